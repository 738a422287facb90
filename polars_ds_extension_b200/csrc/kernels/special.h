// Student-t survival function and quantile for lin_reg_report, host + device.
//
// Reference: student_t_sf / student_t_ppf / checked_beta_reg / inv_beta_reg
// (/root/reference/src/stats_utils/beta.rs:24-37, 62-, 201-361, 365-377 — statrs-derived).  Restated from the
// published definitions: sf(t; v) = 1/2 I_{v/(v+t^2)}(v/2, 1/2) for t >= 0, with the regularised incomplete beta
// evaluated by its continued fraction (modified Lentz), and the quantile by safeguarded Newton on sf.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define PDSB_HD __host__ __device__
#else
#define PDSB_HD
#endif

namespace pdsb {

PDSB_HD inline double betacf(double a, double b, double x) {
  const double FPMIN = 1e-300, EPS = 1e-16;
  double qab = a + b, qap = a + 1.0, qam = a - 1.0;
  double c = 1.0, d = 1.0 - qab * x / qap;
  if (fabs(d) < FPMIN) d = FPMIN;
  d = 1.0 / d;
  double h = d;
  for (int m = 1; m <= 200000; ++m) {
    int m2 = 2 * m;
    double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d; h *= d * c;
    aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
    d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
    c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (fabs(del - 1.0) <= EPS) break;
  }
  return h;
}

// regularised incomplete beta I_x(a, b)
PDSB_HD inline double betai(double a, double b, double x) {
  if (!(x > 0.0)) return 0.0;
  if (!(x < 1.0)) return 1.0;
  double ln_bt = lgamma(a + b) - lgamma(a) - lgamma(b) + a * log(x) + b * log1p(-x);
  double bt = exp(ln_bt);
  if (x < (a + 1.0) / (a + b + 2.0)) return bt * betacf(a, b, x) / a;
  return 1.0 - bt * betacf(b, a, 1.0 - x) / b;
}

// P(T > x) for Student-t with df degrees of freedom
PDSB_HD inline double student_t_sf(double x, double df) {
  if (!(df > 0.0) || x != x) return nan("");
  if (isinf(df)) return 0.5 * erfc(x / sqrt(2.0));
  double h = df / (df + x * x);
  double ib = 0.5 * betai(0.5 * df, 0.5, h);
  return (x <= 0.0) ? 1.0 - ib : ib;
}

PDSB_HD inline double student_t_pdf(double x, double df) {
  double ln = lgamma(0.5 * (df + 1.0)) - lgamma(0.5 * df) - 0.5 * log(df * 3.14159265358979323846) -
              0.5 * (df + 1.0) * log1p(x * x / df);
  return exp(ln);
}

// upper quantile: t such that cdf(t) = prob, for prob in (0.5, 1)
PDSB_HD inline double student_t_ppf(double prob, double df) {
  if (!(df > 0.0)) return nan("");
  const double target = 1.0 - prob;   // sf(t) = target
  double lo = 0.0, hi = 2.0;
  int guard = 0;
  while (student_t_sf(hi, df) > target && guard++ < 200) { lo = hi; hi *= 2.0; }
  double t = 0.5 * (lo + hi);
  for (int it = 0; it < 200; ++it) {
    double f = student_t_sf(t, df) - target;
    if (f > 0.0) lo = t; else hi = t;
    double pdf = student_t_pdf(t, df);
    double tn = (pdf > 0.0) ? t + f / pdf : 0.5 * (lo + hi);   // Newton: d sf/dt = -pdf
    if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
    if (fabs(tn - t) <= 1e-15 * fabs(t)) { t = tn; break; }
    t = tn;
    if (hi - lo <= 1e-15 * hi) break;
  }
  return t;
}

}  // namespace pdsb
