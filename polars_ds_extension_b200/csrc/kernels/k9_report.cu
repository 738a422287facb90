// K9 — lin_reg_report / wls_report statistics, all on the device.
//
// Reference: pl_lin_reg_report / pl_wls_report (/root/reference/src/num_ext/linear_regression.rs:822-1117).
// The reference materialises the p x N matrix (X'X)^-1 X' (:857) and N x N diagonal products for HC0-HC3
// (:880-909).  Here:   moments (K2) -> explicit inverse + beta (K3, method INV) -> residual pass (K4, SSR) ->
// [HC only] per-row weight  omega_i = e_i^2 * {1, 1/(1-h_i), 1/(1-h_i)^2},  h_i = x_i G^-1 x_i^T  (this file) ->
// "meat" X' diag(omega) X as a weighted moments pass (K2a) -> p-length epilogue: G^-1 meat G^-1, SE, t,
// p = 2 sf(|t|, dof), CI = beta -/+ t_{0.975,dof} SE, r2 = 1 - SSR/(var_y * n) (the reference multiplies the
// ddof=1 variance by n, :867 — kept), adj_r2.
#include "../common.h"
#include "kernels.h"
#include "special.h"

namespace pdsb {

namespace {

template <typename T>
__global__ void __launch_bounds__(256)
hc_weights_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ resid, const T* __restrict__ mask,
                  int64_t n, int p, int add_bias, int se_type, const double* __restrict__ ginv /* q x q */,
                  T* __restrict__ omega) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* gi = reinterpret_cast<T*>(smem_raw);
  const int q = p + (add_bias ? 1 : 0);
  for (int i = threadIdx.x; i < q * q; i += blockDim.x) gi[i] = (T)ginv[i];
  __syncthreads();
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
    T e = resid[r];
    T wgt = e * e;
    if (se_type >= 3) {   // hc2, hc3 need the leverage
      T h = T(0);
      for (int a = 0; a < q; ++a) {
        T xa = (a < p) ? X[(int64_t)a * ldx + r] : T(1);
        T inner = T(0);
        for (int b = 0; b < q; ++b) {
          T xb = (b < p) ? X[(int64_t)b * ldx + r] : T(1);
          inner = fma(gi[a * q + b], xb, inner);
        }
        h = fma(xa, inner, h);
      }
      T d = T(1) - h;
      wgt = (se_type == 4) ? wgt / (d * d) : wgt / d;
    }
    if (mask && mask[r] == T(0)) wgt = T(0);
    omega[r] = wgt;
  }
}

struct FinalizeArgs {
  const double* nvalid;  // number of rows that entered the fit
  const double* beta;    // q
  const double* ginv;    // q x q
  const double* ssr;     // [0] unweighted, [4] weighted (layout of predict's ssr buffer: 8 doubles)
  const double* meat;    // moments with omega weights (q1 x q1) or null
  int p, add_bias, se_type, weighted;
  double y_var;
  double* out;           // 8 x q
};

__global__ void report_finalize_kernel(FinalizeArgs a) {
  const int p = a.p, q = p + (a.add_bias ? 1 : 0), q1 = p + 2;
  const int i = threadIdx.x;
  __shared__ double t_alpha;
  const double n = *a.nvalid;
  const double dof = n - (double)q;
  if (i == 0) t_alpha = student_t_ppf(0.975, dof);
  __syncthreads();
  if (i >= q) return;
  const double ssr = a.ssr[0];
  const double ratio = ssr / (a.y_var * n);
  const double r2 = 1.0 - ratio;
  const double adj_r2 = 1.0 - ratio * ((n - 1.0) / (dof - 1.0));
  double se;
  if (a.se_type == 0 || a.weighted) {
    const double mse = (a.weighted ? a.ssr[4] : ssr) / dof;
    se = sqrt(mse * a.ginv[(size_t)i * q + i]);
  } else {
    auto fz = [&](int c) { return c < p ? c : p + 1; };
    double v = 0.0;
    for (int r = 0; r < q; ++r) {
      double inner = 0.0;
      for (int c = 0; c < q; ++c) inner += a.meat[(size_t)fz(r) * q1 + fz(c)] * a.ginv[(size_t)c * q + i];
      v += a.ginv[(size_t)i * q + r] * inner;
    }
    if (a.se_type == 2) v *= n / (n - (double)q);
    se = sqrt(v);
  }
  const double b = a.beta[i];
  const double t = b / se;
  const double pval = 2.0 * student_t_sf(fabs(t), dof);
  double* o = a.out;
  o[0 * q + i] = b;
  o[1 * q + i] = se;
  o[2 * q + i] = t;
  o[3 * q + i] = pval;
  o[4 * q + i] = b - t_alpha * se;
  o[5 * q + i] = b + t_alpha * se;
  o[6 * q + i] = r2;
  o[7 * q + i] = adj_r2;
}

}  // namespace

template <typename T>
int report_stats(const T* X, int64_t ldx, const T* y, const T* w, const T* mask, int64_t n, int p, int add_bias,
                 int se_type, double y_var, double* out, cudaStream_t s) {
  const int q = p + (add_bias ? 1 : 0), q1 = p + 2;
  if (q > 512) { set_error("report: too many features"); return 1; }
  char* base = nullptr;
  // M (q1^2) | meat (q1^2) | beta (q) | ginv (q^2) | ssr (8) | status
  size_t nd = 2 * (size_t)q1 * q1 + q + (size_t)q * q + 8 + 4;
  if (dev_alloc((void**)&base, nd * sizeof(double), s)) return 1;
  double* M = reinterpret_cast<double*>(base);
  double* meat = M + (size_t)q1 * q1;
  double* beta = meat + (size_t)q1 * q1;
  double* ginv = beta + q;
  double* ssr = ginv + (size_t)q * q;
  double* nvalid = ssr + 8;
  int* status = reinterpret_cast<int*>(ssr + 10);
  T* pred = nullptr; T* resid = nullptr; T* omega = nullptr;
  const int64_t ldo = (n + 3) & ~int64_t(3);
  int rc = 0;
  auto cleanup = [&]() {
    if (pred) dev_free(pred, s);
    if (resid) dev_free(resid, s);
    if (omega) dev_free(omega, s);
    dev_free(base, s);
  };
  if (dev_alloc((void**)&pred, (size_t)ldo * sizeof(T), s) || dev_alloc((void**)&resid, (size_t)ldo * sizeof(T), s)) { cleanup(); return 1; }
  rc = mask ? count_mask<T>(mask, n, nvalid, s) : fill_value<double>(nvalid, 1, (double)n, s);
  if (!rc) rc = moments_simt<T>(X, ldx, y, n, w, mask, n, p, 1, M, s);
  pdsb_solve_opts o{};
  o.p = p; o.t = 1; o.add_bias = add_bias; o.method = PDSB_METHOD_INV; o.solver = PDSB_SOLVER_QR;
  if (!rc) rc = solve_from_moments(M, o, beta, status, ginv, s);
  if (!rc) rc = predict_resid<T>(X, ldx, y, n, w, mask, n, p, 1, add_bias, beta, status, pred, resid, ldo, nullptr, ssr, s);
  const bool hc = (se_type >= 1) && !w;
  if (!rc && hc) {
    if (dev_alloc((void**)&omega, (size_t)ldo * sizeof(T), s)) { cleanup(); return 1; }
    size_t smem = (size_t)q * q * sizeof(T);
    auto k = hc_weights_kernel<T>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 8);
    k<<<grid, 256, smem, s>>>(X, ldx, resid, mask, n, p, add_bias, se_type, ginv, omega);
    cudaError_t e = cudaGetLastError(); count_launch();
    if (e != cudaSuccess) { set_error("hc_weights launch failed: %s", cudaGetErrorString(e)); rc = 1; }
    if (!rc) rc = moments_simt<T>(X, ldx, y, n, omega, mask, n, p, 1, meat, s);
  }
  if (!rc) {
    FinalizeArgs a;
    a.nvalid = nvalid; a.beta = beta; a.ginv = ginv; a.ssr = ssr; a.meat = hc ? meat : nullptr;
    a.p = p; a.add_bias = add_bias; a.se_type = hc ? se_type : 0; a.weighted = w ? 1 : 0; a.y_var = y_var; a.out = out;
    int threads = ((q + 31) / 32) * 32;
    report_finalize_kernel<<<1, threads, 0, s>>>(a);
    cudaError_t e = cudaGetLastError(); count_launch();
    if (e != cudaSuccess) { set_error("report finalize launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  cleanup();
  return rc;
}

template int report_stats<float>(const float*, int64_t, const float*, const float*, const float*, int64_t, int, int,
                                 int, double, double*, cudaStream_t);
template int report_stats<double>(const double*, int64_t, const double*, const double*, const double*, int64_t, int,
                                  int, int, double, double*, cudaStream_t);

}  // namespace pdsb
