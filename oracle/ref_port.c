/* oracle/ref_port.c — TEST / BASELINE INFRASTRUCTURE, never linked into the product.
 *
 * C (OpenMP) restatement of the DATA PASSES of the reference's `pl_lr_pred` on its no-null fast path, used as the timed
 * CPU arm (`bench.py --impl reference`, `cpu_baseline`) and cross-checked against the numpy oracle in
 * tests/test_ref_port.py.  The reference is Rust on un-vendored crates (faer, rayon) and cannot be built in this image
 * (DESIGN.md §5): this is kind = "port".  Thread structure follows the reference, not what would be fastest:
 *
 *   pack      series_to_slice_inner fast path (/root/reference/src/utils/mod.rs:118-132): ONE thread extends a fresh
 *             Vec column by column (sequential memcpy + first-touch page faults), + the physical ones column
 *             (src/num_ext/linear_regression.rs:180-182);
 *   gram      get_xtx_with_lambda (src/linear/lr/lr_solvers.rs:183-211): faer matmul, Par::rayon(0) -> all threads;
 *   xty       build_xty (:262-278): faer matmul, all threads;
 *   (solve    col-piv QR of the q x q Gram + rank gate (:329-382): microseconds, done by the caller in numpy)
 *   predict   `x * &coeffs` (linear_regression.rs:782): faer matmul -> all threads;
 *   resid     `y - &pred` (:783): faer's elementwise zip is sequential -> ONE thread;
 *   output    Float32Chunked::from_slice for pred and resid (:790-800): ONE thread copies each into a new buffer.
 *
 * f32 arithmetic with f32 accumulation inside a thread's row block (what an f32 GEMM does), summed across blocks.
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/ref_port.c -o oracle/_ref/libref_port.so   (oracle/Makefile)
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  float* z;        /* packed column-major [n x (1 + q)]: y, x_1 .. x_p, (ones) */
  int64_t n;
  int p, q;        /* q = p + add_bias */
} ref_packed;

double ref_now(void) { return omp_get_wtime(); }
int ref_max_threads(void) { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* pack: cols[0] = y, cols[1..p] = features; sequential like the reference's fast path */
ref_packed* ref_pack_f32(const float* const* cols, int p, int64_t n, int add_bias) {
  ref_packed* P = (ref_packed*)malloc(sizeof(ref_packed));
  if (!P) return NULL;
  P->n = n; P->p = p; P->q = p + (add_bias ? 1 : 0);
  P->z = (float*)malloc((size_t)n * (size_t)(1 + P->q) * sizeof(float));
  if (!P->z) { free(P); return NULL; }
  for (int c = 0; c <= p; ++c) memcpy(P->z + (size_t)c * n, cols[c], (size_t)n * sizeof(float));
  if (add_bias) { float* o = P->z + (size_t)(1 + p) * n; for (int64_t i = 0; i < n; ++i) o[i] = 1.0f; }
  return P;
}

void ref_free(ref_packed* P) { if (P) { free(P->z); free(P); } }

/* xtx [q x q] row-major, xty [q]; all threads, row-blocked */
void ref_gram_f32(const ref_packed* P, float* xtx, float* xty) {
  const int q = P->q;
  const int64_t n = P->n;
  const float* y = P->z;
  const float* X = P->z + n;
  const int nt = omp_get_max_threads();
  float* part = (float*)calloc((size_t)nt * (size_t)(q * q + q), sizeof(float));
#pragma omp parallel
  {
    const int t = omp_get_thread_num();
    float* g = part + (size_t)t * (size_t)(q * q + q);
    float* gy = g + q * q;
    enum { RB = 512 };
#pragma omp for schedule(static)
    for (int64_t r0 = 0; r0 < n; r0 += RB) {
      const int64_t r1 = r0 + RB < n ? r0 + RB : n;
      for (int i = 0; i < q; ++i) {
        const float* xi = X + (size_t)i * n;
        for (int j = i; j < q; ++j) {
          const float* xj = X + (size_t)j * n;
          float s = 0.0f;
          for (int64_t r = r0; r < r1; ++r) s += xi[r] * xj[r];
          g[i * q + j] += s;
        }
        float s = 0.0f;
        for (int64_t r = r0; r < r1; ++r) s += xi[r] * y[r];
        gy[i] += s;
      }
    }
  }
  for (int i = 0; i < q * q + q; ++i) {
    float s = 0.0f;
    for (int t = 0; t < nt; ++t) s += part[(size_t)t * (size_t)(q * q + q) + i];
    if (i < q * q) xtx[i] = s; else xty[i - q * q] = s;
  }
  for (int i = 0; i < q; ++i) for (int j = 0; j < i; ++j) xtx[i * q + j] = xtx[j * q + i];
  free(part);
}

/* pred = X beta (all threads), resid = y - pred (one thread), then the two from_slice copies (one thread each) */
void ref_predict_f32(const ref_packed* P, const float* beta, float* pred_out, float* resid_out) {
  const int q = P->q;
  const int64_t n = P->n;
  const float* y = P->z;
  const float* X = P->z + n;
  float* pred = (float*)malloc((size_t)n * sizeof(float));
  float* resid = (float*)malloc((size_t)n * sizeof(float));
#pragma omp parallel for schedule(static)
  for (int64_t r0 = 0; r0 < n; r0 += 4096) {
    const int64_t r1 = r0 + 4096 < n ? r0 + 4096 : n;
    for (int64_t r = r0; r < r1; ++r) pred[r] = 0.0f;
    for (int j = 0; j < q; ++j) {
      const float b = beta[j];
      const float* xj = X + (size_t)j * n;
      for (int64_t r = r0; r < r1; ++r) pred[r] += xj[r] * b;
    }
  }
  for (int64_t r = 0; r < n; ++r) resid[r] = y[r] - pred[r];
  memcpy(pred_out, pred, (size_t)n * sizeof(float));
  memcpy(resid_out, resid, (size_t)n * sizeof(float));
  free(pred); free(resid);
}
