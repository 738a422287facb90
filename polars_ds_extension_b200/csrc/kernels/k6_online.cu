// K6 / K7 — rolling_lin_reg and recursive_lin_reg without the sequential Woodbury chain.
//
// Reference: faer_rolling_lr / faer_rolling_skipping_lr / faer_recursive_lr + woodbury_step
// (/root/reference/src/linear/online_lr/lr_online_solvers.rs:148-332) and the output loops of pl_rolling_lr /
// pl_recursive_lr (src/num_ext/linear_regression.rs:1121-1283).  The reference walks the rows strictly one after
// another (2 rank-1 updates per row, ~6 small heap allocations per row, single thread).  What its tests pin is the
// mathematical definition — rolling == per-window OLS/ridge, recursive == prefix OLS/ridge
// (tests/test_linear_exprs.py:123-166, 718-854) — and that is what this kernel evaluates directly and in parallel:
//
//   moments of a row   m(z) = [ z_i z_j (i<=j), z_i y, 1 ]  with z = (x_0..x_{p-1}[,1]);  zero for non-finite rows
//                      (OnlineLR::update skips non-finite rows, lr_online_solvers.rs:85-89)
//   pass A  tile sums   S_k = sum of m over tile k (T_ROWS rows), f64
//   pass B  tile scan   C_k = sum_{j<k} S_j  (exclusive, f64)
//   pass C  per row     W_t = C_k + E(t) - [C_k' + H + L(t-w)]   (f64: the global-prefix difference loses
//                       ~1e-16 * (n/w) relative, harmless), then G = W_GG + lambda I_p, solve by an in-register
//                       Cholesky in the data dtype, pred_t = x_t . beta_t.
// Rolling = both sides, recursive = entering side only (w = infinity).  Every row is independent after pass B.
// Bytes per row (algorithmic): (p+1) s read, (p+bias) s + s + 1 written.
#include "../common.h"
#include "kernels.h"

namespace pdsb {

namespace {

constexpr int THREADS = 128;
constexpr int L_ROWS = 8;
constexpr int T_ROWS = THREADS * L_ROWS;  // 1024 rows per tile
constexpr int BSTRIDE = THREADS + 1;      // padded stride of the scan buffer (doubles)

template <int D> struct MomN { static constexpr int NG = D * (D + 1) / 2; static constexpr int NM = NG + D + 1; };

// load row r -> z[D] (features, 1 for bias), yv; returns finite flag
template <typename T, int D>
__device__ __forceinline__ bool load_row(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int p,
                                         int64_t r, T* z, T& yv) {
  bool fin = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    if (c < p) { z[c] = __ldg(X + (int64_t)c * ldx + r); fin = fin && isfinite(z[c]); }
    else z[c] = T(1);
  }
  yv = __ldg(y + r);
  fin = fin && isfinite(yv);
  return fin;
}

template <typename T, int D>
__device__ __forceinline__ void add_moments(double* W, const T* z, T yv, double sign) {
  int k = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { W[k] += sign * ((double)z[i] * (double)z[j]); ++k; }
#pragma unroll
  for (int i = 0; i < D; ++i) { W[k] += sign * ((double)z[i] * (double)yv); ++k; }
  W[k] += sign;
}

// block reduction / exclusive scan of per-thread NM-vectors through a padded shared buffer
//   buf[c * BSTRIDE + tid]
template <int NM>
__device__ __forceinline__ void store_vec(double* buf, const double* v) {
#pragma unroll
  for (int c = 0; c < NM; ++c) buf[c * BSTRIDE + threadIdx.x] = v[c];
}

// ---------------- pass A ----------------
template <typename T, int D>
__global__ void __launch_bounds__(THREADS)
tile_sums_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p,
                 int64_t ntiles, double* __restrict__ S /* [NM][ntiles] */) {
  constexpr int NM = MomN<D>::NM;
  extern __shared__ double buf[];
  const int64_t k = blockIdx.x;
  const int64_t r0 = k * T_ROWS;
  double v[NM];
#pragma unroll
  for (int c = 0; c < NM; ++c) v[c] = 0.0;
  for (int j = 0; j < L_ROWS; ++j) {
    int64_t r = r0 + (int64_t)j * THREADS + threadIdx.x;   // strided: coalesced
    if (r < n) {
      T z[D]; T yv;
      if (load_row<T, D>(X, ldx, y, p, r, z, yv)) add_moments<T, D>(v, z, yv, 1.0);
    }
  }
  store_vec<NM>(buf, v);
  __syncthreads();
  for (int c = threadIdx.x; c < NM; c += THREADS) {
    double s = 0.0;
    for (int i = 0; i < THREADS; ++i) s += buf[c * BSTRIDE + i];
    S[(size_t)c * ntiles + k] = s;
  }
}

// ---------------- pass B: exclusive scan along tiles, one block per component ----------------
__global__ void __launch_bounds__(1024) tile_scan_kernel(double* __restrict__ S, int64_t ntiles) {
  __shared__ double warp_tot[32];
  __shared__ double carry;
  double* row = S + (size_t)blockIdx.x * ntiles;
  if (threadIdx.x == 0) carry = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < ntiles; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    double v = (i < ntiles) ? row[i] : 0.0;
    double inc = v;
    for (int off = 1; off < 32; off <<= 1) { double t = __shfl_up_sync(0xffffffffu, inc, off); if ((threadIdx.x & 31) >= off) inc += t; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    double wpre = 0.0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wpre += warp_tot[w];
    double excl = carry + wpre + (inc - v);
    __syncthreads();
    if (i < ntiles) row[i] = excl;
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
}

// ---------------- in-register Cholesky solve ----------------
template <typename T, int D>
__device__ __forceinline__ bool chol_solve_reg(const double* W, int p, T lambda, T* beta) {
  constexpr int NG = MomN<D>::NG;
  T A[D][D];
  int k = 0;
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = i; j < D; ++j) { A[i][j] = (T)W[k]; A[j][i] = A[i][j]; ++k; }
#pragma unroll
  // The reference builds OnlineLR::new(lambda, false) on a matrix that already holds the physical ones column
  // (lr_online_solvers.rs:163-165, 195-197), so lambda lands on EVERY diagonal entry, the bias one included.
  for (int i = 0; i < D; ++i) { A[i][i] += lambda; beta[i] = (T)W[NG + i]; }
  (void)p;
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    T d = A[c][c];
    if (!(d > T(0)) || !isfinite(d)) ok = false;
    T inv = rsqrt(d);
    A[c][c] = d * inv;   // sqrt(d)
#pragma unroll
    for (int i = c + 1; i < D; ++i) A[i][c] *= inv;
#pragma unroll
    for (int j = c + 1; j < D; ++j)
#pragma unroll
      for (int i = j; i < D; ++i) A[i][j] -= A[i][c] * A[j][c];
  }
  // forward / backward substitution
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = beta[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= A[i][j] * beta[j];
    beta[i] = s / A[i][i];
  }
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    T s = beta[i];
#pragma unroll
    for (int j = i + 1; j < D; ++j) s -= A[j][i] * beta[j];
    beta[i] = s / A[i][i];
  }
  return ok;
}

// ---------------- pass C ----------------
template <typename T, int D>
__global__ void __launch_bounds__(THREADS)
online_main_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p,
                   int64_t window, int64_t min_rows, int skip, T lambda, int64_t ntiles,
                   const double* __restrict__ C /* [NM][ntiles] exclusive tile prefixes */,
                   T* __restrict__ coeffs, T* __restrict__ pred, uint8_t* __restrict__ valid) {
  constexpr int NM = MomN<D>::NM;
  extern __shared__ double buf[];
  __shared__ double base_sh[NM];
  const int64_t k = blockIdx.x;
  const int64_t t0 = k * T_ROWS + (int64_t)threadIdx.x * L_ROWS;   // this thread's first row
  const bool rolling = window > 0;
  // leaving side geometry (rows u = t - window)
  const int64_t lo = rolling ? max((int64_t)0, k * T_ROWS - window) : 0;   // first leaving row that matters (>=0)
  const int64_t kl = lo / T_ROWS;                                          // tile of `lo`
  const int64_t a0 = kl * T_ROWS;
  const bool any_leave = rolling && (k * T_ROWS + T_ROWS - 1 - window >= 0);

  // ---- head total H = sum rows [a0, lo) ----
  double v[NM];
  if (any_leave) {
#pragma unroll
    for (int c = 0; c < NM; ++c) v[c] = 0.0;
    for (int j = 0; j < L_ROWS; ++j) {
      int64_t r = a0 + (int64_t)j * THREADS + threadIdx.x;
      if (r < lo) {
        T z[D]; T yv;
        if (load_row<T, D>(X, ldx, y, p, r, z, yv)) add_moments<T, D>(v, z, yv, 1.0);
      }
    }
    store_vec<NM>(buf, v);
    __syncthreads();
    for (int c = threadIdx.x; c < NM; c += THREADS) {
      double s = 0.0;
      for (int i = 0; i < THREADS; ++i) s += buf[c * BSTRIDE + i];
      base_sh[c] = C[(size_t)c * ntiles + k] - (C[(size_t)c * ntiles + kl] + s);
    }
    __syncthreads();
  } else {
    for (int c = threadIdx.x; c < NM; c += THREADS) base_sh[c] = C[(size_t)c * ntiles + k];
    __syncthreads();
  }

  // ---- per-thread delta = sum(entering rows) - sum(leaving rows) over its L rows ----
#pragma unroll
  for (int c = 0; c < NM; ++c) v[c] = 0.0;
  for (int j = 0; j < L_ROWS; ++j) {
    int64_t r = t0 + j;
    if (r < n) {
      T z[D]; T yv;
      if (load_row<T, D>(X, ldx, y, p, r, z, yv)) add_moments<T, D>(v, z, yv, 1.0);
      if (rolling) {
        int64_t u = r - window;
        if (u >= 0) { if (load_row<T, D>(X, ldx, y, p, u, z, yv)) add_moments<T, D>(v, z, yv, -1.0); }
      }
    }
  }
  store_vec<NM>(buf, v);
  __syncthreads();
  // exclusive scan along threads, one component per scanning thread
  for (int c = threadIdx.x; c < NM; c += THREADS) {
    double run = base_sh[c];
    for (int i = 0; i < THREADS; ++i) { double x = buf[c * BSTRIDE + i]; buf[c * BSTRIDE + i] = run; run += x; }
  }
  __syncthreads();
  double W[NM];
#pragma unroll
  for (int c = 0; c < NM; ++c) W[c] = buf[c * BSTRIDE + threadIdx.x];

  // ---- walk the L rows ----
  for (int j = 0; j < L_ROWS; ++j) {
    int64_t r = t0 + j;
    if (r >= n) break;
    T z[D]; T yv; T zl[D]; T yl;
    bool fin = load_row<T, D>(X, ldx, y, p, r, z, yv);
    if (fin) add_moments<T, D>(W, z, yv, 1.0);
    if (rolling) {
      int64_t u = r - window;
      if (u >= 0) { if (load_row<T, D>(X, ldx, y, p, u, zl, yl)) add_moments<T, D>(W, zl, yl, -1.0); }
    }
    const double cnt = W[NM - 1];
    bool ok;
    if (rolling) ok = (r >= window - 1) && (!skip || cnt >= (double)min_rows - 0.5);
    else ok = skip ? (fin && cnt >= (double)min_rows - 0.5) : (r >= min_rows - 1);
    T beta[D];
    T pr = T(0);
    if (ok) {
      bool pd = chol_solve_reg<T, D>(W, p, lambda, beta);
      if (!pd) {
#pragma unroll
        for (int i = 0; i < D; ++i) beta[i] = (T)nan("");
      }
#pragma unroll
      for (int i = 0; i < D; ++i) pr = fma(z[i], beta[i], pr);
    } else {
#pragma unroll
      for (int i = 0; i < D; ++i) beta[i] = T(0);
    }
#pragma unroll
    for (int i = 0; i < D; ++i) coeffs[r * D + i] = beta[i];
    pred[r] = pr;
    valid[r] = ok ? 1 : 0;
  }
}

template <typename T, int D>
int run_online(const T* X, int64_t ldx, const T* y, int64_t n, int p, int64_t window, int64_t min_rows, int skip,
               double lambda, T* coeffs, T* pred, uint8_t* valid, cudaStream_t s) {
  constexpr int NM = MomN<D>::NM;
  const int64_t ntiles = ceil_div(n, T_ROWS);
  double* S = nullptr;
  if (dev_alloc((void**)&S, (size_t)NM * ntiles * sizeof(double), s)) return 1;
  const size_t smem = (size_t)NM * BSTRIDE * sizeof(double);
  auto ka = tile_sums_kernel<T, D>;
  auto kc = online_main_kernel<T, D>;
  if (smem > 48 * 1024) {
    PDSB_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PDSB_CUDA_OK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  ka<<<(unsigned)ntiles, THREADS, smem, s>>>(X, ldx, y, n, p, ntiles, S);
  cudaError_t e = cudaGetLastError(); count_launch();
  if (e == cudaSuccess) { tile_scan_kernel<<<NM, 1024, 0, s>>>(S, ntiles); e = cudaGetLastError(); count_launch(); }
  if (e == cudaSuccess) {
    kc<<<(unsigned)ntiles, THREADS, smem, s>>>(X, ldx, y, n, p, window, min_rows, skip, (T)lambda, ntiles, S, coeffs, pred, valid);
    e = cudaGetLastError(); count_launch();
  }
  dev_free(S, s);
  if (e != cudaSuccess) { set_error("online lin_reg launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

template <typename T>
int online_lin_reg(const T* X, int64_t ldx, const T* y, int64_t n, int p, int add_bias, int64_t window,
                   int64_t min_rows, int skip, double lambda, T* coeffs, T* pred, uint8_t* valid, cudaStream_t s) {
  if (n <= 0) return 0;
  const int d = p + (add_bias ? 1 : 0);
  if (n / T_ROWS > 2000000000LL) { set_error("online lin_reg: too many rows"); return 1; }
#define CASE_D(DD) case DD: return run_online<T, DD>(X, ldx, y, n, p, window, min_rows, skip, lambda, coeffs, pred, valid, s);
  switch (d) {
    CASE_D(1) CASE_D(2) CASE_D(3) CASE_D(4) CASE_D(5) CASE_D(6) CASE_D(7) CASE_D(8) CASE_D(9) CASE_D(10)
    CASE_D(11) CASE_D(12)
    default:
      set_error("rolling/recursive lin_reg: %d coefficients not supported on device (max 12)", d);
      return 1;
  }
#undef CASE_D
}

template int online_lin_reg<float>(const float*, int64_t, const float*, int64_t, int, int, int64_t, int64_t, int,
                                   double, float*, float*, uint8_t*, cudaStream_t);
template int online_lin_reg<double>(const double*, int64_t, const double*, int64_t, int, int, int64_t, int64_t, int,
                                    double, double*, double*, uint8_t*, cudaStream_t);

}  // namespace pdsb
