// K6 / K7 — rolling_lin_reg and recursive_lin_reg without the sequential Woodbury chain.
//
// Reference: faer_rolling_lr / faer_rolling_skipping_lr / faer_recursive_lr + woodbury_step
// (/root/reference/src/linear/online_lr/lr_online_solvers.rs:148-332) and the output loops of pl_rolling_lr /
// pl_recursive_lr (src/num_ext/linear_regression.rs:1121-1283).  The reference walks the rows strictly one after
// another (2 rank-1 updates per row, ~6 small heap allocations per row, single thread).  What its tests pin is the
// mathematical definition — rolling == per-window OLS/ridge, recursive == prefix OLS/ridge
// (tests/test_linear_exprs.py:123-166, 718-854) — and that is what this kernel evaluates directly and in parallel:
//
//   row vector          e = (z_0..z_{D-1}, y, 1)  with z = (x_0..x_{p-1}[,1]);  all zero for a non-finite row
//                       (OnlineLR::update skips non-finite rows, lr_online_solvers.rs:85-89)
//   moments of a row    m = [ e_i e_j (i<=j<D), e_i y, 1*1 ]          NM = D(D+1)/2 + D + 1 numbers
//   pass A  chain sums  S_k = sum of m over chain k (1024 rows), f64
//   pass B  chain scan  C_k = sum_{j<k} S_j  (exclusive, f64)
//   pass C  per row     W_t = C_k + E(t) - [C_k' + H + L(t-w)]   (f64: the global-prefix difference loses
//                       ~1e-16 * (n/w) relative, harmless), then G = W_GG + lambda I, solved by an in-register
//                       Cholesky in the data dtype, pred_t = x_t . beta_t.
// Rolling = both sides, recursive = entering side only (w = infinity).  Every chain is independent after pass B.
//
// Work decomposition (passes A and C): ONE WARP PER CHAIN, and inside the warp two alternating roles per 32-row batch
//   * lane-per-row:       coalesced loads of the batch's entering (and leaving) rows, e-vectors written to shared
//                         memory as doubles; later the Cholesky solve of "its" row from the staged Gram, and the stores;
//   * lane-per-component: each lane owns <= 3 of the NM running moments in f64 registers and walks the 32 rows
//                         (one LDS.64 + one LDS.128 feed two DFMAs per row and side), dropping W_t (cast to the data
//                         dtype) into shared memory for the row's solver lane.
// Pass A has no per-row output and stays thread-per-row (NM register accumulators, one shuffle reduction per chain).
// This keeps the running state at 2-3 doubles per lane (the thread-per-row-chain version of this kernel carried all
// NM in registers: 254 registers, 8 warps per SM, 29 % issue utilisation — profiles/README.md) and needs no
// intra-tile scan at all.
// Bytes per row (algorithmic): (p+1) s read, (p+bias) s + s + 1 written.
#include "../common.h"
#include "kernels.h"

namespace pdsb {

namespace {

constexpr int WARPS = 5;                 // independent chains per CTA (no block-level synchronisation anywhere);
                                         // 3 CTAs x 5 warps fit the 8-feature + bias f32 case in shared memory and registers
constexpr int CTA_THREADS = WARPS * 32;
constexpr int BATCH = 32;                // rows per role switch
constexpr int CHAIN_ROWS = 1024;         // rows per warp = granularity of the prefix arrays

// A lane-per-component "task" (a, b) accumulates e_a*e_b and e_a*e_{b+1} (b even): one LDS.64 + one LDS.128 feed two
// DFMAs.  Row a of the upper triangle needs b = a&~1, .., <= D (the y column is entry D); one more task holds 1*1.
template <int D> constexpr int n_tasks() {
  int t = 0;
  for (int a = 0; a < D; ++a) for (int b = a & ~1; b <= D; b += 2) ++t;
  return t + 1;
}
constexpr int even_odd_half(int x) { int e = (x + 1) & ~1; return ((e / 2) & 1) ? e : e + 2; }   // even, half of it odd

template <int D> struct MomN {
  static constexpr int NG = D * (D + 1) / 2;
  static constexpr int NM = NG + D + 1;
  static constexpr int ES = even_odd_half(((D + 1) | 1) + 1);   // 16-byte pairs; <= 2-way conflicts when staging
  static constexpr int GS = (NM + 1) | 1;                  // odd stride; slot NM is the dump slot of unused products
  static constexpr int NT = n_tasks<D>();
  static constexpr int TPL = (NT + 31) / 32;               // tasks per lane
};

template <typename T, int D>
constexpr size_t warp_smem_bytes() {
  return (size_t)(2 * BATCH * MomN<D>::ES + BATCH) * sizeof(double) + (size_t)BATCH * MomN<D>::GS * sizeof(T);
}

// packed slot of the unordered pair (i <= j) in the solver's Gram order, or the dump slot
template <int D>
__device__ __forceinline__ int pair_slot(int i, int j) {
  constexpr int NG = MomN<D>::NG, NM = MomN<D>::NM;
  if (i > j || i >= D) return (i == D + 1 && j == D + 1) ? NM - 1 : NM;
  if (j < D) return i * D - i * (i - 1) / 2 + (j - i);
  if (j == D) return NG + i;
  return NM;
}

// task t -> offsets (in doubles) of e_a, e_b inside a staged row and the Gram slots of its two products
template <int D>
__device__ __forceinline__ void task_of(int t, int& a, int& b, int& k0, int& k1) {
  constexpr int NT = MomN<D>::NT;
  a = D + 1; b = (D + 1) & ~1;
  if (t < NT - 1) {
    int cur = 0;
    for (int aa = 0; aa < D; ++aa)
      for (int bb = aa & ~1; bb <= D; bb += 2) { if (cur == t) { a = aa; b = bb; } ++cur; }
  }
  if (t >= NT) { k0 = k1 = MomN<D>::NM; return; }
  k0 = pair_slot<D>(a, b);
  k1 = pair_slot<D>(a, b + 1);
}

// lane-per-row: raw loads of row r (clamped address, no branch); `inr` says whether the row exists
template <typename T, int D>
__device__ __forceinline__ void load_raw(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int p,
                                         int64_t r, int64_t n, T* z, T& yv) {
  const int64_t rc = min(max(r, (int64_t)0), n - 1);
  const T* q = X + rc;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    if (c < p) { z[c] = __ldg(q); q += ldx; }
    else z[c] = T(1);
  }
  yv = __ldg(y + rc);
}

// lane-per-row: e-vector of a row (all zero when the row is missing or not finite) -> dst[0..D+2)
template <typename T, int D>
__device__ __forceinline__ bool put_row(double* __restrict__ dst, const T* z, T yv, bool inr) {
  T acc = yv * T(0);
#pragma unroll
  for (int c = 0; c < D; ++c) acc = fma(z[c], T(0), acc);     // 0 when every entry is finite, NaN otherwise
  const bool fin = inr && (acc == T(0));
#pragma unroll
  for (int c = 0; c < D; ++c) dst[c] = (double)(fin ? z[c] : T(0));
  dst[D] = (double)(fin ? yv : T(0));
  dst[D + 1] = fin ? 1.0 : 0.0;
  if (((D + 1) | 1) != D + 1) dst[(D + 1) | 1] = 0.0;
  return fin;
}

// lane-per-component: walk the 32 staged rows
template <typename T, int D, bool BOTH, bool STORE>
__device__ __forceinline__ void walk_batch(const double* __restrict__ ee, const double* __restrict__ el,
                                           const int* ta, const int* tb, const int* k0, const int* k1,
                                           double (*W)[2], T* __restrict__ gs, double* __restrict__ cnt, bool cnt_lane) {
  constexpr int ES = MomN<D>::ES, GS = MomN<D>::GS, TPL = MomN<D>::TPL, NT = MomN<D>::NT;
#pragma unroll
  for (int r = 0; r < BATCH; ++r) {
    const double* e = ee + r * ES;
    const double* l = el + r * ES;
#pragma unroll
    for (int m = 0; m < TPL; ++m) {
      const double ea = e[ta[m]];
      const double2 eb = *reinterpret_cast<const double2*>(e + tb[m]);
      W[m][0] = fma(ea, eb.x, W[m][0]);
      W[m][1] = fma(ea, eb.y, W[m][1]);
      if (BOTH) {
        const double la = l[ta[m]];
        const double2 lb = *reinterpret_cast<const double2*>(l + tb[m]);
        W[m][0] = fma(-la, lb.x, W[m][0]);
        W[m][1] = fma(-la, lb.y, W[m][1]);
      }
    }
    if (STORE) {
#pragma unroll
      for (int m = 0; m < TPL; ++m) {
        gs[r * GS + k0[m]] = (T)W[m][0];
        gs[r * GS + k1[m]] = (T)W[m][1];
      }
      if (cnt_lane) cnt[r] = W[(NT - 1) / 32][(D + 1) & 1];     // the row count stays exact in f64
    }
  }
}

// ---------------- pass A: chain sums.  No per-row output here, so thread-per-row with all NM accumulators in
// registers is the cheap way: lane l takes rows chain0 + 32 j + l, one shuffle reduction per chain. ----------------
template <typename T, int D>
__global__ void __launch_bounds__(CTA_THREADS)
chain_sums_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p,
                  int64_t nchains, double* __restrict__ S /* [NM][nchains] */) {
  constexpr int NM = MomN<D>::NM;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * WARPS + wid;
  if (k >= nchains) return;
  double v[NM];
#pragma unroll
  for (int c = 0; c < NM; ++c) v[c] = 0.0;
  const int64_t chain0 = k * CHAIN_ROWS;
#pragma unroll 2
  for (int j = 0; j < CHAIN_ROWS / 32; ++j) {
    const int64_t r = chain0 + (int64_t)j * 32 + lane;
    T z[D]; T yv;
    load_raw<T, D>(X, ldx, y, p, r, n, z, yv);
    T acc = yv * T(0);
#pragma unroll
    for (int c = 0; c < D; ++c) acc = fma(z[c], T(0), acc);
    const bool fin = (r < n) && (acc == T(0));
    double dz[D];
#pragma unroll
    for (int c = 0; c < D; ++c) dz[c] = fin ? (double)z[c] : 0.0;
    const double dy = fin ? (double)yv : 0.0;
    int q = 0;
#pragma unroll
    for (int i = 0; i < D; ++i)
#pragma unroll
      for (int jj = i; jj < D; ++jj) { v[q] = fma(dz[i], dz[jj], v[q]); ++q; }
#pragma unroll
    for (int i = 0; i < D; ++i) { v[q] = fma(dz[i], dy, v[q]); ++q; }
    v[q] += fin ? 1.0 : 0.0;
  }
#pragma unroll
  for (int c = 0; c < NM; ++c) {
    double x = v[c];
#pragma unroll
    for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    if (lane == (c & 31)) S[(size_t)c * nchains + k] = x;
  }
}

// ---------------- pass B: exclusive scan along tiles, one block per component ----------------
// M0 (optional): moments [X | y | 1]' [X | y | 1] ((p+2)^2, row-major f64) of the rows that precede this shard; they
// seed the prefix so that a row shard continues the expanding fit of the shards before it.
__global__ void __launch_bounds__(1024) tile_scan_kernel(double* __restrict__ S, int64_t ntiles,
                                                         const double* __restrict__ M0, int p, int d) {
  __shared__ double warp_tot[32];
  __shared__ double carry;
  double* row = S + (size_t)blockIdx.x * ntiles;
  if (threadIdx.x == 0) {
    double c0 = 0.0;
    if (M0) {
      // component blockIdx.x -> (i, j) over (z_0..z_{d-1}, y, 1);  z_c = x_c for c < p, the ones column otherwise
      const int ng = d * (d + 1) / 2;
      int c = blockIdx.x, i, j;
      if (c >= ng + d) { i = j = d + 1; }
      else if (c >= ng) { i = c - ng; j = d; }
      else { i = 0; while (c >= d - i) { c -= d - i; ++i; } j = i + c; }
      auto col = [&](int a) { return a < p ? a : (a == d ? p : p + 1); };   // y -> p, bias / ones -> p + 1
      c0 = M0[(size_t)col(i) * (p + 2) + col(j)];
    }
    carry = c0;
  }
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

// ---------------- in-register Cholesky solve on the packed Gram ----------------
// g: NM values in the component order (upper triangle row-major, then X'y, then the count).  lower(i, j) for i >= j
// lives at the slot of (j, i).  The diagonal is overwritten with 1/sqrt(pivot), so the substitutions multiply.
template <int D> __device__ __forceinline__ constexpr int gidx(int i, int j) { return j * D - j * (j - 1) / 2 + (i - j); }

template <typename T, int D>
__device__ __forceinline__ bool chol_solve_packed(T* g, T lambda, T* beta) {
  constexpr int NG = MomN<D>::NG;
  // The reference builds OnlineLR::new(lambda, false) on a matrix that already holds the physical ones column
  // (lr_online_solvers.rs:163-165, 195-197), so lambda lands on EVERY diagonal entry, the bias one included.
#pragma unroll
  for (int i = 0; i < D; ++i) { g[gidx<D>(i, i)] += lambda; beta[i] = g[NG + i]; }
  bool ok = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const T d = g[gidx<D>(c, c)];
    if (!(d > T(0)) || !isfinite(d)) ok = false;
    const T inv = rsqrt(d);
    g[gidx<D>(c, c)] = inv;
#pragma unroll
    for (int i = c + 1; i < D; ++i) g[gidx<D>(i, c)] *= inv;
#pragma unroll
    for (int j = c + 1; j < D; ++j)
#pragma unroll
      for (int i = j; i < D; ++i) g[gidx<D>(i, j)] -= g[gidx<D>(i, c)] * g[gidx<D>(j, c)];
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    T s = beta[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= g[gidx<D>(i, j)] * beta[j];
    beta[i] = s * g[gidx<D>(i, i)];
  }
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    T s = beta[i];
#pragma unroll
    for (int j = i + 1; j < D; ++j) s -= g[gidx<D>(j, i)] * beta[j];
    beta[i] = s * g[gidx<D>(i, i)];
  }
  return ok;
}

// ---------------- pass C ----------------
template <typename T, int D>
__global__ void __launch_bounds__(CTA_THREADS)
online_main_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p,
                   int64_t window, int64_t min_rows, int skip, T lambda, int64_t row0, int64_t nchains,
                   const double* __restrict__ C /* [NM][nchains] exclusive chain prefixes */,
                   T* __restrict__ coeffs, T* __restrict__ pred, uint8_t* __restrict__ valid) {
  constexpr int NM = MomN<D>::NM, ES = MomN<D>::ES, GS = MomN<D>::GS, TPL = MomN<D>::TPL, NT = MomN<D>::NT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * WARPS + wid;
  if (k >= nchains) return;
  double* ee = reinterpret_cast<double*>(smem_raw + wid * warp_smem_bytes<T, D>());   // entering rows [32][ES]
  double* el = ee + BATCH * ES;                                                          // leaving rows  [32][ES]
  double* cnt = el + BATCH * ES;                                                         // finite-row count per row
  T* gs = reinterpret_cast<T*>(cnt + BATCH);                                             // W_t per row    [32][GS]
  const bool rolling = window > 0;
  const int64_t chain0 = k * CHAIN_ROWS;
  const bool cnt_lane = lane == ((NT - 1) & 31);

  int ta[TPL], tb[TPL], k0[TPL], k1[TPL];
  double W[TPL][2];
#pragma unroll
  for (int m = 0; m < TPL; ++m) {
    task_of<D>(lane + 32 * m, ta[m], tb[m], k0[m], k1[m]);
    W[m][0] = (k0[m] < NM) ? C[(size_t)k0[m] * nchains + k] : 0.0;
    W[m][1] = (k1[m] < NM) ? C[(size_t)k1[m] * nchains + k] : 0.0;
  }
  if (rolling) {
    // rows [lo, chain0) are inside the window of the chain's first row: subtract the prefix up to lo = C[kl] + head
    const int64_t lo = max((int64_t)0, chain0 - window);
    const int64_t kl = lo / CHAIN_ROWS;
    double H[TPL][2];
#pragma unroll
    for (int m = 0; m < TPL; ++m) {
      H[m][0] = (k0[m] < NM) ? C[(size_t)k0[m] * nchains + kl] : 0.0;
      H[m][1] = (k1[m] < NM) ? C[(size_t)k1[m] * nchains + kl] : 0.0;
    }
    for (int64_t rb = kl * CHAIN_ROWS; rb < lo; rb += BATCH) {     // empty when the window is a multiple of the chain
      T z[D]; T yv;
      const int64_t r = rb + lane;
      load_raw<T, D>(X, ldx, y, p, r, n, z, yv);
      put_row<T, D>(ee + lane * ES, z, yv, r < lo);
      __syncwarp();
      walk_batch<T, D, false, false>(ee, ee, ta, tb, k0, k1, H, (T*)nullptr, (double*)nullptr, false);
      __syncwarp();
    }
#pragma unroll
    for (int m = 0; m < TPL; ++m) { W[m][0] -= H[m][0]; W[m][1] -= H[m][1]; }
  }

  // software pipeline: the raw rows of batch b+1 are in flight while batch b is walked and solved
  T zn[D], yn, zln[D], yln;
  load_raw<T, D>(X, ldx, y, p, chain0 + lane, n, zn, yn);
  if (rolling) load_raw<T, D>(X, ldx, y, p, chain0 + lane - window, n, zln, yln);
  for (int b = 0; b < CHAIN_ROWS / BATCH; ++b) {
    const int64_t rb = chain0 + (int64_t)b * BATCH;
    if (rb >= n) break;
    const int64_t r = rb + lane;
    // ---- lane-per-row: stage ----
    T z[D]; T yv;
#pragma unroll
    for (int c = 0; c < D; ++c) z[c] = zn[c];
    yv = yn;
    const bool fin = put_row<T, D>(ee + lane * ES, z, yv, r < n);
    if (rolling) put_row<T, D>(el + lane * ES, zln, yln, r - window >= 0 && r - window < n);
    load_raw<T, D>(X, ldx, y, p, r + BATCH, n, zn, yn);
    if (rolling) load_raw<T, D>(X, ldx, y, p, r + BATCH - window, n, zln, yln);
    __syncwarp();
    // ---- lane-per-component: walk ----
    if (rolling) walk_batch<T, D, true, true>(ee, el, ta, tb, k0, k1, W, gs, cnt, cnt_lane);
    else walk_batch<T, D, false, true>(ee, ee, ta, tb, k0, k1, W, gs, cnt, cnt_lane);
    __syncwarp();
    // ---- lane-per-row: solve ----
    T g[NM];
#pragma unroll
    for (int c = 0; c < NM; ++c) g[c] = gs[lane * GS + c];
    const double cn = cnt[lane];
    __syncwarp();
    bool ok;
    if (rolling) ok = (r >= window - 1) && (!skip || cn >= (double)min_rows - 0.5);
    else ok = skip ? (fin && cn >= (double)min_rows - 0.5) : (r + row0 >= min_rows - 1);
    T beta[D];
    T pr = T(0);
    if (ok) {
      const bool pd = chol_solve_packed<T, D>(g, lambda, beta);
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
    for (int i = 0; i < D; ++i) gs[lane * D + i] = beta[i];
    __syncwarp();
    const int nout = (int)min((int64_t)BATCH, n - rb) * D;
    T* cdst = coeffs + rb * D;
#pragma unroll
    for (int i = 0; i < D; ++i) { const int idx = i * 32 + lane; if (idx < nout) cdst[idx] = gs[idx]; }   // coalesced
    if (r < n) { pred[r] = pr; valid[r] = ok ? 1 : 0; }
    __syncwarp();
  }
}

template <typename T, int D>
int run_online(const T* X, int64_t ldx, const T* y, int64_t n, int p, int64_t window, int64_t min_rows, int skip,
               double lambda, const double* m0, int64_t row0, T* coeffs, T* pred, uint8_t* valid, cudaStream_t s) {
  constexpr int NM = MomN<D>::NM;
  const int64_t nchains = ceil_div(n, CHAIN_ROWS);
  double* S = nullptr;
  if (dev_alloc((void**)&S, (size_t)NM * nchains * sizeof(double), s)) return 1;
  const size_t smem = WARPS * warp_smem_bytes<T, D>();
  auto ka = chain_sums_kernel<T, D>;
  auto kc = online_main_kernel<T, D>;
  if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)ceil_div(nchains, WARPS);
  ka<<<grid, CTA_THREADS, 0, s>>>(X, ldx, y, n, p, nchains, S);
  cudaError_t e = cudaGetLastError(); count_launch();
  if (e == cudaSuccess) { tile_scan_kernel<<<NM, 1024, 0, s>>>(S, nchains, m0, p, D); e = cudaGetLastError(); count_launch(); }
  if (e == cudaSuccess) {
    kc<<<grid, CTA_THREADS, smem, s>>>(X, ldx, y, n, p, window, min_rows, skip, (T)lambda, row0, nchains, S, coeffs, pred, valid);
    e = cudaGetLastError(); count_launch();
  }
  dev_free(S, s);
  if (e != cudaSuccess) { set_error("online lin_reg launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

}  // namespace

template <typename T>
int online_lin_reg(const T* X, int64_t ldx, const T* y, int64_t n, int p, int add_bias, int64_t window,
                   int64_t min_rows, int skip, double lambda, const double* m0, int64_t row0, T* coeffs, T* pred,
                   uint8_t* valid, cudaStream_t s) {
  if (n <= 0) return 0;
  if (window > 0 && m0) { set_error("online lin_reg: preceding-row moments only apply to the recursive fit"); return 1; }
  const int d = p + (add_bias ? 1 : 0);
  if (n / CHAIN_ROWS > 2000000000LL) { set_error("online lin_reg: too many rows"); return 1; }
#define CASE_D(DD) case DD: return run_online<T, DD>(X, ldx, y, n, p, window, min_rows, skip, lambda, m0, row0, coeffs, pred, valid, s);
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
                                   double, const double*, int64_t, float*, float*, uint8_t*, cudaStream_t);
template int online_lin_reg<double>(const double*, int64_t, const double*, int64_t, int, int, int64_t, int64_t, int,
                                    double, const double*, int64_t, double*, double*, uint8_t*, cudaStream_t);

}  // namespace pdsb
