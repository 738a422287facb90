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
#include <cstdlib>

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
        // products no moment wants (slot NM) are not stored: several lanes would write that slot, harmlessly, but
        // compute-sanitizer's racecheck reports every one of them as a WAW hazard
        if (k0[m] < MomN<D>::NM) gs[r * GS + k0[m]] = (T)W[m][0];
        if (k1[m] < MomN<D>::NM) gs[r * GS + k1[m]] = (T)W[m][1];
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
  // Accumulation in the DATA type inside a lane (32 rows of the chain per lane), f64 across lanes: for f32 data the
  // per-lane sums are 32-term f32 FMA chains — the same precision pass C's scan works in, and a window only ever sees the
  // DIFFERENCE of a few chain sums, so the error does not grow with n.  (With f64 products this pass ran at 1.36 ms per
  // 1e8 x 9 f32 = 40 % of the HBM peak: 45 DFMA + 9 conversions per row on a part whose FP64 pipe is narrow.)
  constexpr int NMP = (NM + 31) / 32 * 32;               // padded to whole 32s for the transpose-reduce at the end
  T v[NM];
#pragma unroll
  for (int c = 0; c < NM; ++c) v[c] = T(0);
  const int64_t chain0 = k * CHAIN_ROWS;
  // The pass is a pure stream (read (p+1) s bytes per row, keep NM sums): what bounds it is bytes in flight.  ncu, round
  // 2: with one 32-row batch of loads per warp outstanding the kernel sat at 25 % of the HBM peak (2.1 ms per 1e8 x 9
  // f32).  PF batches are kept in flight per warp (register ring, statically indexed by unrolling the loop PF times).
  constexpr int PF = (D <= 6 || (sizeof(T) == 4 && D <= 9)) ? 4 : 2;
  T zb[PF][D]; T yb[PF];
#pragma unroll
  for (int u = 0; u < PF; ++u) load_raw<T, D>(X, ldx, y, p, chain0 + (int64_t)u * 32 + lane, n, zb[u], yb[u]);
  for (int j0 = 0; j0 < CHAIN_ROWS / 32; j0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int j = j0 + u;
      const int64_t r = chain0 + (int64_t)j * 32 + lane;
      T z[D]; T yv;
#pragma unroll
      for (int c = 0; c < D; ++c) z[c] = zb[u][c];
      yv = yb[u];
      if (j + PF < CHAIN_ROWS / 32) load_raw<T, D>(X, ldx, y, p, r + PF * 32, n, zb[u], yb[u]);
      T acc = yv * T(0);
#pragma unroll
      for (int c = 0; c < D; ++c) acc = fma(z[c], T(0), acc);
      const bool fin = (r < n) && (acc == T(0));
#pragma unroll
      for (int c = 0; c < D; ++c) z[c] = fin ? z[c] : T(0);
      yv = fin ? yv : T(0);
      int q = 0;
#pragma unroll
      for (int i = 0; i < D; ++i)
#pragma unroll
        for (int jj = i; jj < D; ++jj) { v[q] = fma(z[i], z[jj], v[q]); ++q; }
#pragma unroll
      for (int i = 0; i < D; ++i) { v[q] = fma(z[i], yv, v[q]); ++q; }
      v[q] += fin ? T(1) : T(0);
    }
  }
  // Transpose-reduce over the warp, in f64: at every step a lane hands the half of its values it does not keep to its
  // partner and adds the partner's half it keeps -> NMP - NMP / 32 double shuffles and DADDs for all NM sums, where the
  // butterfly per value cost 5 of each (a quarter of this kernel's instructions at NM = 46).  (The same tree in f32 was
  // 2 % faster still but put one rolling parity case outside the two-sided f32 rule of tests/parity_rule.py.)
  double dv[NMP];
#pragma unroll
  for (int c = 0; c < NMP; ++c) dv[c] = c < NM ? (double)v[c] : 0.0;
#pragma unroll
  for (int s = 16, m = NMP / 2; s >= 1; s >>= 1, m >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < m; ++i) {
      const double give = upper ? dv[i] : dv[i + m];
      const double keep = upper ? dv[i + m] : dv[i];
      dv[i] = keep + __shfl_xor_sync(0xffffffffu, give, s);
    }
  }
  const int cbase = ((lane >> 4) & 1) * (NMP / 2) + ((lane >> 3) & 1) * (NMP / 4) + ((lane >> 2) & 1) * (NMP / 8) +
                    ((lane >> 1) & 1) * (NMP / 16) + (lane & 1) * (NMP / 32);
#pragma unroll
  for (int j = 0; j < NMP / 32; ++j)
    if (cbase + j < NM) S[(size_t)(cbase + j) * nchains + k] = dv[j];
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

// =====================================================================================================================
// Pass C, f32, packed: the production kernel of rolling / recursive for float data (D <= 12).
//
// ncu on the kernel above (profiles/README.md, round 1): issue-slot-bound at ~40 warp-instructions per row — the f64
// lane-per-moment walk (4 LDS + 4 DFMA per row), the staging of e-vectors as doubles, and a 32-wide Cholesky.  This
// version does the same mathematics with a quarter of the instructions:
//   * a warp takes 64 rows per step and every lane owns the ADJACENT row pair (2l, 2l+1): all per-row arithmetic is
//     packed f32x2 (FFMA2 / FMUL2 / FADD2, sm_100), so products, window sums, the Cholesky and the prediction of two rows
//     cost one instruction stream;
//   * moments of a step are formed as products  m_k(t) = e_a e_b - l_a l_b  by the row lanes (2 packed instructions per
//     moment per 2 rows), written to shared memory [moment][row], and turned into running sums by lane-per-moment serial
//     scans over the 64 rows (LDS.128 / 4 FADD / STS.128) — in f32, from zero, so the rounding is that of a 64-term sum;
//   * f64 lives only at step boundaries: each moment lane keeps the exact-to-f64 base W_k (chain prefix from passes A/B
//     plus the step totals) and publishes it as f32 once per step;  W_t = base + scan(t)  is formed by the row lanes.
// The Gram the solver sees is an f32 rounding of the window moments, exactly as before (the old kernel cast its f64
// walk to T before the solve); what changes is 12 instead of 40 warp-instructions per row.
// =====================================================================================================================
constexpr int SB = 64;                   // rows per step of the packed kernel
constexpr int SBP = 68;                  // row stride of the product tile: 272 B = 16 (mod 128) -> conflict-free LDS.128 scans
constexpr int V2_WARPS = 4;
constexpr int K6_DEFAULT_VAR = 0;         // see online_main_f32x2_kernel (PDSB_K6_VAR overrides for A/B timing)

template <int D> struct V2 {
  static constexpr int NG = D * (D + 1) / 2;
  static constexpr int NM = NG + D + 1;
  static constexpr int TPL = (NM + 31) / 32;
  static constexpr int NB = (NM + 3) & ~3;
  static constexpr size_t warp_floats = (size_t)NM * SBP + NB;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }

// raw pair load (rows r, r+1; clamped addresses, no branches); z[D] (features, ones for the bias slot) and y
template <int D, bool CLAMP = true>
__device__ __forceinline__ void load_pair(const float* __restrict__ X, int64_t ldx, const float* __restrict__ y, int p,
                                          int64_t r, int64_t n, float2* z, float2& yv) {
  const int64_t r0 = CLAMP ? min(max(r, (int64_t)0), n - 1) : r, r1 = CLAMP ? min(max(r + 1, (int64_t)0), n - 1) : r + 1;
  const float* q = X;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    if (c < p) { z[c] = f2(__ldg(q + r0), __ldg(q + r1)); q += ldx; }
    else z[c] = f2(1.0f, 1.0f);
  }
  yv = f2(__ldg(y + r0), __ldg(y + r1));
}

// e-vector pair: zero where the row is missing or holds a non-finite value; fin = 1 / 0 per row
template <int D>
__device__ __forceinline__ void finish_pair(float2* z, float2& yv, bool in0, bool in1, float2& fin) {
  float2 acc = __fmul2_rn(yv, bc(0.0f));
#pragma unroll
  for (int c = 0; c < D; ++c) acc = __ffma2_rn(z[c], bc(0.0f), acc);      // 0 when every entry is finite, NaN otherwise
  const bool k0 = in0 && (acc.x == 0.0f), k1 = in1 && (acc.y == 0.0f);
  fin = f2(k0 ? 1.0f : 0.0f, k1 ? 1.0f : 0.0f);
  if (__all_sync(0xffffffffu, k0 && k1)) return;             // the usual case: every row of the step exists and is finite
#pragma unroll
  for (int c = 0; c < D; ++c) z[c] = f2(k0 ? z[c].x : 0.0f, k1 ? z[c].y : 0.0f);
  yv = f2(k0 ? yv.x : 0.0f, k1 ? yv.y : 0.0f);
}

// row lanes: products of one step -> tile[k][2 lane .. 2 lane + 1]
template <int D, bool BOTH>
__device__ __forceinline__ void put_products(float* __restrict__ tile, int lane, const float2* e, float2 ey, float2 ef,
                                             const float2* l, float2 ly, float2 lf) {
  constexpr int NG = V2<D>::NG, NM = V2<D>::NM;
  float2* out = reinterpret_cast<float2*>(tile) + lane;
  constexpr int ST = SBP / 2;                                // stride in float2
  int k = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float2 nli = BOTH ? __fmul2_rn(l[i], bc(-1.0f)) : bc(0.0f);
#pragma unroll
    for (int j = i; j < D; ++j) {
      float2 m = BOTH ? __ffma2_rn(e[i], e[j], __fmul2_rn(nli, l[j])) : __fmul2_rn(e[i], e[j]);
      out[(size_t)k * ST] = m;
      ++k;
    }
    float2 my = BOTH ? __ffma2_rn(e[i], ey, __fmul2_rn(nli, ly)) : __fmul2_rn(e[i], ey);
    out[(size_t)(NG + i) * ST] = my;
  }
  out[(size_t)(NM - 1) * ST] = BOTH ? __fadd2_rn(ef, __fmul2_rn(lf, bc(-1.0f))) : ef;
}

// moment lanes: in-place inclusive scan of the 64 rows of their moments; returns the step totals.
// Two levels, so that nothing but a 16-long carry chain is serial: all 16 LDS.128 of a row are issued at once, the four
// values of every float4 are prefix-summed independently (16-way ILP), the carries run over the 16 group totals, and the
// carry is added back with packed adds.  (The first version walked the 64 values in one dependent LDS -> FADD chain:
// ncu attributed a quarter of all stall samples to those FADDs waiting on shared-memory loads.)
template <int D>
__device__ __forceinline__ void scan_products(float* __restrict__ tile, int lane, float* tot) {
  constexpr int NM = V2<D>::NM, TPL = V2<D>::TPL;
#pragma unroll
  for (int m = 0; m < TPL; ++m) {
    tot[m] = 0.0f;
    if (lane + 32 * m < NM) {
      float4* row = reinterpret_cast<float4*>(tile + (size_t)(lane + 32 * m) * SBP);
      float4 v[SB / 4];
#pragma unroll
      for (int i = 0; i < SB / 4; ++i) v[i] = row[i];
#pragma unroll
      for (int i = 0; i < SB / 4; ++i) { v[i].y += v[i].x; v[i].w += v[i].z; v[i].z += v[i].y; v[i].w += v[i].y; }
      float c = 0.0f;
#pragma unroll
      for (int i = 0; i < SB / 4; ++i) {
        const float t = v[i].w;
        const float2 lo = __fadd2_rn(make_float2(v[i].x, v[i].y), make_float2(c, c));
        const float2 hi = __fadd2_rn(make_float2(v[i].z, v[i].w), make_float2(c, c));
        row[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
        c += t;
      }
      tot[m] = c;
    }
  }
}

// packed Cholesky solve of two rows at once (same algorithm as chol_solve_packed)
template <int D>
__device__ __forceinline__ void chol_solve_pair(float2* g, float lambda, float2* beta, bool& ok0, bool& ok1) {
  constexpr int NG = V2<D>::NG;
#pragma unroll
  for (int i = 0; i < D; ++i) { g[gidx<D>(i, i)] = __fadd2_rn(g[gidx<D>(i, i)], bc(lambda)); beta[i] = g[NG + i]; }
  ok0 = ok1 = true;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    const float2 d = g[gidx<D>(c, c)];
    if (!(d.x > 0.0f) || !isfinite(d.x)) ok0 = false;
    if (!(d.y > 0.0f) || !isfinite(d.y)) ok1 = false;
    const float2 inv = f2(rsqrtf(d.x), rsqrtf(d.y));
    g[gidx<D>(c, c)] = inv;
#pragma unroll
    for (int i = c + 1; i < D; ++i) g[gidx<D>(i, c)] = __fmul2_rn(g[gidx<D>(i, c)], inv);
#pragma unroll
    for (int j = c + 1; j < D; ++j) {
      const float2 nj = __fmul2_rn(g[gidx<D>(j, c)], bc(-1.0f));
#pragma unroll
      for (int i = j; i < D; ++i) g[gidx<D>(i, j)] = __ffma2_rn(g[gidx<D>(i, c)], nj, g[gidx<D>(i, j)]);
    }
  }
  float2 nb[D];                       // -beta: the packed FMA has no negate modifier, so the sign is carried by the vector
#pragma unroll
  for (int i = 0; i < D; ++i) {
    float2 s = beta[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s = __ffma2_rn(g[gidx<D>(i, j)], nb[j], s);
    beta[i] = __fmul2_rn(s, g[gidx<D>(i, i)]);
    nb[i] = __fmul2_rn(beta[i], bc(-1.0f));
  }
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    float2 s = beta[i];
#pragma unroll
    for (int j = i + 1; j < D; ++j) s = __ffma2_rn(g[gidx<D>(j, i)], nb[j], s);
    beta[i] = __fmul2_rn(s, g[gidx<D>(i, i)]);
    nb[i] = __fmul2_rn(beta[i], bc(-1.0f));
  }
}

// VAR: 0 = register prefetch of the next step's rows (4-warp CTAs, 2 CTAs per SM at D = 8);
//      1 / 2 = no register prefetch and a register cap for 3 / 4 CTAs per SM (more warps hide the load latency instead)
template <int D, bool ROLLING, int VAR>
__global__ void __launch_bounds__(V2_WARPS * 32, VAR == 0 ? 1 : (VAR == 1 ? 3 : 4))
online_main_f32x2_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ y, int64_t n, int p,
                         int64_t window, int64_t min_rows, int skip, float lambda, int64_t row0, int64_t nchains,
                         const double* __restrict__ C /* [NM][nchains] exclusive chain prefixes */,
                         float* __restrict__ coeffs, float* __restrict__ pred, uint8_t* __restrict__ valid) {
  constexpr int NM = V2<D>::NM, TPL = V2<D>::TPL, NB = V2<D>::NB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t k = (int64_t)blockIdx.x * V2_WARPS + wid;
  if (k >= nchains) return;
  float* tile = reinterpret_cast<float*>(smem_raw) + (size_t)wid * V2<D>::warp_floats;     // [NM][SBP]
  float* basef = tile + (size_t)NM * SBP;                                                   // [NB]
  const int64_t chain0 = k * CHAIN_ROWS;

  // moment lanes: f64 base of "their" moments at the chain start
  double W[TPL];
#pragma unroll
  for (int m = 0; m < TPL; ++m) W[m] = (lane + 32 * m < NM) ? C[(size_t)(lane + 32 * m) * nchains + k] : 0.0;
  if (ROLLING) {
    // rows [lo, chain0) are inside the window of the chain's first row: subtract the prefix up to lo = C[kl] + head
    const int64_t lo = max((int64_t)0, chain0 - window);
    const int64_t kl = lo / CHAIN_ROWS;
#pragma unroll
    for (int m = 0; m < TPL; ++m) if (lane + 32 * m < NM) W[m] -= C[(size_t)(lane + 32 * m) * nchains + kl];
    for (int64_t rb = kl * CHAIN_ROWS; rb < lo; rb += SB) {       // empty when the window is a multiple of the chain
      float2 z[D], yv, fin;
      const int64_t r = rb + 2 * lane;
      load_pair<D>(X, ldx, y, p, r, n, z, yv);
      finish_pair<D>(z, yv, r < lo, r + 1 < lo, fin);
      put_products<D, false>(tile, lane, z, yv, fin, z, yv, fin);
      __syncwarp();
      float tot[TPL];
      scan_products<D>(tile, lane, tot);
#pragma unroll
      for (int m = 0; m < TPL; ++m) W[m] -= (double)tot[m];
      __syncwarp();
    }
  }

  // interior chains (all but the first and the last few) need neither clamped addresses nor row-exists tests
  const bool interior = (chain0 + CHAIN_ROWS + SB <= n) && (!ROLLING || chain0 - window >= 0);
  constexpr bool PREF = (VAR == 0);
  // VAR 0: software pipeline — the raw rows of step b+1 are in flight while step b is scanned and solved
  float2 zn[PREF ? D : 1], yn, zln[PREF ? D : 1], yln;
  if (PREF) {
    load_pair<D>(X, ldx, y, p, chain0 + 2 * lane, n, zn, yn);
    if (ROLLING) load_pair<D>(X, ldx, y, p, chain0 + 2 * lane - window, n, zln, yln);
  }
  for (int b = 0; b < CHAIN_ROWS / SB; ++b) {
    const int64_t rb = chain0 + (int64_t)b * SB;
    if (rb >= n) break;
    const int64_t r = rb + 2 * lane;
    // ---- row lanes: products of the step ----
    float2 z[D], yv, fin, zl[D], yl;
    if (PREF) {
#pragma unroll
      for (int c = 0; c < D; ++c) { z[c] = zn[c]; if (ROLLING) zl[c] = zln[c]; }
      yv = yn; yl = yln;
    } else if (interior) {
      load_pair<D, false>(X, ldx, y, p, r, n, z, yv);
      if (ROLLING) load_pair<D, false>(X, ldx, y, p, r - window, n, zl, yl);
    } else {
      load_pair<D>(X, ldx, y, p, r, n, z, yv);
      if (ROLLING) load_pair<D>(X, ldx, y, p, r - window, n, zl, yl);
    }
    finish_pair<D>(z, yv, interior || r < n, interior || r + 1 < n, fin);
    if (ROLLING) {
      float2 lf;
      finish_pair<D>(zl, yl, interior || (r - window >= 0 && r - window < n), interior || (r + 1 - window >= 0 && r + 1 - window < n), lf);
      put_products<D, true>(tile, lane, z, yv, fin, zl, yl, lf);
    } else {
      put_products<D, false>(tile, lane, z, yv, fin, z, yv, fin);
    }
    if (PREF) {
      if (interior) {
        load_pair<D, false>(X, ldx, y, p, r + SB, n, zn, yn);
        if (ROLLING) load_pair<D, false>(X, ldx, y, p, r + SB - window, n, zln, yln);
      } else {
        load_pair<D>(X, ldx, y, p, r + SB, n, zn, yn);
        if (ROLLING) load_pair<D>(X, ldx, y, p, r + SB - window, n, zln, yln);
      }
    }
    __syncwarp();
    // ---- moment lanes: publish the base, scan, advance the base ----
    {
      float tot[TPL];
#pragma unroll
      for (int m = 0; m < TPL; ++m) if (lane + 32 * m < NM) basef[lane + 32 * m] = (float)W[m];
      scan_products<D>(tile, lane, tot);
#pragma unroll
      for (int m = 0; m < TPL; ++m) W[m] += (double)tot[m];
    }
    __syncwarp();
    // ---- row lanes: W_t = base + scan(t), solve both rows, predict ----
    float2 g[NM];
    {
      const float2* src = reinterpret_cast<const float2*>(tile) + lane;
#pragma unroll
      for (int c4 = 0; c4 < NB / 4; ++c4) {
        const float4 b4 = *reinterpret_cast<const float4*>(basef + 4 * c4);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = 4 * c4 + u;
          if (c < NM) g[c] = __fadd2_rn(src[(size_t)c * (SBP / 2)], bc(bb[u]));
        }
      }
    }
    __syncwarp();
    const float2 cn = g[NM - 1];
    bool ok0, ok1;
    if (ROLLING) {
      ok0 = (r >= window - 1) && (!skip || cn.x >= (float)min_rows - 0.5f);
      ok1 = (r + 1 >= window - 1) && (!skip || cn.y >= (float)min_rows - 0.5f);
    } else {
      ok0 = skip ? (fin.x != 0.0f && cn.x >= (float)min_rows - 0.5f) : (r + row0 >= min_rows - 1);
      ok1 = skip ? (fin.y != 0.0f && cn.y >= (float)min_rows - 0.5f) : (r + 1 + row0 >= min_rows - 1);
    }
    float2 beta[D];
    bool pd0, pd1;
    chol_solve_pair<D>(g, lambda, beta, pd0, pd1);
    float2 pr = bc(0.0f);
#pragma unroll
    for (int i = 0; i < D; ++i) {
      beta[i] = f2(ok0 ? (pd0 ? beta[i].x : nanf("")) : 0.0f, ok1 ? (pd1 ? beta[i].y : nanf("")) : 0.0f);
      pr = __ffma2_rn(z[i], beta[i], pr);
    }
    // rows 2l, 2l+1 are adjacent: 2 D consecutive floats per lane, consecutive lanes consecutive -> coalesced
    if (r + 1 < n) {
      float* cdst = coeffs + r * D;
      if ((D & 1) == 0) {
        float2* c2 = reinterpret_cast<float2*>(cdst);
#pragma unroll
        for (int i = 0; i < D / 2; ++i) c2[i] = f2(beta[2 * i].x, beta[2 * i + 1].x);
#pragma unroll
        for (int i = 0; i < D / 2; ++i) c2[D / 2 + i] = f2(beta[2 * i].y, beta[2 * i + 1].y);
      } else {
#pragma unroll
        for (int i = 0; i < D; ++i) cdst[i] = beta[i].x;
#pragma unroll
        for (int i = 0; i < D; ++i) cdst[D + i] = beta[i].y;
      }
      pred[r] = pr.x; pred[r + 1] = pr.y;
      valid[r] = ok0 ? 1 : 0; valid[r + 1] = ok1 ? 1 : 0;
    } else if (r < n) {
      float* cdst = coeffs + r * D;
#pragma unroll
      for (int i = 0; i < D; ++i) cdst[i] = beta[i].x;
      pred[r] = pr.x;
      valid[r] = ok0 ? 1 : 0;
    }
  }
}

// =====================================================================================================================
// Generic pass C for any number of coefficients (D > 12): the reference takes any p (lr_online_solvers.rs:148-301).
// One warp per chain, rows in order; the window moments live in shared memory as f64 ((D+2)(D+3)/2 packed upper
// triangle over (z, y, 1)), lanes stride over the moments for the rank-1 updates and cooperate on an f64 Cholesky of a
// scratch copy per row.  O(D^3 / 32) per row — a correctness path, not a roofline one.
// =====================================================================================================================
template <typename T>
__global__ void __launch_bounds__(32)
online_generic_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p, int d,
                      int64_t window, int64_t min_rows, int skip, double lambda, int64_t row0, int64_t nchains,
                      const double* __restrict__ C /* [NM][nchains] */, T* __restrict__ coeffs, T* __restrict__ pred,
                      uint8_t* __restrict__ valid) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int ng = d * (d + 1) / 2, nm = ng + d + 1;
  double* W = reinterpret_cast<double*>(smem_raw);      // [nm] window moments, component order of the chain sums
  double* A = W + nm;                                   // [d*d] scratch (lower triangle used)
  double* e = A + d * d;                                // [d+2] entering row
  double* l = e + d + 2;                                // [d+2] leaving row
  double* bt = l + d + 2;                               // [d]
  const bool rolling = window > 0;
  const int64_t chain0 = k * CHAIN_ROWS;
  auto load_row = [&](int64_t r, double* dst, bool use) {
    bool fin = use && r >= 0 && r < n;
    double acc = 0.0;
    for (int c = lane; c <= d; c += 32) {
      double v = 0.0;
      if (fin) v = (c < p) ? (double)X[(size_t)c * ldx + r] : (c < d ? 1.0 : (double)y[r]);
      dst[c] = v;
      acc += v * 0.0;
    }
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    fin = fin && (acc == 0.0);
    __syncwarp();
    if (!fin) for (int c = lane; c <= d; c += 32) dst[c] = 0.0;
    if (lane == 0) dst[d + 1] = fin ? 1.0 : 0.0;
    __syncwarp();
    return fin;
  };
  auto rank1 = [&](const double* v, double sgn) {       // W += sgn * m(v)
    for (int c = lane; c < nm; c += 32) {
      int i, j;
      if (c >= ng + d) { i = j = d + 1; }
      else if (c >= ng) { i = c - ng; j = d; }
      else { int cc = c; i = 0; while (cc >= d - i) { cc -= d - i; ++i; } j = i + cc; }
      W[c] += sgn * v[i] * v[j];
    }
    __syncwarp();
  };
  for (int c = lane; c < nm; c += 32) W[c] = C[(size_t)c * nchains + k];
  __syncwarp();
  if (rolling) {
    const int64_t lo = max((int64_t)0, chain0 - window);
    const int64_t kl = lo / CHAIN_ROWS;
    for (int c = lane; c < nm; c += 32) W[c] -= C[(size_t)c * nchains + kl];
    __syncwarp();
    for (int64_t r = kl * CHAIN_ROWS; r < lo; ++r) { load_row(r, e, true); rank1(e, -1.0); }
  }
  for (int64_t r = chain0; r < min(chain0 + (int64_t)CHAIN_ROWS, n); ++r) {
    const bool fin = load_row(r, e, true);
    rank1(e, 1.0);
    if (rolling) { load_row(r - window, l, true); rank1(l, -1.0); }
    const double cn = W[nm - 1];
    bool ok;
    if (rolling) ok = (r >= window - 1) && (!skip || cn >= (double)min_rows - 0.5);
    else ok = skip ? (fin && cn >= (double)min_rows - 0.5) : (r + row0 >= min_rows - 1);
    bool pd = true;
    if (ok) {
      // A = lower triangle of G + lambda I, rounded through T like the packed kernels; bt = X'y
      for (int c = lane; c < ng; c += 32) {
        int cc = c, i = 0; while (cc >= d - i) { cc -= d - i; ++i; } const int j = i + cc;
        A[j * d + i] = (double)(T)W[c] + (i == j ? lambda : 0.0);
      }
      for (int c = lane; c < d; c += 32) bt[c] = (double)(T)W[ng + c];
      __syncwarp();
      for (int c = 0; c < d; ++c) {
        const double dg = A[c * d + c];
        if (!(dg > 0.0) || !isfinite(dg)) pd = false;
        const double inv = rsqrt(dg);
        __syncwarp();
        for (int i = c + 1 + lane; i < d; i += 32) A[i * d + c] *= inv;
        if (lane == 0) A[c * d + c] = inv;
        __syncwarp();
        for (int idx = lane; idx < (d - c - 1) * (d - c - 1); idx += 32) {
          const int j = c + 1 + idx / (d - c - 1), i = c + 1 + idx % (d - c - 1);
          if (i >= j) A[i * d + j] -= A[i * d + c] * A[j * d + c];
        }
        __syncwarp();
      }
      for (int i = 0; i < d; ++i) {            // forward substitution (lane-parallel dot)
        double s = 0.0;
        for (int j = lane; j < i; j += 32) s += A[i * d + j] * bt[j];
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) bt[i] = (bt[i] - s) * A[i * d + i];
        __syncwarp();
      }
      for (int i = d - 1; i >= 0; --i) {       // backward
        double s = 0.0;
        for (int j = i + 1 + lane; j < d; j += 32) s += A[j * d + i] * bt[j];
        for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) bt[i] = (bt[i] - s) * A[i * d + i];
        __syncwarp();
      }
    }
    double prs = 0.0;
    for (int c = lane; c < d; c += 32) {
      const double bv = ok ? (pd ? bt[c] : nan("")) : 0.0;
      coeffs[r * d + c] = (T)bv;
      const double zc = (c < p) ? (double)X[(size_t)c * ldx + r] : 1.0;
      prs += ok ? zc * (double)(T)bv : 0.0;
    }
    for (int off = 16; off; off >>= 1) prs += __shfl_xor_sync(0xffffffffu, prs, off);
    if (lane == 0) { pred[r] = (T)prs; valid[r] = ok ? 1 : 0; }
    __syncwarp();
  }
}

// chain sums for any d (generic path): one warp per chain, lanes stride over the moments
template <typename T>
__global__ void __launch_bounds__(32)
chain_sums_generic_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, int64_t n, int p, int d,
                          int64_t nchains, double* __restrict__ S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int ng = d * (d + 1) / 2, nm = ng + d + 1;
  double* acc = reinterpret_cast<double*>(smem_raw);   // [nm]
  double* e = acc + nm;                                // [d+2]
  for (int c = lane; c < nm; c += 32) acc[c] = 0.0;
  __syncwarp();
  for (int64_t r = k * CHAIN_ROWS; r < min((k + 1) * (int64_t)CHAIN_ROWS, n); ++r) {
    double z0 = 0.0;
    for (int c = lane; c <= d; c += 32) {
      const double v = (c < p) ? (double)X[(size_t)c * ldx + r] : (c < d ? 1.0 : (double)y[r]);
      e[c] = v; z0 += v * 0.0;
    }
    for (int off = 16; off; off >>= 1) z0 += __shfl_xor_sync(0xffffffffu, z0, off);
    __syncwarp();
    if (z0 == 0.0) {
      if (lane == 0) e[d + 1] = 1.0;
      __syncwarp();
      for (int c = lane; c < nm; c += 32) {
        int i, j;
        if (c >= ng + d) { i = j = d + 1; }
        else if (c >= ng) { i = c - ng; j = d; }
        else { int cc = c; i = 0; while (cc >= d - i) { cc -= d - i; ++i; } j = i + cc; }
        acc[c] += e[i] * e[j];
      }
    }
    __syncwarp();
  }
  for (int c = lane; c < nm; c += 32) S[(size_t)c * nchains + k] = acc[c];
}

template <typename T>
int run_online_generic(const T* X, int64_t ldx, const T* y, int64_t n, int p, int d, int64_t window, int64_t min_rows,
                       int skip, double lambda, const double* m0, int64_t row0, T* coeffs, T* pred, uint8_t* valid,
                       cudaStream_t s) {
  const int nm = d * (d + 1) / 2 + d + 1;
  const int64_t nchains = ceil_div(n, CHAIN_ROWS);
  double* S = nullptr;
  if (dev_alloc((void**)&S, (size_t)nm * nchains * sizeof(double), s)) return 1;
  const size_t smem_a = (size_t)(nm + d + 2) * sizeof(double);
  const size_t smem_c = (size_t)(nm + d * d + 2 * (d + 2) + d) * sizeof(double);
  if (smem_c > 200 * 1024) { dev_free(S, s); set_error("rolling/recursive lin_reg: %d coefficients exceed the device limit (about 150)", d); return 1; }
  auto ka = chain_sums_generic_kernel<T>;
  auto kc = online_generic_kernel<T>;
  if (smem_a > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(ka, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
  if (smem_c > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
  ka<<<(unsigned)nchains, 32, smem_a, s>>>(X, ldx, y, n, p, d, nchains, S);
  cudaError_t e = cudaGetLastError(); count_launch();
  if (e == cudaSuccess) { tile_scan_kernel<<<nm, 1024, 0, s>>>(S, nchains, m0, p, d); e = cudaGetLastError(); count_launch(); }
  if (e == cudaSuccess) {
    kc<<<(unsigned)nchains, 32, smem_c, s>>>(X, ldx, y, n, p, d, window, min_rows, skip, lambda, row0, nchains, S, coeffs, pred, valid);
    e = cudaGetLastError(); count_launch();
  }
  dev_free(S, s);
  if (e != cudaSuccess) { set_error("online lin_reg (generic) launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

// packed f32 pass C (only meaningful for T = float)
template <int D, bool ROLLING, int VAR>
int launch_main_f32x2_v(const float* X, int64_t ldx, const float* y, int64_t n, int p, int64_t window, int64_t min_rows,
                        int skip, double lambda, int64_t row0, int64_t nchains, const double* S, float* coeffs, float* pred,
                        uint8_t* valid, cudaStream_t s) {
  const size_t smem = V2_WARPS * V2<D>::warp_floats * sizeof(float);
  const unsigned grid = (unsigned)ceil_div(nchains, V2_WARPS);
  auto kc = online_main_f32x2_kernel<D, ROLLING, VAR>;
  if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kc<<<grid, V2_WARPS * 32, smem, s>>>(X, ldx, y, n, p, window, min_rows, skip, (float)lambda, row0, nchains, S, coeffs, pred, valid);
  return 0;
}
inline int k6_variant() {
  static int v = [] { const char* e = getenv("PDSB_K6_VAR"); int x = e ? atoi(e) : K6_DEFAULT_VAR; return (x >= 0 && x <= 2) ? x : K6_DEFAULT_VAR; }();
  return v;
}
template <int D>
int launch_main_f32x2(const float* X, int64_t ldx, const float* y, int64_t n, int p, int64_t window, int64_t min_rows,
                      int skip, double lambda, int64_t row0, int64_t nchains, const double* S, float* coeffs, float* pred,
                      uint8_t* valid, cudaStream_t s) {
#define PDSB_K6_GO(R, V) return launch_main_f32x2_v<D, R, V>(X, ldx, y, n, p, window, min_rows, skip, lambda, row0, nchains, S, coeffs, pred, valid, s)
  const int var = k6_variant();
  if (window > 0) { if (var == 1) PDSB_K6_GO(true, 1); if (var == 2) PDSB_K6_GO(true, 2); PDSB_K6_GO(true, 0); }
  if (var == 1) PDSB_K6_GO(false, 1);
  if (var == 2) PDSB_K6_GO(false, 2);
  PDSB_K6_GO(false, 0);
#undef PDSB_K6_GO
}
template <typename T, int D> struct MainF32x2 {
  static bool use() { return false; }
  static int run(const T*, int64_t, const T*, int64_t, int, int64_t, int64_t, int, double, int64_t, int64_t, const double*, T*, T*, uint8_t*, cudaStream_t) { return 1; }
};
template <int D> struct MainF32x2<float, D> {
  // D <= 9 keeps the packed kernel free of register spills (201 registers at D = 8, 236 at D = 9); wider fits use the
  // lane-per-moment kernel above.  PDSB_ONLINE_V1=1 forces the old kernel (A/B timing).
  static bool use() { static int v = [] { const char* e = getenv("PDSB_ONLINE_V1"); return e && atoi(e) ? 0 : 1; }(); return v != 0 && D <= 9; }
  static int run(const float* X, int64_t ldx, const float* y, int64_t n, int p, int64_t window, int64_t min_rows, int skip,
                 double lambda, int64_t row0, int64_t nchains, const double* S, float* coeffs, float* pred, uint8_t* valid,
                 cudaStream_t s) {
    return launch_main_f32x2<D>(X, ldx, y, n, p, window, min_rows, skip, lambda, row0, nchains, S, coeffs, pred, valid, s);
  }
};

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
    // The packed kernel sums a step's products in f32 from zero: for a window shorter than the 64-row step the partial sums
    // (a random walk over 64 rows) are larger than the window sum they end in and the rounding shows (3e-6 relative at
    // window 2); such windows keep the f64 lane-per-moment walk.  Recursive fits and windows >= 64 take the packed kernel.
    if (MainF32x2<T, D>::use() && (window == 0 || window >= SB)) {
      if (MainF32x2<T, D>::run(X, ldx, y, n, p, window, min_rows, skip, lambda, row0, nchains, S, coeffs, pred, valid, s)) { dev_free(S, s); return 1; }
    } else {
      kc<<<grid, CTA_THREADS, smem, s>>>(X, ldx, y, n, p, window, min_rows, skip, (T)lambda, row0, nchains, S, coeffs, pred, valid);
    }
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
    default:   // any number of coefficients, like the reference (lr_online_solvers.rs:148-301): generic shared-memory path
      return run_online_generic<T>(X, ldx, y, n, p, d, window, min_rows, skip, lambda, m0, row0, coeffs, pred, valid, s);
  }
#undef CASE_D
}

template int online_lin_reg<float>(const float*, int64_t, const float*, int64_t, int, int, int64_t, int64_t, int,
                                   double, const double*, int64_t, float*, float*, uint8_t*, cudaStream_t);
template int online_lin_reg<double>(const double*, int64_t, const double*, int64_t, int, int, int64_t, int64_t, int,
                                    double, const double*, int64_t, double*, double*, uint8_t*, cudaStream_t);

}  // namespace pdsb
