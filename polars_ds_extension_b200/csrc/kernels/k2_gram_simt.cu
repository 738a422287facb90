// K2a — generic SIMT moments kernel:  M = [X|Y|1]^T diag(w) [X|Y|1]   (q1 x q1, f64, row-major)
//
// Replaces the reference's get_xtx_with_lambda + build_xty + column sums
// (/root/reference/src/linear/lr/lr_solvers.rs:183-211, 262-278, 483-484) and x^T w x
// (src/num_ext/linear_regression.rs:1026-1027) with ONE pass over the data.
//
// This is the path for f64 data, weighted fits, tiny inputs and shapes the tcgen05 kernel does not take
// (k2_gram_tcgen05.cu is the f32 headline path).  Layout: X col-major [n x p] (ldx), Y col-major [n x t] (ldy).
// Each CTA walks row tiles of TILE_R rows: tile -> shared memory (row-major, row stride S), every thread owns
// up to MAXT 4x4 blocks of the upper triangle; per tile the block is accumulated in T (a short FMA chain)
// and then added to f64 accumulators, so f32 rounding never grows with n.  When there are fewer blocks than threads
// (q1 <= 88, the usual case) the 256 threads are dealt as (block, row slice): 256 / #blocks slices each take every
// nslices-th row of the tile, so all threads work (with one thread per block only 45 of 256 had work at q1 = 34).
// Per-(CTA, slice) partials are reduced in a fixed order by a second kernel -> bit-reproducible results.
#include "../common.h"
#include "kernels.h"

namespace pdsb {

template <typename T>
struct Vec4 { T v[4]; };

template <typename T, int MAXT, bool WEIGHTED>
__global__ void __launch_bounds__(256)
gram_simt_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ Y, int64_t ldy,
                 const T* __restrict__ w, const T* __restrict__ mask, int64_t n, int p, int t,
                 int tile_r, int S, double* __restrict__ partials, int64_t bstride /* 0: column-major; else frame block stride */) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Zs = reinterpret_cast<T*>(smem_raw);            // [tile_r][S]
  T* ws = Zs + (size_t)tile_r * S;                   // [tile_r]
  const int q1 = p + t + 1;
  const int nt = (q1 + 3) / 4;
  const int ntp = nt * (nt + 1) / 2;
  const int tid = threadIdx.x;

  // (block, row slice) decode for this thread; more than 256 blocks -> MAXT blocks per thread, one slice
  const int nslices = (MAXT == 1 && ntp <= 256) ? 256 / ntp : 1;
  const int slice = (MAXT == 1) ? tid / ntp : 0;
  int ti[MAXT], tj[MAXT];
  bool act[MAXT];
#pragma unroll
  for (int m = 0; m < MAXT; ++m) {
    int idx = (MAXT == 1) ? tid % ntp : tid + m * 256;
    act[m] = (MAXT == 1) ? (slice < nslices) : (idx < ntp);
    int a = 0, rem = act[m] ? idx : 0;
    // row a of the upper triangle has (nt - a) entries
    while (rem >= nt - a) { rem -= nt - a; ++a; }
    ti[m] = a; tj[m] = a + rem;
  }
  double acc[MAXT][16];
#pragma unroll
  for (int m = 0; m < MAXT; ++m)
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[m][k] = 0.0;

  // zero the padding columns once
  for (int i = tid; i < tile_r * S; i += 256) Zs[i] = T(0);
  __syncthreads();

  const int64_t ntiles = (n + tile_r - 1) / tile_r;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * tile_r;
    // ---- load tile (coalesced along rows).  LU independent loads are issued before the first shared-memory store:
    // with one load in flight per thread this phase was pure latency (~10 us per 128-row tile) ----
    constexpr int LU = 8;
    const int total = tile_r * q1;
    const int tshift = 31 - __clz(tile_r);          // tile_r is a power of two
    for (int base = tid; base < total; base += 256 * LU) {
      T v[LU];
#pragma unroll
      for (int u = 0; u < LU; ++u) {
        const int idx = base + u * 256;
        v[u] = T(0);
        if (idx < total) {
          const int c = idx >> tshift, r = idx & (tile_r - 1);
          const int64_t row = row0 + r;
          if (row < n) {
            if (bstride) {   // row-blocked frame: [block][column][FRAME_ROWS]
              const int64_t o = (row >> 7) * bstride + (row & 127);
              if (c < p) v[u] = X[o + ((int64_t)c << 7)];
              else if (c < p + t) v[u] = Y[o + ((int64_t)(c - p) << 7)];
              else v[u] = mask ? mask[row] : T(1);
            } else if (c < p) v[u] = X[(int64_t)c * ldx + row];
            else if (c < p + t) v[u] = Y[(int64_t)(c - p) * ldy + row];
            else v[u] = mask ? mask[row] : T(1);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < LU; ++u) {
        const int idx = base + u * 256;
        if (idx < total) Zs[(idx & (tile_r - 1)) * S + (idx >> tshift)] = v[u];
      }
    }
    if (WEIGHTED) {
      for (int r = tid; r < tile_r; r += 256) ws[r] = (row0 + r < n) ? w[row0 + r] : T(0);
    }
    __syncthreads();
    // ---- accumulate ----
#pragma unroll
    for (int m = 0; m < MAXT; ++m) {
      if (!act[m]) continue;
      T loc[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) loc[k] = T(0);
      const T* pa = Zs + 4 * ti[m];
      const T* pb = Zs + 4 * tj[m];
      for (int r = slice; r < tile_r; r += nslices) {
        Vec4<T> a = *reinterpret_cast<const Vec4<T>*>(pa + r * S);
        Vec4<T> b = *reinterpret_cast<const Vec4<T>*>(pb + r * S);
        if (WEIGHTED) {
          T wr = ws[r];
#pragma unroll
          for (int k = 0; k < 4; ++k) a.v[k] *= wr;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) loc[i * 4 + j] = fma(a.v[i], b.v[j], loc[i * 4 + j]);
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[m][k] += (double)loc[k];
    }
    __syncthreads();
  }
  // ---- write this CTA's partial (full symmetric) ----
  double* out = partials + ((size_t)blockIdx.x * nslices + slice) * q1 * q1;
#pragma unroll
  for (int m = 0; m < MAXT; ++m) {
    if (!act[m]) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int gi = 4 * ti[m] + i, gj = 4 * tj[m] + j;
        if (gi < q1 && gj < q1) {
          double v = acc[m][i * 4 + j];
          if (ti[m] == tj[m]) {
            // diagonal block: both (i,j) and (j,i) were accumulated by this thread; with weights
            // a_i*w*b_j vs a_j*w*b_i round identically (commutative), so keep it symmetric anyway
            if (gi <= gj) { out[gi * q1 + gj] = v; out[gj * q1 + gi] = v; }
          } else {
            out[gi * q1 + gj] = v;
            out[gj * q1 + gi] = v;
          }
        }
      }
  }
}

// one CTA per moment entry: thread k sums parts k, k+128, ... in order, then a fixed-shape tree -> reproducible
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const double* __restrict__ partials, int nparts, int len, double* __restrict__ out) {
  __shared__ double sh[128];
  const int i = blockIdx.x;
  double s = 0.0;
  for (int k = threadIdx.x; k < nparts; k += 128) s += partials[(size_t)k * len + i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 64; off; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[i] = sh[0];
}

// ------------------------------------------------------------------------------------------------------------
// f64 data: the same moments on the FP64 tensor-core path (mma.sync m8n8k4, SASS DMMA).  Z~ = [X | Y | 1] is cut into
// NB blocks of 8 columns.  For a 4-row step, lane l holds ONE element per block: Z~[row k0 + l%4][col 8b + l/4] — which
// is at the same time the A fragment (8 x 4, "row") of block b and the B fragment (4 x 8, "col") of block b, so a step
// costs NB loads and NB(NB+1)/2 DMMAs per warp, no shared memory and 8x fewer issue slots than the DFMA kernel above.
// Weights scale the A side only.  Accumulation is f64 throughout; per-CTA partials are combined warp by warp in a
// fixed order, then across CTAs by reduce_partials_kernel: reproducible.
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NB>
__global__ void __launch_bounds__(256)
gram_dmma_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy,
                 const double* __restrict__ w, const double* __restrict__ mask, int64_t n, int p, int t,
                 double* __restrict__ partials /* [grid][q1*q1] */) {
  constexpr int NP = NB * (NB + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);      // [NP][64] CTA-level sum of the warps' accumulators
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, k = lane & 3;
  const int q1 = p + t + 1;
  const double* colp[NB];
  int kind[NB];                                           // 0 data, 1 ones / mask, 2 zero padding
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int c = 8 * b + g;
    kind[b] = c < p + t ? 0 : (c == p + t ? 1 : 2);
    colp[b] = c < p ? X + (int64_t)c * ldx : (c < p + t ? Y + (int64_t)(c - p) * ldy : X);
  }
  double acc[NP][2];
#pragma unroll
  for (int i = 0; i < NP; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }

  const int64_t stride = (int64_t)gridDim.x * 8 * 4;
  auto load_step = [&](int64_t r0, double* z, double& wv) {
    const int64_t r = r0 + k;
    const bool in = r < n;
    wv = (in && w) ? w[r] : 1.0;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      double v = 0.0;
      if (in) {
        if (kind[b] == 0) v = colp[b][r];
        else if (kind[b] == 1) v = mask ? mask[r] : 1.0;
      }
      z[b] = v;
    }
  };
  int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * 4;
  double z[NB], zn[NB], wv = 1.0, wn = 1.0;
  if (r0 < n) load_step(r0, z, wv);
  for (; r0 < n; r0 += stride) {
    const bool more = r0 + stride < n;
    if (more) load_step(r0 + stride, zn, wn);             // next step's loads fly while this step's DMMAs issue
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const double a = z[i] * wv;
#pragma unroll
      for (int j = i; j < NB; ++j) { dmma884(acc[idx][0], acc[idx][1], a, z[j]); ++idx; }
    }
    if (more) {
#pragma unroll
      for (int b = 0; b < NB; ++b) z[b] = zn[b];
      wv = wn;
    }
  }
  // ---- CTA reduction in a fixed order: warp 0 stores, warps 1..7 add one after the other ----
  for (int wturn = 0; wturn < 8; ++wturn) {
    if (warp == wturn) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double* d = sm + i * 64 + g * 8 + 2 * k;
        if (wturn == 0) { d[0] = acc[i][0]; d[1] = acc[i][1]; }
        else { d[0] += acc[i][0]; d[1] += acc[i][1]; }
      }
    }
    __syncthreads();
  }
  // ---- this CTA's partial, full symmetric q1 x q1 (upper entries of the diagonal blocks are mirrored) ----
  double* out = partials + (size_t)blockIdx.x * q1 * q1;
  for (int e = threadIdx.x; e < NP * 64; e += 256) {
    const int blk = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
    int i = 0, rem = blk;
    while (rem >= NB - i) { rem -= NB - i; ++i; }
    const int j = i + rem;
    const int a = 8 * i + rr, b = 8 * j + cc;
    if (a < q1 && b < q1 && (i != j || a <= b)) {
      const double v = sm[e];
      out[(size_t)a * q1 + b] = v;
      out[(size_t)b * q1 + a] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Wide variant: the direct kernel keeps 4 rows x (p + t) x 8 B in flight per warp and measured 33 % of the HBM peak with
// the DMMA pipe at 43 % of ITS measured peak (profiles/fp64_peak.cu: 37 TFLOP/s for every mma.sync f64 shape, 33 for
// DFMA) — it waits on memory latency, not on issue slots.  Here a lane loads TWO consecutive rows (16 bytes) per column
// block: lane (g, k) reads rows r0 + 2k, r0 + 2k + 1 of column 8b + g, an 8-row batch per warp.  The rows of a DMMA's
// k index can be any four rows, so the .x halves (rows r0 + {0, 2, 4, 6}) make one m8n8k4 step and the .y halves the
// next.  DEPTH batches live in a register ring (the loop is unrolled over the ring, so every index is static): DEPTH - 1
// batches = 8 (DEPTH - 1) rows per warp are in flight while one is multiplied.  The mask is just the "ones" column read
// from memory.  128-thread CTAs, MINB per SM.  Needs 16-byte aligned columns and even leading dimensions.
template <int NB, int DEPTH, int MINB, bool WEIGHTED>
__global__ void __launch_bounds__(128, MINB)
gram_dmma_wide_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy,
                      const double* __restrict__ w, const double* __restrict__ mask, int64_t n, int p, int t,
                      double* __restrict__ partials /* [grid][q1*q1] */) {
  constexpr int NP = NB * (NB + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);      // [NP][64]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, k = lane & 3;
  const int q1 = p + t + 1;
  const double* colp[NB];
  int kind[NB];                                           // 0 load, 1 ones, 2 zero padding
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int c = 8 * b + g;
    kind[b] = c < p + t ? 0 : (c == p + t ? (mask ? 0 : 1) : 2);
    colp[b] = c < p ? X + (int64_t)c * ldx : (c < p + t ? Y + (int64_t)(c - p) * ldy : (mask ? mask : X));
  }
  double acc[NP][2];
#pragma unroll
  for (int i = 0; i < NP; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }

  double2 buf[DEPTH][NB];
  double2 wb[DEPTH];
  const int64_t S = (int64_t)gridDim.x * 4 * 8;
  const int64_t r0 = ((int64_t)blockIdx.x * 4 + warp) * 8;
  int64_t rl = r0 + 2 * k;                                // this lane's first row of the NEXT batch to load
  const double* wp = WEIGHTED ? w + rl : nullptr;
#pragma unroll
  for (int b = 0; b < NB; ++b) colp[b] += rl;
  auto load_batch = [&](double2* z, double2& wv) {        // n is a multiple of 8 here: a batch is all in or all out
    const bool in = rl < n;
    if constexpr (WEIGHTED) {
      wv = make_double2(0.0, 0.0);
      if (in) wv = __ldcs(reinterpret_cast<const double2*>(wp));
      wp += S;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      double2 v = make_double2(0.0, 0.0);
      if (in) {
        if (kind[b] == 0) v = __ldcs(reinterpret_cast<const double2*>(colp[b]));
        else if (kind[b] == 1) v = make_double2(1.0, 1.0);
      }
      z[b] = v;
      colp[b] += S;
    }
    rl += S;
  };
  auto multiply = [&](const double2* z, const double2& wv) {
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      const double a0 = WEIGHTED ? z[i].x * wv.x : z[i].x;
      const double a1 = WEIGHTED ? z[i].y * wv.y : z[i].y;
#pragma unroll
      for (int j = i; j < NB; ++j) {
        dmma884(acc[idx][0], acc[idx][1], a0, z[j].x);
        dmma884(acc[idx][0], acc[idx][1], a1, z[j].y);
        ++idx;
      }
    }
  };

#pragma unroll
  for (int d = 0; d < DEPTH - 1; ++d) load_batch(buf[d], wb[d]);
  for (int64_t rc = r0; rc < n; rc += DEPTH * S) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      load_batch(buf[(d + DEPTH - 1) % DEPTH], wb[(d + DEPTH - 1) % DEPTH]);
      multiply(buf[d], wb[d]);
    }
  }
  // ---- CTA reduction in a fixed order: warp 0 stores, warps 1..3 add one after the other ----
  for (int wturn = 0; wturn < 4; ++wturn) {
    if (warp == wturn) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double* d = sm + i * 64 + g * 8 + 2 * k;
        if (wturn == 0) { d[0] = acc[i][0]; d[1] = acc[i][1]; }
        else { d[0] += acc[i][0]; d[1] += acc[i][1]; }
      }
    }
    __syncthreads();
  }
  double* out = partials + (size_t)blockIdx.x * q1 * q1;
  for (int e = threadIdx.x; e < NP * 64; e += 128) {
    const int blk = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
    int i = 0, rem = blk;
    while (rem >= NB - i) { rem -= NB - i; ++i; }
    const int j = i + rem;
    const int a = 8 * i + rr, b = 8 * j + cc;
    if (a < q1 && b < q1 && (i != j || a <= b)) {
      const double v = sm[e];
      out[(size_t)a * q1 + b] = v;
      out[(size_t)b * q1 + a] = v;
    }
  }
}

template <int NB> struct WideCfg {
  static constexpr int DEPTH = NB <= 2 ? 6 : (NB <= 4 ? 4 : (NB == 5 ? 3 : 2));
  static constexpr int MINB = NB <= 2 ? 4 : (NB <= 6 ? 3 : 2);
};

// ------------------------------------------------------------------------------------------------------------
// Staged variant (the production f64 path for n >= 4096 rows): the direct kernel above loads one 8-byte scalar per lane
// and 4-row step, i.e. eight 32-byte sectors of eight different columns per load instruction, with 4 rows in flight per
// warp — 33 % of the HBM peak at 33 columns (round 1).  Here every CTA streams 128-row tiles: one elected thread issues a
// 1-D bulk async copy (cp.async.bulk, the TMA engine) per column — 1 KiB contiguous each — into a shared-memory tile
// [column][132] (pitch 132 doubles: the 8-column x 4-row fragment read of a warp is bank-conflict-free), completion on an
// mbarrier ring of STAGES tiles.  Warps take 16 rows of a tile each: per 4-row step NB LDS.64 feed the same
// NB (NB + 1) / 2 DMMAs as before.  Rows in flight per CTA: STAGES x 128.  Only whole 128-row tiles are handled here; the
// tail (< 128 rows) goes through the direct kernel into further partials, and a fixed-order reduce joins both.
constexpr int DT_R = 128;                 // rows per tile
constexpr int DT_RP = 132;                // pitch in doubles (132 * 8 B = 32 B mod 128 B)

__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NB>
__global__ void __launch_bounds__(256)
gram_dmma_staged_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy,
                        const double* __restrict__ w, const double* __restrict__ mask, int64_t ntiles, int p, int t,
                        int stages, double* __restrict__ partials /* [grid][q1*q1] */) {
  constexpr int NP = NB * (NB + 1) / 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int ncol = p + t + (w ? 1 : 0) + (mask ? 1 : 0);       // staged columns: data, then weights, then mask
  const int wcol = p + t, mcol = p + t + (w ? 1 : 0);
  const size_t tile_doubles = (size_t)ncol * DT_RP;
  double* tiles = reinterpret_cast<double*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)stages * tile_doubles);
  uint64_t* empty = full + stages;
  double* sm = tiles;                                          // reused for the CTA reduction after the main loop
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, k = lane & 3;
  const int q1 = p + t + 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(&empty[i])), "r"(8));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t my_tiles = ntiles > (int64_t)blockIdx.x ? (uint32_t)((ntiles - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
  auto issue = [&](uint32_t it) {      // one thread: all columns of tile `it` of this CTA -> stage it % stages
    const int st = it % stages;
    const int64_t r0 = ((int64_t)it * gridDim.x + blockIdx.x) * DT_R;
    const uint32_t bar = smem_addr_u32(&full[st]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(ncol * DT_R * 8)) : "memory");
    double* dst = tiles + (size_t)st * tile_doubles;
    for (int c = 0; c < ncol; ++c) {
      const double* src = c < p ? X + (int64_t)c * ldx + r0 : (c < p + t ? Y + (int64_t)(c - p) * ldy + r0 : ((w && c == wcol) ? w + r0 : mask + r0));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_addr_u32(dst + (size_t)c * DT_RP)), "l"(src), "r"((uint32_t)(DT_R * 8)), "r"(bar) : "memory");
    }
  };
  auto wait = [&](uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
      asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                   : "=r"(done) : "r"(smem_addr_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
      if (!done && spins > (1u << 24)) __trap();
    }
  };
  if (threadIdx.x == 0)
    for (uint32_t it = 0; it < my_tiles && it < (uint32_t)stages; ++it) issue(it);

  int kind[NB], cidx[NB];                 // 0 data column cidx, 1 ones / mask, 2 zero padding
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int c = 8 * b + g;
    kind[b] = c < p + t ? 0 : (c == p + t ? 1 : 2);
    cidx[b] = c < p + t ? c : 0;
  }
  double acc[NP][2];
#pragma unroll
  for (int i = 0; i < NP; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }

  for (uint32_t it = 0; it < my_tiles; ++it) {
    const int st = it % stages;
    const uint32_t ph = (it / stages) & 1;
    wait(&full[st], ph);
    const double* tile = tiles + (size_t)st * tile_doubles;
#pragma unroll
    for (int step = 0; step < 4; ++step) {
      const int r = warp * 16 + step * 4 + k;
      double z[NB];
      const double mk = mask ? tile[(size_t)mcol * DT_RP + r] : 1.0;
      const double wv = w ? tile[(size_t)wcol * DT_RP + r] : 1.0;
#pragma unroll
      for (int b = 0; b < NB; ++b) z[b] = kind[b] == 0 ? tile[(size_t)cidx[b] * DT_RP + r] : (kind[b] == 1 ? mk : 0.0);
      int idx = 0;
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const double a = z[i] * wv;
#pragma unroll
        for (int j = i; j < NB; ++j) { dmma884(acc[idx][0], acc[idx][1], a, z[j]); ++idx; }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr_u32(&empty[st])) : "memory");
    if (threadIdx.x == 0 && it + stages < my_tiles) {     // refill this stage once all 8 warps have left it
      wait(&empty[st], ph);
      issue(it + stages);
    }
  }
  __syncthreads();       // every tile consumed: the tile memory becomes the reduction scratch
  for (int wturn = 0; wturn < 8; ++wturn) {
    if (warp == wturn) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        double* d = sm + i * 64 + g * 8 + 2 * k;
        if (wturn == 0) { d[0] = acc[i][0]; d[1] = acc[i][1]; }
        else { d[0] += acc[i][0]; d[1] += acc[i][1]; }
      }
    }
    __syncthreads();
  }
  double* out = partials + (size_t)blockIdx.x * q1 * q1;
  for (int e = threadIdx.x; e < NP * 64; e += 256) {
    const int blk = e >> 6, rr = (e >> 3) & 7, cc = e & 7;
    int i = 0, rem = blk;
    while (rem >= NB - i) { rem -= NB - i; ++i; }
    const int j = i + rem;
    const int a = 8 * i + rr, b = 8 * j + cc;
    if (a < q1 && b < q1 && (i != j || a <= b)) {
      const double v = sm[e];
      out[(size_t)a * q1 + b] = v;
      out[(size_t)b * q1 + a] = v;
    }
  }
}

template <int NB>
static int launch_dmma_staged(const double* X, int64_t ldx, const double* Y, int64_t ldy, const double* w, const double* mask,
                              int64_t ntiles, int p, int t, int grid, int stages, size_t smem, double* partials, cudaStream_t s) {
  auto k = gram_dmma_staged_kernel<NB>;
  PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, ntiles, p, t, stages, partials);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// m16n8k8 variant (production): the same data flow as gram_dmma_kernel with the larger FP64 tensor-core shape.
// An 8-row step gives every lane two elements per 8-column block, Z~[k0 + l%4][8b + l/4] and Z~[k0 + l%4 + 4][8b + l/4];
// they are the B fragment (8 x 8, "col") of block b and one half of the A fragment (16 x 8, "row") of the block pair
// that contains b.  A block pair I (16 columns) meets block J (8 columns) in one mma.sync.m16n8k8 (2048 flops); only
// the pairs that touch the upper triangle (J >= 2 I) are issued: 9 instructions per 8 rows at 33..40 columns, where the
// m8n8k4 kernel issues 30.  Round 2 measured the m8n8k4 kernel at 33 % of the HBM peak with the FP64 pipe far from its
// peak: the small shape is issue-limited.
__device__ __forceinline__ void dmma1688(double* c, double a0, double a1, double a2, double a3, double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}

template <int NB> struct Dmma16 {
  static constexpr int NI = (NB + 1) / 2;                 // 16-column block pairs
  static constexpr int count() { int c = 0; for (int i = 0; i < NI; ++i) for (int j = 2 * i; j < NB; ++j) ++c; return c; }
  static constexpr int NMMA = count();
};

constexpr int D16_WARPS = 4;   // 128-thread CTAs: 152 registers at 33..40 columns -> 3 CTAs (12 warps) per SM
template <int NB>
__global__ void __launch_bounds__(D16_WARPS * 32)
gram_dmma16_kernel(const double* __restrict__ X, int64_t ldx, const double* __restrict__ Y, int64_t ldy,
                   const double* __restrict__ w, const double* __restrict__ mask, int64_t n, int p, int t,
                   double* __restrict__ partials /* [grid][q1*q1] */) {
  constexpr int NI = Dmma16<NB>::NI, NMMA = Dmma16<NB>::NMMA;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);      // [NMMA][128] CTA-level sum of the warps' accumulators
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, k = lane & 3;
  const int q1 = p + t + 1;
  const double* colp[NB];
  int kind[NB];                                           // 0 data, 1 ones / mask, 2 zero padding
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int c = 8 * b + g;
    kind[b] = c < p + t ? 0 : (c == p + t ? 1 : 2);
    colp[b] = c < p ? X + (int64_t)c * ldx : (c < p + t ? Y + (int64_t)(c - p) * ldy : X);
  }
  double acc[NMMA][4];
#pragma unroll
  for (int i = 0; i < NMMA; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; acc[i][2] = 0.0; acc[i][3] = 0.0; }

  const int64_t stride = (int64_t)gridDim.x * D16_WARPS * 8;
  auto load_step = [&](int64_t r0, double (*z)[2], double* wv) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t r = r0 + k + 4 * h;
      const bool in = r < n;
      wv[h] = (in && w) ? w[r] : 1.0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        double v = 0.0;
        if (in) {
          if (kind[b] == 0) v = colp[b][r];
          else if (kind[b] == 1) v = mask ? mask[r] : 1.0;
        }
        z[b][h] = v;
      }
    }
  };
  int64_t r0 = ((int64_t)blockIdx.x * D16_WARPS + warp) * 8;
  double z[NB][2], zn[NB][2], wv[2] = {1.0, 1.0}, wn[2] = {1.0, 1.0};
  if (r0 < n) load_step(r0, z, wv);
  for (; r0 < n; r0 += stride) {
    const bool more = r0 + stride < n;
    if (more) load_step(r0 + stride, zn, wn);             // next step's loads fly while this step's DMMAs issue
    int idx = 0;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      // A fragment of the block pair (2i, 2i+1): a0 (row g, k), a1 (row g+8, k), a2 (row g, k+4), a3 (row g+8, k+4)
      const double a0 = z[2 * i][0] * wv[0], a2 = z[2 * i][1] * wv[1];
      const double a1 = (2 * i + 1 < NB) ? z[2 * i + 1 < NB ? 2 * i + 1 : 0][0] * wv[0] : 0.0;
      const double a3 = (2 * i + 1 < NB) ? z[2 * i + 1 < NB ? 2 * i + 1 : 0][1] * wv[1] : 0.0;
#pragma unroll
      for (int j = 2 * i; j < NB; ++j) { dmma1688(acc[idx], a0, a1, a2, a3, z[j][0], z[j][1]); ++idx; }
    }
    if (more) {
#pragma unroll
      for (int b = 0; b < NB; ++b) { z[b][0] = zn[b][0]; z[b][1] = zn[b][1]; }
      wv[0] = wn[0]; wv[1] = wn[1];
    }
  }
  // ---- CTA reduction in a fixed order: warp 0 stores, the other warps add one after the other ----
  // C fragment of (I, J): c0 (row g, col 2k), c1 (g, 2k+1), c2 (g+8, 2k), c3 (g+8, 2k+1); tile slot = row * 8 + col
  for (int wturn = 0; wturn < D16_WARPS; ++wturn) {
    if (warp == wturn) {
#pragma unroll
      for (int i = 0; i < NMMA; ++i) {
        double* d = sm + i * 128;
        const int s0 = g * 8 + 2 * k, s1 = (g + 8) * 8 + 2 * k;
        if (wturn == 0) { d[s0] = acc[i][0]; d[s0 + 1] = acc[i][1]; d[s1] = acc[i][2]; d[s1 + 1] = acc[i][3]; }
        else { d[s0] += acc[i][0]; d[s0 + 1] += acc[i][1]; d[s1] += acc[i][2]; d[s1 + 1] += acc[i][3]; }
      }
    }
    __syncthreads();
  }
  // ---- this CTA's partial, full symmetric q1 x q1 ----
  double* out = partials + (size_t)blockIdx.x * q1 * q1;
  for (int e = threadIdx.x; e < NMMA * 128; e += D16_WARPS * 32) {
    const int blk = e >> 7, rr = (e >> 3) & 15, cc = e & 7;
    int i = 0, rem = blk;
    while (rem >= NB - 2 * i) { rem -= NB - 2 * i; ++i; }
    const int j = 2 * i + rem;
    const int a = 16 * i + rr, b = 8 * j + cc;
    if (a < q1 && b < q1 && a <= b) {
      const double v = sm[e];
      out[(size_t)a * q1 + b] = v;
      out[(size_t)b * q1 + a] = v;
    }
  }
}

template <int NB>
static int launch_dmma16(const double* X, int64_t ldx, const double* Y, int64_t ldy, const double* w, const double* mask,
                         int64_t n, int p, int t, int grid, double* partials, cudaStream_t s) {
  const size_t smem = (size_t)Dmma16<NB>::NMMA * 128 * sizeof(double);
  auto k = gram_dmma16_kernel<NB>;
  if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, D16_WARPS * 32, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, partials);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template <int NB>
static int launch_dmma(const double* X, int64_t ldx, const double* Y, int64_t ldy, const double* w, const double* mask,
                       int64_t n, int p, int t, int grid, double* partials, cudaStream_t s) {
  const size_t smem = (size_t)(NB * (NB + 1) / 2) * 64 * sizeof(double);
  gram_dmma_kernel<NB><<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, partials);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template <int NB>
static int launch_dmma_wide(const double* X, int64_t ldx, const double* Y, int64_t ldy, const double* w, const double* mask,
                            int64_t n, int p, int t, int grid, double* partials, cudaStream_t s) {
  const size_t smem = (size_t)(NB * (NB + 1) / 2) * 64 * sizeof(double);
  if (w) gram_dmma_wide_kernel<NB, WideCfg<NB>::DEPTH, WideCfg<NB>::MINB, true><<<grid, 128, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, partials);
  else gram_dmma_wide_kernel<NB, WideCfg<NB>::DEPTH, WideCfg<NB>::MINB, false><<<grid, 128, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, partials);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

// 0 ok, 1 error, -1 not applicable (caller uses the DFMA kernel).  PDSB_K2A_DMMA=0 disables.
static int moments_dmma_f64(const double* X, int64_t ldx, const double* Y, int64_t ldy, const double* w,
                            const double* mask, int64_t n, int p, int t, double* M, cudaStream_t s) {
  static const bool enabled = [] { const char* e = getenv("PDSB_K2A_DMMA"); return !(e && e[0] == '0'); }();
  const int q1 = p + t + 1;
  const int nb = (q1 + 7) / 8;
  if (!enabled || nb < 1 || nb > 8 || n < 1) return -1;
  // staged kernel for the whole 128-row tiles (needs 16-byte aligned columns), direct kernel for the rest
  // PDSB_K2A_KERNEL: 8 (default) = m8n8k4 direct, 16 = m16n8k8 direct, 0 = bulk-copy staged m8n8k4 (+ m8n8k4 tail).
  // Measured in one call (B200, 2e7 x 33 f64 / 5e7 x 9 f64, profiles/r02/k2a_f64.txt): m8n8k4 direct 33.4 % / 30.0 % of the
  // HBM peak, m16n8k8 direct 26.3 % / 24.6 %, staged 22.9 % / 31.0 %.
  static const int kern = [] { const char* e = getenv("PDSB_K2A_KERNEL"); return e ? atoi(e) : 2; }();
  const bool staged_on = kern == 0;
  const int ncol = p + t + (w ? 1 : 0) + (mask ? 1 : 0);
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const bool aligned = al16(X) && al16(Y) && (ldx % 2 == 0) && (ldy % 2 == 0) && (!w || al16(w)) && (!mask || al16(mask));
  const size_t tile_bytes = (size_t)ncol * DT_RP * sizeof(double);
  int stages = (int)std::min<size_t>(4, (200 * 1024 - 256) / tile_bytes);
  const size_t red_bytes = (size_t)(nb * (nb + 1) / 2) * 64 * sizeof(double);
  const bool wide = kern == 2 && aligned && n >= 8;       // whole 8-row batches; the last n % 8 rows go to the direct kernel
  const int64_t ntiles = (staged_on && aligned && stages >= 2 && n >= 4096) ? n / DT_R : 0;
  const int64_t n_main = wide ? n - n % 8 : ntiles * DT_R;
  const int64_t n_tail = n - n_main;
  int grid_main = 0, grid = 0;
  if (wide) grid_main = (int)std::min<int64_t>(ceil_div(n_main, 32), (int64_t)sm_count() * (nb <= 2 ? 4 : (nb <= 6 ? 3 : 2)));
  if (ntiles > 0) {
    // one or two CTAs per SM, whatever shared memory allows
    const size_t smem_need = std::max((size_t)stages * tile_bytes, red_bytes) + 2 * stages * sizeof(uint64_t) + 128;
    const int per_sm = smem_need <= 100 * 1024 ? 2 : 1;
    grid_main = (int)std::min<int64_t>(ntiles, (int64_t)sm_count() * per_sm);
  }
  if (n_tail > 0) {
    if (kern == 16) grid = (int)std::min<int64_t>(ceil_div(n_tail, 8 * D16_WARPS), (int64_t)sm_count() * (nb <= 4 ? 4 : (nb <= 5 ? 3 : 2)));
    else grid = (int)std::min<int64_t>(ceil_div(n_tail, 32), (int64_t)sm_count() * 4);
  }
  if (grid_main + grid < 1) grid = 1;
  double* partials = nullptr;
  if (dev_alloc((void**)&partials, (size_t)(grid_main + grid) * q1 * q1 * sizeof(double), s)) return 1;
  int rc = 0;
  if (wide) {
    switch (nb) {
      case 1: rc = launch_dmma_wide<1>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 2: rc = launch_dmma_wide<2>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 3: rc = launch_dmma_wide<3>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 4: rc = launch_dmma_wide<4>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 5: rc = launch_dmma_wide<5>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 6: rc = launch_dmma_wide<6>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      case 7: rc = launch_dmma_wide<7>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
      default: rc = launch_dmma_wide<8>(X, ldx, Y, ldy, w, mask, n_main, p, t, grid_main, partials, s); break;
    }
    if (rc) { dev_free(partials, s); return rc; }
  } else if (grid_main > 0) {
    const size_t smem = std::max((size_t)stages * tile_bytes, red_bytes) + 2 * stages * sizeof(uint64_t) + 128;
#define PDSB_ST(NBV) rc = launch_dmma_staged<NBV>(X, ldx, Y, ldy, w, mask, ntiles, p, t, grid_main, stages, smem, partials, s)
    switch (nb) {
      case 1: PDSB_ST(1); break; case 2: PDSB_ST(2); break; case 3: PDSB_ST(3); break; case 4: PDSB_ST(4); break;
      case 5: PDSB_ST(5); break; case 6: PDSB_ST(6); break; case 7: PDSB_ST(7); break; default: PDSB_ST(8); break;
    }
#undef PDSB_ST
    if (rc) { dev_free(partials, s); return rc; }
  }
  if (grid > 0) {
    const double* Xt = X + n_main; const double* Yt = Y + n_main;
    const double* wt = w ? w + n_main : nullptr; const double* mt = mask ? mask + n_main : nullptr;
    double* pt = partials + (size_t)grid_main * q1 * q1;
    if (kern == 16) {
      switch (nb) {
        case 1: rc = launch_dmma16<1>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 2: rc = launch_dmma16<2>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 3: rc = launch_dmma16<3>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 4: rc = launch_dmma16<4>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 5: rc = launch_dmma16<5>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 6: rc = launch_dmma16<6>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        case 7: rc = launch_dmma16<7>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
        default: rc = launch_dmma16<8>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      }
    } else
    switch (nb) {
      case 1: rc = launch_dmma<1>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 2: rc = launch_dmma<2>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 3: rc = launch_dmma<3>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 4: rc = launch_dmma<4>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 5: rc = launch_dmma<5>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 6: rc = launch_dmma<6>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      case 7: rc = launch_dmma<7>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
      default: rc = launch_dmma<8>(Xt, ldx, Yt, ldy, wt, mt, n_tail, p, t, grid, pt, s); break;
    }
    if (rc) { dev_free(partials, s); return rc; }
  }
  grid += grid_main;
  const int len = q1 * q1;
  reduce_partials_kernel<<<len, 128, 0, s>>>(partials, grid, len, M);
  cudaError_t e = cudaGetLastError();
  count_launch();
  dev_free(partials, s);
  if (e != cudaSuccess) { set_error("reduce_partials launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

template <typename T, int MAXT>
static int launch_gram(const T* X, int64_t ldx, const T* Y, int64_t ldy, const T* w, const T* mask,
                       int64_t n, int p, int t, int tile_r, int S, int grid, size_t smem, double* partials,
                       int64_t bstride, cudaStream_t s) {
  if (w) {
    auto k = gram_simt_kernel<T, MAXT, true>;
    PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, partials, bstride);
  } else {
    auto k = gram_simt_kernel<T, MAXT, false>;
    PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, partials, bstride);
  }
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template <typename T>
int moments_simt(const T* X, int64_t ldx, const T* Y, int64_t ldy, const T* w, const T* mask, int64_t n,
                 int p, int t, double* M, cudaStream_t s, int64_t bstride) {
  const int q1 = p + t + 1;
  if (p < 0 || t < 0 || q1 > 260) { set_error("moments: p+t+1=%d out of range (max 260)", q1); return 1; }
  if constexpr (sizeof(T) == 8) {
    if (bstride == 0) {      // column-major f64: FP64 tensor-core path for up to 64 columns
      const int rc = moments_dmma_f64(X, ldx, Y, ldy, w, mask, n, p, t, M, s);
      if (rc >= 0) return rc;
    }
  }
  const int nt = (q1 + 3) / 4, ntp = nt * (nt + 1) / 2;
  const int maxt = (ntp + 255) / 256;
  const int S = nt * 4;  // row stride, multiple of 4 elements
  // tile rows: keep the tile under ~64 KB
  int tile_r = 128;
  while (tile_r > 16 && (size_t)tile_r * (S + 1) * sizeof(T) > 64 * 1024) tile_r >>= 1;
  size_t smem = (size_t)tile_r * (S + 1) * sizeof(T);
  int64_t ntiles = ceil_div(n > 0 ? n : 1, tile_r);
  // persistent grid: as many CTAs per SM as registers / shared memory allow (load and FMA phases of different CTAs overlap)
  int per_sm = 2;
  {
    int occ = 0;
    cudaError_t oe = cudaSuccess;
    if (maxt == 1) {
      if (w) oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gram_simt_kernel<T, 1, true>, 256, smem);
      else oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gram_simt_kernel<T, 1, false>, 256, smem);
      if (oe == cudaSuccess && occ >= 1) per_sm = std::min(occ, 4);
      else (void)cudaGetLastError();
    }
  }
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count() * per_sm);
  if (grid < 1) grid = 1;
  const int nslices = (maxt == 1) ? 256 / ntp : 1;      // must match the kernel's decode
  const int nparts = grid * nslices;
  double* partials = nullptr;
  if (dev_alloc((void**)&partials, (size_t)nparts * q1 * q1 * sizeof(double), s)) return 1;
  int rc;
  switch (maxt) {
    case 1: rc = launch_gram<T, 1>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    case 2: rc = launch_gram<T, 2>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    case 3: case 4: rc = launch_gram<T, 4>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    default: rc = launch_gram<T, 9>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
  }
  if (rc) { dev_free(partials, s); return rc; }
  int len = q1 * q1;
  reduce_partials_kernel<<<len, 128, 0, s>>>(partials, nparts, len, M);
  cudaError_t e = cudaGetLastError();
  count_launch();
  dev_free(partials, s);
  if (e != cudaSuccess) { set_error("reduce_partials launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

template int moments_simt<float>(const float*, int64_t, const float*, int64_t, const float*, const float*,
                                 int64_t, int, int, double*, cudaStream_t, int64_t);
template int moments_simt<double>(const double*, int64_t, const double*, int64_t, const double*, const double*,
                                  int64_t, int, int, double*, cudaStream_t, int64_t);

}  // namespace pdsb
