// K5 — group_by(seg).agg(pds.lin_reg(...)) as a batched, segmented problem.
//
// In the reference there is no grouped code: Polars calls `_polars_plugin_pl_lr` once per group from its rayon pool
// (/root/reference/src/utils/mod.rs:81-84 "tight group-by loops"; tests/test_linear_exprs.py:918-953).
// Each call packs the group, builds X'X / X'y with a sequential matmul (lr_solvers.rs:186-190) and runs the gated
// QR solve (:329-382).  Here the whole frame stays in HBM in key order and one launch sequence does all groups:
//   pass 0  work list: a group of len rows becomes ceil(len / CHUNK) items of EQUAL size; item -> rows map (one thread per item);
//   pass 1  group moments: one warp per item, taken from a global counter by the warps of a persistent grid; lanes stride
//           rows (coalesced column reads), the (p+2)(p+3)/2 moments live in registers (f32 FMA chains of <= CHUNK/32
//           terms -> f64 warp reduce); item partials are combined in a fixed order -> bit-reproducible whatever warp ran them;
//   pass 2  batched solve: one thread per group on an interleaved (component-major) workspace — shared memory of a
//           32-thread CTA up to ~20 coefficients, global memory (coalesced across groups) above: ridge, rank gate
//           (ln|det| - sum ln diag <= ln tol), pivoted Householder QR (or Cholesky for solver="choleskey"),
//           back-substitution, coordinate descent / NNLS per group.
// HBM-bound: algorithmic bytes per row = (p+1) * s.
#include "../common.h"
#include "kernels.h"
#include <cstdlib>

namespace pdsb {

namespace {

constexpr int CHUNK = 8192;  // rows per work item of the whole-frame caller; upper bound per (group, chunk) item
// PDSB_K5_CHUNK: rows per (group, chunk) item of the grouped path (a group of len rows is cut into ceil(len / chunk)
// items of EQUAL size, see item_range)
static int64_t chunk_rows() {
  static const int64_t v = [] { const char* e = getenv("PDSB_K5_CHUNK"); const long c = e ? atol(e) : 0; return (int64_t)(c >= 128 ? c : CHUNK); }();
  return v;
}

// rows [r0, r1) of item `item` of group g: the group's rows are dealt EVENLY over its items (rounded up to 128 rows, one
// trip of the register kernel).  Cutting at fixed CHUNK boundaries left a 1e4-row group as 8192 + 1808 rows, and with
// items dealt round-robin over an even number of warps half the warps got only the long ones (56 % of the HBM peak).
__device__ __forceinline__ void item_range(const int64_t* __restrict__ offsets, const int64_t* __restrict__ item_start, int64_t g,
                                           int64_t item, int64_t& r0, int64_t& r1) {
  const int64_t beg = offsets[g], end = offsets[g + 1];
  const int64_t cnt = item_start[g + 1] - item_start[g];
  const int64_t per = (((end - beg) + cnt - 1) / cnt + 127) & ~int64_t(127);
  r0 = min(beg + (item - item_start[g]) * per, end);
  r1 = min(r0 + per, end);
}
constexpr int GEN_ACC = 70;   // moments per lane of the generic-p kernel: (p+2)(p+3)/2 <= 2240 -> p <= 64

// ---------------- pass 0: work list ----------------
__global__ void count_items_kernel(const int64_t* __restrict__ offsets, int64_t n_groups, int64_t chunk, int64_t* __restrict__ item_start) {
  // item_start[g] = number of chunks of group g (scanned on the host side of this file by a tiny kernel below)
  for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (int64_t)gridDim.x * blockDim.x) {
    int64_t len = offsets[g + 1] - offsets[g];
    item_start[g] = len > 0 ? (len + chunk - 1) / chunk : 1;
  }
}

__global__ void scan_items_kernel(int64_t* __restrict__ item_start, int64_t n_groups) {
  // single block exclusive scan with running carry (n_groups is at most a few million)
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_groups + 1; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    int64_t v = (i < n_groups) ? item_start[i] : 0;
    int64_t inc = v;
    for (int off = 1; off < 32; off <<= 1) { int64_t y = __shfl_up_sync(0xffffffffu, inc, off); if ((threadIdx.x & 31) >= off) inc += y; }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
    __syncthreads();
    int64_t wpre = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wpre += warp_tot[w];
    int64_t excl = carry + wpre + inc - v;
    __syncthreads();
    if (i <= n_groups) item_start[i] = excl;
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
}

// ---------------- pass 1: per-(group, chunk) moments ----------------
// Z = [x_0 .. x_{P-1}, y, 1];  Q1 = P + 2 columns; NM = Q1 (Q1+1) / 2 packed upper-triangular moments.
// SKIP_NAN: null rows arrive as NaN and drop out of their group (the grouped path).  The whole-frame caller (moments_small)
// passes false: there a NaN must poison the moments exactly like in the other moments kernels.
// MODE bit 0: the item's rows come from a precomputed [n_items][2] array (item_rows_kernel) instead of a 14-step binary
// search in front of every item; bit 1: warps take items from a global counter (persistent grid) instead of a fixed
// stride, so a warp with a short item does not idle while its CTA's longest one finishes.
template <typename T, int P, bool SKIP_NAN = true, int MODE = 0>
__global__ void __launch_bounds__(256)
group_moments_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y,
                     const int64_t* __restrict__ offsets, const int64_t* __restrict__ item_start,
                     int64_t n_groups, int64_t n_items, double* __restrict__ part /* [NM][n_items] */,
                     const int64_t* __restrict__ rows = nullptr, unsigned long long* __restrict__ queue = nullptr) {
  constexpr int Q1 = P + 2;
  constexpr int NM = Q1 * (Q1 + 1) / 2;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  auto next_item = [&](int64_t cur) -> int64_t {
    if constexpr ((MODE & 2) != 0) {
      unsigned long long v = 0;
      if (lane == 0) v = atomicAdd(queue, 1ULL);
      return (int64_t)__shfl_sync(0xffffffffu, v, 0);
    } else {
      return cur + nwarps;
    }
  };
  for (int64_t item = (MODE & 2) ? next_item(0) : warp_global; item < n_items; item = next_item(item)) {
    int64_t r0, r1;
    if constexpr ((MODE & 1) != 0) {
      r0 = rows[2 * item];
      r1 = rows[2 * item + 1];
    } else {
      // binary search the group of this item: item_start[g] <= item < item_start[g+1]
      int64_t lo = 0, hi = n_groups;
      while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (item_start[mid] <= item) lo = mid; else hi = mid; }
      item_range(offsets, item_start, lo, item, r0, r1);
    }
    T acc[NM];
#pragma unroll
    for (int k = 0; k < NM; ++k) acc[k] = T(0);
    // U rows per lane and trip: all U x (P+1) loads are issued before the first FMA (memory-level parallelism is what
    // this kernel lives on: one row per trip left ~9 loads in flight per warp and 29 % of the HBM roofline)
    constexpr int U = 4;
    for (int64_t r = r0 + lane; r < r1; r += 32 * U) {
      T z[U][Q1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t rr = r + 32 * u;
        const int64_t rc = rr < r1 ? rr : r;          // clamped address; the row is zeroed below
#pragma unroll
        for (int c = 0; c < P; ++c) z[u][c] = X[(int64_t)c * ldx + rc];
        z[u][P] = y[rc];
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // null rows arrive as NaN (null_policy="skip"): they drop out of their group -> the whole row becomes 0
        T probe = T(0);
#pragma unroll
        for (int c = 0; c <= P; ++c) probe = fma(z[u][c], T(0), probe);
        const bool use = (r + 32 * u < r1) && (!SKIP_NAN || probe == T(0));
#pragma unroll
        for (int c = 0; c <= P; ++c) z[u][c] = use ? z[u][c] : T(0);
        z[u][P + 1] = use ? T(1) : T(0);
        int k = 0;
#pragma unroll
        for (int i = 0; i < Q1; ++i)
#pragma unroll
          for (int j = i; j < Q1; ++j) { acc[k] = fma(z[u][i], z[u][j], acc[k]); ++k; }
      }
    }
    // (a transpose-reduce in T — NM - NM / 32 shuffles instead of 5 f64 shuffles per value — was measured twice in round 2:
    // ptxas then keeps fewer of the U x (P + 1) loads in flight and the kernel drops from 4.2 to 2.5 TB/s)
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      double v = (double)acc[k];
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) part[(size_t)k * n_items + item] = v;
    }
  }
}

// whole-frame use of the register kernel (moments_small): one "group" = all rows, item i = rows [i CHUNK, (i+1) CHUNK)
__global__ void small_items_kernel(int64_t n, int64_t n_items, int64_t* __restrict__ offsets /* [2] */, int64_t* __restrict__ item_start /* [2] */) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { offsets[0] = 0; offsets[1] = n; item_start[0] = 0; item_start[1] = n_items; }
}
// packed moment k of Z = [x.., y, 1] summed over the items in a fixed order -> M[(i, j)] and M[(j, i)]
__global__ void __launch_bounds__(128) small_reduce_kernel(const double* __restrict__ part, int64_t n_items, int q1, double* __restrict__ M) {
  __shared__ double red[128];
  const int k = blockIdx.x;
  double v = 0.0;
  for (int64_t it = threadIdx.x; it < n_items; it += 128) v += part[(size_t)k * n_items + it];
  red[threadIdx.x] = v;
  __syncthreads();
  for (int off = 64; off; off >>= 1) { if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) {
    int i = 0, rem = k;
    while (rem >= q1 - i) { rem -= q1 - i; ++i; }
    const int j = i + rem;
    M[(size_t)i * q1 + j] = red[0];
    M[(size_t)j * q1 + i] = red[0];
  }
}

// ---------------- pass 1, staged (PDSB_K5_STAGED=1; measured slower than the register kernel, see launch_moments_p) -------
// ncu on the register kernel above (profiles/r02: C3 shape): 128 registers -> 16 warps per SM, the dominant stall is the
// long scoreboard (2.97 of 4.8 warp-cycles per issue): every byte in flight occupies a register, so the kernel cannot
// keep enough of them in flight to cover the HBM latency (43 % of the DRAM peak).  Here the bytes in flight live in
// shared memory instead: one producer thread per CTA streams the work items through a ring of STAGES tiles of TILE rows
// with 1-D bulk async copies (cp.async.bulk, one per column and tile, completion on an mbarrier), CONSUMERS warps read
// their rows from the tile (conflict-free: consecutive threads, consecutive rows of a column) into the same register
// moments as before.  An item's tiles start at the 16-byte boundary below its first row; rows outside [r0, r1) are
// masked.  The last < 16 bytes of a column that is not a whole number of 16-byte groups long are read from global memory.
constexpr int ST_TILE = 512;          // rows per stage
constexpr int ST_STAGES = 6;
constexpr int ST_CONSUMERS = 8;       // consumer warps per CTA (+ 1 producer warp)
constexpr int ST_THREADS = (ST_CONSUMERS + 1) * 32;

struct StageMeta { int lo, hi, last; int64_t item; int64_t row_begin; };

__device__ __forceinline__ uint32_t k5_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void k5_mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(done) : "r"(k5_smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}

// item -> (first row, end row): one thread per item, so that the staged kernel's single producer thread never runs the
// binary search (14 dependent global loads per item) in front of its copies
__global__ void item_rows_kernel(const int64_t* __restrict__ offsets, const int64_t* __restrict__ item_start, int64_t n_groups,
                                 int64_t n_items, int64_t* __restrict__ rows /* [n_items][2] */) {
  for (int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n_groups;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (item_start[mid] <= item) lo = mid; else hi = mid; }
    int64_t r0, r1;
    item_range(offsets, item_start, lo, item, r0, r1);
    rows[2 * item] = r0;
    rows[2 * item + 1] = r1;
  }
}

// packed pair helpers: a consumer thread owns the ADJACENT rows (2 tid, 2 tid + 1) of a tile, so one 8-byte (f32) load per
// column brings both and every moment costs one FFMA2 per two rows (sm_100 packed f32); f64 data keeps scalar arithmetic
template <typename T> struct Pair;
template <> struct Pair<float> {
  using type = float2;
  static __device__ __forceinline__ float2 load(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
  static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
};
template <> struct Pair<double> {
  using type = double2;
  static __device__ __forceinline__ double2 load(const double* p) { return *reinterpret_cast<const double2*>(p); }
  static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
  static __device__ __forceinline__ double2 fma(double2 a, double2 b, double2 c) { return make_double2(::fma(a.x, b.x, c.x), ::fma(a.y, b.y, c.y)); }
};

template <typename T, int P>
__global__ void __launch_bounds__(ST_THREADS)
group_moments_staged_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y, const int64_t* __restrict__ item_rows,
                            int64_t n_items, int64_t n, double* __restrict__ part /* [NM][n_items] */) {
  using PT = typename Pair<T>::type;
  constexpr int Q1 = P + 2;
  constexpr int NM = Q1 * (Q1 + 1) / 2;
  constexpr int V = 16 / (int)sizeof(T);
  static_assert(ST_TILE == ST_CONSUMERS * 32 * 2, "one row pair per consumer thread and stage");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);                                        // [STAGES][P + 1][TILE]
  uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)ST_STAGES * (P + 1) * ST_TILE);
  uint64_t* empty = full + ST_STAGES;
  StageMeta* meta = reinterpret_cast<StageMeta*>(empty + ST_STAGES);
  double* red = reinterpret_cast<double*>(meta + ST_STAGES);                         // [CONSUMERS][NM]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_floor = (n / V) * V;                     // rows below this are covered by whole 16-byte groups
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_STAGES; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k5_smem_u32(&full[i])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(k5_smem_u32(&empty[i])), "r"(ST_CONSUMERS));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == ST_CONSUMERS) {
    // ---------------- producer warp: lane 0 keeps the barriers and the stage metadata, lane c issues column c's copy
    // (a bulk-copy instruction costs its issuing thread ~50 cycles: nine of them from one thread were the pace of the ring)
    uint32_t st = 0, ph = 0;
    for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int64_t r0 = item_rows[2 * item], r1 = item_rows[2 * item + 1];
      const int64_t a0 = (r0 / V) * V;
      const int64_t ntiles = r1 > a0 ? (r1 - a0 + ST_TILE - 1) / ST_TILE : 1;
      for (int64_t tl = 0; tl < ntiles; ++tl) {
        k5_mbar_wait(&empty[st], ph ^ 1);
        const int64_t rb = a0 + tl * ST_TILE;
        int64_t rows = min((int64_t)ST_TILE, n_floor - rb);     // whole 16-byte groups only
        if (rows < 0) rows = 0;
        const uint32_t bytes = (uint32_t)(rows * (int64_t)sizeof(T));
        if (lane == 0) {
          StageMeta m;
          m.lo = (int)max(r0 - rb, (int64_t)0); m.hi = (int)min(r1 - rb, (int64_t)ST_TILE); m.last = (tl == ntiles - 1);
          m.item = item; m.row_begin = rb;
          meta[st] = m;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(k5_smem_u32(&full[st])), "r"(bytes * (P + 1)) : "memory");
        }
        __syncwarp();
        if (bytes && lane <= P) {
          T* tile = tiles + (size_t)st * (P + 1) * ST_TILE;
          const T* src = (lane < P ? X + (int64_t)lane * ldx : y) + rb;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(k5_smem_u32(tile + (size_t)lane * ST_TILE)), "l"(src), "r"(bytes), "r"(k5_smem_u32(&full[st])) : "memory");
        }
        if (++st == ST_STAGES) { st = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ---------------- consumers: thread tid owns rows 2 tid, 2 tid + 1 of every tile ----------------
  PT acc[NM];
#pragma unroll
  for (int k = 0; k < NM; ++k) acc[k] = Pair<T>::make(T(0), T(0));
  const int rr = 2 * threadIdx.x;
  uint32_t st = 0, ph = 0;
  for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    for (;;) {
      k5_mbar_wait(&full[st], ph);
      const StageMeta m = meta[st];
      const T* tile = tiles + (size_t)st * (P + 1) * ST_TILE;
      PT z[Q1];
      if (m.row_begin + rr + 1 < n_floor) {
#pragma unroll
        for (int c = 0; c <= P; ++c) z[c] = Pair<T>::load(tile + (size_t)c * ST_TILE + rr);
      } else {
        // the ragged end of the columns (fewer than 16 bytes were not copied): straight from global memory, clamped address
        const int64_t g0 = min(m.row_begin + rr, n - 1), g1 = min(m.row_begin + rr + 1, n - 1);
#pragma unroll
        for (int c = 0; c < P; ++c) z[c] = Pair<T>::make(X[(int64_t)c * ldx + g0], X[(int64_t)c * ldx + g1]);
        z[P] = Pair<T>::make(y[g0], y[g1]);
      }
      // null rows arrive as NaN (null_policy="skip"): they drop out of their group -> the whole row becomes 0
      PT probe = Pair<T>::make(T(0), T(0));
#pragma unroll
      for (int c = 0; c <= P; ++c) probe = Pair<T>::fma(z[c], Pair<T>::make(T(0), T(0)), probe);
      const bool use0 = (rr >= m.lo) && (rr < m.hi) && (probe.x == T(0));
      const bool use1 = (rr + 1 >= m.lo) && (rr + 1 < m.hi) && (probe.y == T(0));
      if (!__all_sync(0xffffffffu, use0 && use1)) {       // tile edges and null rows only
#pragma unroll
        for (int c = 0; c <= P; ++c) z[c] = Pair<T>::make(use0 ? z[c].x : T(0), use1 ? z[c].y : T(0));
      }
      z[P + 1] = Pair<T>::make(use0 ? T(1) : T(0), use1 ? T(1) : T(0));
      int k = 0;
#pragma unroll
      for (int i = 0; i < Q1; ++i)
#pragma unroll
        for (int j = i; j < Q1; ++j) { acc[k] = Pair<T>::fma(z[i], z[j], acc[k]); ++k; }
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(k5_smem_u32(&empty[st])) : "memory");
      if (++st == ST_STAGES) { st = 0; ph ^= 1; }
      if (m.last) break;
    }
    // the item is complete: warp tree sums in the data type (f64 arithmetic costs a warp 10-30 cycles per instruction on
    // this part: 8 warps x 55 moments x 5 f64 shuffle steps per item took as long as streaming the item), then the eight
    // warp sums of every moment are added in f64 in a fixed order (bit-reproducible)
#pragma unroll
    for (int k = 0; k < NM; ++k) {
      T v = acc[k].x + acc[k].y;
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp * NM + k] = (double)v;
      acc[k] = Pair<T>::make(T(0), T(0));
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ST_CONSUMERS * 32) : "memory");
    for (int k = threadIdx.x; k < NM; k += ST_CONSUMERS * 32) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < ST_CONSUMERS; ++w) v += red[w * NM + k];
      part[(size_t)k * n_items + item] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ST_CONSUMERS * 32) : "memory");
  }
}

template <typename T, int P>
constexpr size_t staged_smem() {
  return (size_t)ST_STAGES * (P + 1) * ST_TILE * sizeof(T) + 2 * ST_STAGES * sizeof(uint64_t) + ST_STAGES * sizeof(StageMeta) +
         (size_t)ST_CONSUMERS * ((P + 2) * (P + 3) / 2) * sizeof(double) + 128;
}

// generic-P variant: lanes still stride rows but moments are accumulated through shared memory per warp
template <typename T>
__global__ void __launch_bounds__(128)
group_moments_generic_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ y,
                             const int64_t* __restrict__ offsets, const int64_t* __restrict__ item_start,
                             int64_t n_groups, int64_t n_items, int p, double* __restrict__ part) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int q1 = p + 2;
  const int nm = q1 * (q1 + 1) / 2;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  T* tile = reinterpret_cast<T*>(smem_raw) + (size_t)wid * 32 * (q1 + 1);   // [32][q1+1]
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int per_lane = (nm + 31) / 32;
  for (int64_t item = warp_global; item < n_items; item += nwarps) {
    int64_t lo = 0, hi = n_groups;
    while (hi - lo > 1) { int64_t mid = (lo + hi) >> 1; if (item_start[mid] <= item) lo = mid; else hi = mid; }
    const int64_t g = lo;
    int64_t r0, r1;
    item_range(offsets, item_start, g, item, r0, r1);
    double acc[GEN_ACC];   // supports nm <= 32 * GEN_ACC (p <= 64)
    for (int k = 0; k < GEN_ACC; ++k) acc[k] = 0.0;
    for (int64_t rb = r0; rb < r1; rb += 32) {
      const int64_t r = rb + lane;
      bool fin = r < r1;
      for (int c = 0; c < q1; ++c) {
        T v = T(0);
        if (r < r1) v = (c < p) ? X[(int64_t)c * ldx + r] : (c == p ? y[r] : T(1));
        fin = fin && isfinite(v);
        tile[lane * (q1 + 1) + c] = v;
      }
      if (!fin) for (int c = 0; c < q1; ++c) tile[lane * (q1 + 1) + c] = T(0);
      __syncwarp();
      for (int m = 0; m < per_lane; ++m) {
        int k = lane + 32 * m;
        if (k < nm) {
          int i = 0, rem = k;
          while (rem >= q1 - i) { rem -= q1 - i; ++i; }
          int j = i + rem;
          T s = T(0);
          for (int rr = 0; rr < 32; ++rr) s = fma(tile[rr * (q1 + 1) + i], tile[rr * (q1 + 1) + j], s);
          acc[m] += (double)s;
        }
      }
      __syncwarp();
    }
    for (int m = 0; m < per_lane; ++m) {
      int k = lane + 32 * m;
      if (k < nm) part[(size_t)k * n_items + item] = acc[m];
    }
  }
}

// ---------------- pass 2: batched solve, one thread per group, interleaved workspace ----------------
struct GroupSolveArgs {
  const double* part; const int64_t* item_start; const int64_t* offsets;
  int64_t n_groups, n_items; int p, add_bias, solver; double l2, tol;
  int method, positive, max_iter; double l1, cd_tol;      // PDSB_METHOD_CD / _NNLS per group (lr_solvers.rs:426-600)
  double* ws;      // [(q*q + 2q) ][n_groups]  interleaved; nullptr: the workspace lives in shared memory, [(q*q + 2q)][blockDim.x]
  double* beta;    // [n_groups][q]
  int* status;
};

__global__ void __launch_bounds__(128) group_solve_kernel(GroupSolveArgs a) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= a.n_groups) return;
  const int p = a.p, q = p + (a.add_bias ? 1 : 0), q1 = p + 2;
  // one thread per group walks dependent chains over its q x q system: in global memory every step pays a DRAM / L2 round
  // trip (0.113 ms for 1e4 groups of 9 coefficients, 14 % of the C3 step); up to ~20 coefficients the workspace of a
  // 32-thread CTA fits shared memory instead (same interleaved layout, stride = threads per CTA)
  extern __shared__ double gs_smem[];
  const int64_t NG = a.ws ? a.n_groups : (int64_t)blockDim.x;
  double* base = a.ws ? a.ws + g : gs_smem + threadIdx.x;
  double* A = base;                                      // A(i,j) = A[(i + j*q) * NG]
  double* b = base + (size_t)q * q * NG;                 // b(i)   = b[i * NG]
  double* cn = base + ((size_t)q * q + q) * NG;          // scratch q
#define AA(i, j) A[((size_t)(i) + (size_t)(j) * q) * NG]
#define BB(i) b[(size_t)(i) * NG]
#define CN(i) cn[(size_t)(i) * NG]
  // moment (i,j) of Z=[x.., y, 1], i<=j, packed index
  auto midx = [&](int i, int j) { if (i > j) { int t = i; i = j; j = t; } return i * q1 - i * (i - 1) / 2 + (j - i); };
  auto fz = [&](int i) { return i < p ? i : p + 1; };    // coefficient index -> Z column (bias -> ones)
  const int64_t it0 = a.item_start[g], it1 = a.item_start[g + 1];
  const double* __restrict__ part = a.part;
  const int nm = q1 * (q1 + 1) / 2;
  // the group's moments, summed over its items in a fixed order.  With the shared-memory workspace they are gathered
  // first, eight independent loads at a time (a load followed at once by its consumer serialises on the in-order issue:
  // 64 dependent round trips were 40 of the kernel's 97 microseconds at 9 coefficients)
  double* ms = a.ws ? nullptr : base + ((size_t)q * q + 2 * q) * NG;     // ms(k) = ms[k * NG]
  if (ms) {
    for (int k0 = 0; k0 < nm; k0 += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int k = k0 + u < nm ? k0 + u : nm - 1;
        v[u] = part[(size_t)k * a.n_items + it0];
      }
      for (int64_t it = it0 + 1; it < it1; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k0 + u < nm ? k0 + u : nm - 1;
          v[u] += part[(size_t)k * a.n_items + it];
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < nm) ms[(size_t)(k0 + u) * NG] = v[u];
    }
  }
  auto mom = [&](int k) {
    if (ms) return ms[(size_t)k * NG];
    double s = 0.0;
    for (int64_t it = it0; it < it1; ++it) s += part[(size_t)k * a.n_items + it];
    return s;
  };
  const int64_t nrows = (int64_t)llround(mom(midx(p + 1, p + 1)));   // valid rows of the group
  double* out = a.beta + (size_t)g * q;
  if (nrows < q) {  // "#Data < #features" would be an error for a single call; per group it yields null
    a.status[g] = PDSB_GATED;
    for (int i = 0; i < q; ++i) out[i] = nan("");
    return;
  }
  for (int j = 0; j < q; ++j)
    for (int i = 0; i <= j; ++i) {
      double v = mom(midx(fz(i), fz(j)));
      if (a.method == PDSB_METHOD_LSTSQ && i == j && i < p && a.l2 > 0.0) v += a.l2;   // ridge; CD scales l2 by n itself
      AA(i, j) = v; AA(j, i) = v;
    }
  for (int i = 0; i < q; ++i) BB(i) = mom(midx(fz(i), p));
  if (a.method == PDSB_METHOD_CD || a.method == PDSB_METHOD_NNLS) {
    // Iterative solvers on the group's Gram, one thread per group (same recurrences as the single-problem kernel,
    // k3_solve.cu; faer_coordinate_descent lr_solvers.rs:426-538, faer_nn_lr :542-600).  The ridge term a.l2 was added to
    // the diagonal above only for LSTSQ callers: CD scales its penalties by the row count itself.
    for (int i = 0; i < q; ++i) CN(i) = 0.0;                      // beta
    if (a.method == PDSB_METHOD_CD) {
      const double mcount = (double)nrows, lambda_l1 = mcount * a.l1, l2n = mcount * a.l2;
      const double y_sum = mom(midx(p, p + 1));
      for (int it = 0; it < a.max_iter; ++it) {
        double max_change = 0.0;
        for (int j = 0; j < p; ++j) {
          const double before = CN(j);
          double part = 0.0;
          for (int i = 0; i < q; ++i) if (i != j) part += AA(i, j) * CN(i);
          const double mu = BB(j) - part;
          double after;
          if (a.positive && mu < 0.0) after = 0.0;
          else {
            const double sgn = (mu > 0.0) ? 1.0 : ((mu < 0.0) ? -1.0 : 0.0);
            after = sgn * fmax(fabs(mu) - lambda_l1, 0.0) / (AA(j, j) + l2n);
          }
          CN(j) = after;
          max_change = fmax(max_change, fabs(after - before));
        }
        if (a.add_bias) {
          double part = 0.0;
          for (int j = 0; j < p; ++j) part += CN(j) * mom(midx(j, p + 1));
          CN(p) = (y_sum - part) / mcount;
        }
        if (max_change < a.cd_tol) break;
      }
    } else {
      // mu = G beta - X'y lives in the b slots (negated X'y to start with)
      for (int i = 0; i < q; ++i) BB(i) = -BB(i);
      for (int it = 0; it < a.max_iter; ++it) {
        bool ok = true;
        for (int i = 0; i < q; ++i) {
          const double mu = BB(i);
          if (!(mu >= -a.cd_tol)) ok = false;
          if (CN(i) > 0.0 && !(mu <= a.cd_tol)) ok = false;
        }
        if (ok) break;
        for (int k = 0; k < q; ++k) {
          const double bk = CN(k);
          double upd = bk - BB(k) / AA(k, k);
          if (!a.add_bias || k < q - 1) upd = fmax(upd, 0.0);
          const double diff = upd - bk;
          for (int i = 0; i < q; ++i) BB(i) += diff * AA(i, k);
          CN(k) = upd;
        }
      }
    }
    for (int i = 0; i < q; ++i) out[i] = CN(i);
    a.status[g] = PDSB_OK;
    return;
  }
  const bool gated = a.tol > 0.0;
  double ln_den = 0.0;
  if (gated) {
    for (int i = 0; i < q; ++i) {
      double d = AA(i, i);
      if (d <= 0.0) { a.status[g] = PDSB_GATED; for (int k = 0; k < q; ++k) out[k] = nan(""); return; }
      ln_den += log(d);
    }
    // NaN diagonal: the reference's comparisons are all false -> it solves and returns NaN coefficients (k3_solve.cu)
    if (isnan(ln_den)) { a.status[g] = PDSB_OK; for (int k = 0; k < q; ++k) out[k] = nan(""); return; }
  }
  const double ln_tol = gated ? log(a.tol) : 0.0;
  int perm[66];
  bool done = false;
  if (a.solver == PDSB_SOLVER_CHOLESKEY) {
    bool ok = true;
    double ln_det = 0.0;
    for (int k = 0; k < q && ok; ++k) {
      double d = AA(k, k);
      if (!(d > 0.0) || !isfinite(d)) { ok = false; break; }
      d = sqrt(d); AA(k, k) = d; ln_det += 2.0 * log(d);
      for (int i = k + 1; i < q; ++i) AA(i, k) = AA(i, k) / d;
      for (int j = k + 1; j < q; ++j) { double ljk = AA(j, k); for (int i = j; i < q; ++i) AA(i, j) -= AA(i, k) * ljk; }
    }
    if (ok) {
      if (gated && ln_det - ln_den <= ln_tol) { a.status[g] = PDSB_GATED; for (int k = 0; k < q; ++k) out[k] = nan(""); return; }
      for (int i = 0; i < q; ++i) { double s = BB(i); for (int j = 0; j < i; ++j) s -= AA(i, j) * BB(j); BB(i) = s / AA(i, i); }
      for (int i = q - 1; i >= 0; --i) { double s = BB(i); for (int j = i + 1; j < q; ++j) s -= AA(j, i) * BB(j); BB(i) = s / AA(i, i); }
      for (int i = 0; i < q; ++i) out[i] = BB(i);
      done = true;
    } else if (gated) {
      a.status[g] = PDSB_GATED; for (int k = 0; k < q; ++k) out[k] = nan(""); return;
    } else {  // restore and fall through to QR (lr_solvers.rs:288-291)
      for (int j = 0; j < q; ++j)
        for (int i = 0; i <= j; ++i) {
          double v = mom(midx(fz(i), fz(j)));
          if (a.method == PDSB_METHOD_LSTSQ && i == j && i < p && a.l2 > 0.0) v += a.l2;   // ridge; CD scales l2 by n itself
          AA(i, j) = v; AA(j, i) = v;
        }
      for (int i = 0; i < q; ++i) BB(i) = mom(midx(fz(i), p));
    }
  }
  if (!done) {
    // Householder QR with column pivoting
    for (int j = 0; j < q; ++j) perm[j] = j;
    double ln_det = 0.0;
    for (int k = 0; k < q; ++k) {
      int piv = k; double best = -1.0;
      for (int j = k; j < q; ++j) {
        double s = 0.0;
        for (int i = k; i < q; ++i) { double v = AA(i, j); s += v * v; }
        CN(j) = s;
        if (s > best) { best = s; piv = j; }
      }
      if (piv != k) {
        for (int i = 0; i < q; ++i) { double t = AA(i, k); AA(i, k) = AA(i, piv); AA(i, piv) = t; }
        int t = perm[k]; perm[k] = perm[piv]; perm[piv] = t;
      }
      double s = CN(piv);
      double normx = sqrt(s), x0 = AA(k, k);
      double alpha = (x0 >= 0.0) ? -normx : normx;
      double v0 = x0 - alpha;
      double vtv = v0 * v0 + (s - x0 * x0);
      double hb = (vtv > 0.0 && isfinite(vtv)) ? 2.0 / vtv : 0.0;
      if (hb != 0.0) {
        for (int j = k + 1; j < q; ++j) {
          double dot = v0 * AA(k, j);
          for (int i = k + 1; i < q; ++i) dot += AA(i, k) * AA(i, j);
          double f = hb * dot;
          AA(k, j) -= f * v0;
          for (int i = k + 1; i < q; ++i) AA(i, j) -= f * AA(i, k);
        }
        double dot = v0 * BB(k);
        for (int i = k + 1; i < q; ++i) dot += AA(i, k) * BB(i);
        double f = hb * dot;
        BB(k) -= f * v0;
        for (int i = k + 1; i < q; ++i) BB(i) -= f * AA(i, k);
      }
      AA(k, k) = alpha;
      ln_det += log(fabs(alpha));
    }
    if (gated && ln_det - ln_den <= ln_tol) {
      a.status[g] = PDSB_GATED; for (int k = 0; k < q; ++k) out[k] = nan(""); return;
    }
    for (int i = q - 1; i >= 0; --i) {
      double s = BB(i);
      for (int j = i + 1; j < q; ++j) s -= AA(i, j) * BB(j);
      double r = AA(i, i);
      BB(i) = (r != 0.0) ? s / r : 0.0;
    }
    for (int i = 0; i < q; ++i) out[perm[i]] = BB(i);
  }
  a.status[g] = PDSB_OK;
#undef AA
#undef BB
#undef CN
}

template <typename T, int P>
int launch_moments_p(const T* X, int64_t ldx, const T* y, const int64_t* offsets, const int64_t* item_start,
                     int64_t n_groups, int64_t n_items, int64_t n, double* part, cudaStream_t s) {
  int64_t warps = n_items;
  int grid = (int)std::min<int64_t>(ceil_div(warps, 8), (int64_t)sm_count() * 16);
  if (grid < 1) grid = 1;
  constexpr int V = 16 / (int)sizeof(T);
  // (a 16-byte-load variant of the register kernel measured 1.33 ms against 0.995 ms on C3 in round 2 — 140 registers, two
  // 16-byte loads per column in flight instead of four 4-byte ones — and was removed)
  // measured (B200, C3): register kernel 0.996 ms (56 % of HBM); staged kernel 1.38 ms scalar rows -> 1.20 ms packed row
  // pairs -> 1.09 ms with the f32 per-item tree sum (51 %): 110 accumulator registers leave one CTA per SM, and the
  // per-item CTA-wide reduction drains the ring every 16 tiles.  The register kernel stays the default; PDSB_K5_STAGED=1
  // selects the staged one (kept as the recorded experiment, covered by the grouped parity tests under that setting).
  static const bool staged_on = [] { const char* e = getenv("PDSB_K5_STAGED"); return e && e[0] == '1'; }();
  // packed row pairs hold 2 NM accumulators per thread: f32 up to 9 features, f64 up to 5 (more would spill)
  constexpr bool st_fits = (sizeof(T) == 4) ? (P <= 9) : (P <= 5);
  bool launched = false;
  if constexpr (st_fits) {
    constexpr size_t st_smem = staged_smem<T, P>();
    const bool st_aligned = (reinterpret_cast<uintptr_t>(X) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0) && (ldx % V == 0);
    if (staged_on && st_aligned && st_smem <= (size_t)220 * 1024 && n_items >= 1) {
      auto k = group_moments_staged_kernel<T, P>;
      PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)st_smem));
      int per_sm = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, ST_THREADS, st_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
      const int sgrid = (int)std::min<int64_t>(n_items, (int64_t)sm_count() * per_sm);
      int64_t* rows = nullptr;
      if (dev_alloc((void**)&rows, (size_t)n_items * 2 * sizeof(int64_t), s)) return 1;
      item_rows_kernel<<<(int)std::min<int64_t>(ceil_div(n_items, 256), 2048), 256, 0, s>>>(offsets, item_start, n_groups, n_items, rows);
      count_launch();
      k<<<sgrid, ST_THREADS, st_smem, s>>>(X, ldx, y, rows, n_items, n, part);
      dev_free(rows, s);
      launched = true;
    }
  }
  if (!launched) {
    // PDSB_K5_MODE: 0 = binary search + fixed stride (round 1), 1 = precomputed rows, 2 = work queue, 3 = both (default;
    // C3 in one call: 0.861 / 0.852 / 0.810 / 0.805 ms, profiles/r02/k5_mode_sweep.txt)
    static const int mode = [] { const char* e = getenv("PDSB_K5_MODE"); return e ? atoi(e) : 3; }();
    int64_t* rows = nullptr;
    unsigned long long* queue = nullptr;
    if (mode & 1) {
      if (dev_alloc((void**)&rows, (size_t)n_items * 2 * sizeof(int64_t), s)) return 1;
      item_rows_kernel<<<(int)std::min<int64_t>(ceil_div(n_items, 256), 2048), 256, 0, s>>>(offsets, item_start, n_groups, n_items, rows);
      count_launch();
    }
    if (mode & 2) {
      if (dev_alloc((void**)&queue, sizeof(unsigned long long), s)) return 1;
      PDSB_CUDA_OK(cudaMemsetAsync(queue, 0, sizeof(unsigned long long), s));
    }
    auto launch = [&](auto kern) {
      int g = grid;
      if (mode & 2) {      // persistent: exactly the resident CTAs
        int occ = 2;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, 0) != cudaSuccess || occ < 1) { (void)cudaGetLastError(); occ = 2; }
        g = (int)std::min<int64_t>(ceil_div(n_items, (int64_t)8), (int64_t)sm_count() * occ);
        if (g < 1) g = 1;
      }
      kern<<<g, 256, 0, s>>>(X, ldx, y, offsets, item_start, n_groups, n_items, part, rows, queue);
    };
    switch (mode & 3) {
      case 1: launch(group_moments_kernel<T, P, true, 1>); break;
      case 2: launch(group_moments_kernel<T, P, true, 2>); break;
      case 3: launch(group_moments_kernel<T, P, true, 3>); break;
      default: launch(group_moments_kernel<T, P, true, 0>); break;
    }
    if (rows) dev_free(rows, s);
    if (queue) dev_free(queue, s);
  }
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace

// Moments [X | y | 1]' [X | y | 1] of a whole frame with few features on the register kernel above.  The tensor-core
// kernel pays ~800 cycles per 128-row stage whatever the stage holds (p = 8: 1.6 TB/s), while 45..105 FMAs per row fit
// under the HBM time on the FP32 pipes: measured (B200, 2e8 x 9 f32) 3.7 TB/s here.  f32 chains of CHUNK / 32 rows per
// lane, f64 across lanes and items, fixed order -> reproducible.  Returns -1 when the shape is not taken.
template <typename T>
int moments_small(const T* X, int64_t ldx, const T* y, int64_t n, int p, double* M, cudaStream_t s) {
  if (p < 1 || p > 10 || n < 1) return -1;
  const int q1 = p + 2, nm = q1 * (q1 + 1) / 2;
  const int64_t n_items = ceil_div(n, (int64_t)CHUNK);
  int64_t* meta = nullptr;
  double* part = nullptr;
  if (dev_alloc((void**)&meta, 4 * sizeof(int64_t), s)) return 1;
  if (dev_alloc((void**)&part, (size_t)nm * n_items * sizeof(double), s)) { dev_free(meta, s); return 1; }
  small_items_kernel<<<1, 32, 0, s>>>(n, n_items, meta, meta + 2);
  count_launch();
  int grid = (int)std::min<int64_t>(ceil_div(n_items, (int64_t)8), (int64_t)sm_count() * 16);
  if (grid < 1) grid = 1;
  int rc = 0;
#define CASE_P(PP) case PP: group_moments_kernel<T, PP, false><<<grid, 256, 0, s>>>(X, ldx, y, meta, meta + 2, 1, n_items, part); break;
  switch (p) {
    CASE_P(1) CASE_P(2) CASE_P(3) CASE_P(4) CASE_P(5) CASE_P(6) CASE_P(7) CASE_P(8) CASE_P(9) CASE_P(10)
    default: rc = -1;
  }
#undef CASE_P
  if (!rc) {
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) { set_error("small moments launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  if (!rc) {
    small_reduce_kernel<<<nm, 128, 0, s>>>(part, n_items, q1, M);
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) { set_error("small moments reduce launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  dev_free(part, s);
  dev_free(meta, s);
  return rc;
}
template int moments_small<float>(const float*, int64_t, const float*, int64_t, int, double*, cudaStream_t);

template <typename T>
int grouped_lin_reg(const T* X, int64_t ldx, const T* y, const int64_t* offsets, int64_t n_groups, int64_t n,
                    int p, const pdsb_solve_opts& o, double* beta, int* status, cudaStream_t s) {
  if (n_groups <= 0) return 0;
  if (o.method != PDSB_METHOD_LSTSQ && o.method != PDSB_METHOD_CD && o.method != PDSB_METHOD_NNLS) {
    set_error("grouped lin_reg: method %d is not batched", o.method);
    return 1;
  }
  const int q = p + (o.add_bias ? 1 : 0);
  if (p < 1 || p > 64) { set_error("grouped lin_reg: p=%d not supported (1..64)", p); return 1; }
  const int q1 = p + 2, nm = q1 * (q1 + 1) / 2;
  // work list: we need n_items on the host to size the partial buffer -> one small D2H
  int64_t* item_start = nullptr;
  if (dev_alloc((void**)&item_start, (size_t)(n_groups + 1) * sizeof(int64_t), s)) return 1;
  count_items_kernel<<<(int)std::min<int64_t>(ceil_div(n_groups, 256), 1024), 256, 0, s>>>(offsets, n_groups, chunk_rows(), item_start);
  count_launch();
  scan_items_kernel<<<1, 1024, 0, s>>>(item_start, n_groups);
  count_launch();
  int64_t n_items = 0;
  PDSB_CUDA_OK(cudaMemcpyAsync(&n_items, item_start + n_groups, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  PDSB_CUDA_OK(cudaStreamSynchronize(s));
  double* part = nullptr;
  if (dev_alloc((void**)&part, (size_t)nm * n_items * sizeof(double), s)) { dev_free(item_start, s); return 1; }
  int rc = 0;
#define CASE_P(PP) case PP: rc = launch_moments_p<T, PP>(X, ldx, y, offsets, item_start, n_groups, n_items, n, part, s); break;
  switch (p) {
    CASE_P(1) CASE_P(2) CASE_P(3) CASE_P(4) CASE_P(5) CASE_P(6) CASE_P(7) CASE_P(8) CASE_P(9) CASE_P(10)
    default: {
      size_t smem = (size_t)4 * 32 * (q1 + 1) * sizeof(T);
      int grid = (int)std::min<int64_t>(ceil_div(n_items, 4), (int64_t)sm_count() * 16);
      if (grid < 1) grid = 1;
      group_moments_generic_kernel<T><<<grid, 128, smem, s>>>(X, ldx, y, offsets, item_start, n_groups, n_items, p, part);
      cudaError_t e = cudaGetLastError();
      count_launch();
      if (e != cudaSuccess) { set_error("group moments launch failed: %s", cudaGetErrorString(e)); rc = 1; }
    }
  }
#undef CASE_P
  double* ws = nullptr;
  const size_t ws_per_thread = ((size_t)q * q + 2 * q + (size_t)(p + 2) * (p + 3) / 2) * sizeof(double);   // A, b, scratch, moments
  const bool ws_shared = ws_per_thread * 32 <= (size_t)200 * 1024;
  if (!rc && !ws_shared && dev_alloc((void**)&ws, ((size_t)q * q + 2 * q) * sizeof(double) * n_groups, s)) rc = 1;
  if (!rc) {
    GroupSolveArgs a;
    a.part = part; a.item_start = item_start; a.offsets = offsets; a.n_groups = n_groups; a.n_items = n_items;
    a.p = p; a.add_bias = o.add_bias; a.solver = o.solver; a.l2 = o.l2_reg; a.tol = o.singular_x_tol;
    a.method = o.method; a.positive = o.positive; a.max_iter = o.max_iter; a.l1 = o.l1_reg; a.cd_tol = o.tol;
    a.ws = ws; a.beta = beta; a.status = status;
    if (ws_shared) {
      const size_t smem = ws_per_thread * 32;
      if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(group_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      group_solve_kernel<<<(int)ceil_div(n_groups, 32), 32, smem, s>>>(a);
    } else {
      group_solve_kernel<<<(int)ceil_div(n_groups, 128), 128, 0, s>>>(a);
    }
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) { set_error("group solve launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  if (ws) dev_free(ws, s);
  dev_free(part, s);
  dev_free(item_start, s);
  (void)n;
  return rc;
}

template int grouped_lin_reg<float>(const float*, int64_t, const float*, const int64_t*, int64_t, int64_t, int,
                                    const pdsb_solve_opts&, double*, int*, cudaStream_t);
template int grouped_lin_reg<double>(const double*, int64_t, const double*, const int64_t*, int64_t, int64_t, int,
                                     const pdsb_solve_opts&, double*, int*, cudaStream_t);

}  // namespace pdsb
