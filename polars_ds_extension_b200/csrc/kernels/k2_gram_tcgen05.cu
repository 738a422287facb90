// K2b — the headline kernel: moments  M = [X|Y|1]^T [X|Y|1]  for f32 frames on the 5th-gen tensor cores.
//
// Replaces faer's matmul in get_xtx_with_lambda / build_xty (/root/reference/src/linear/lr/lr_solvers.rs:183-211,
// 262-278) and the column sums of faer_coordinate_descent (:483-484): one pass over the frame instead of three.
//
// Shape of the problem: X has p <= 64 columns and n ~ 1e8 rows, i.e. a GEMM with M = N = p and K = n — tiny output,
// enormous reduction.  At p = 32 the FP32 SIMT pipes cannot keep up with HBM (595 FMA per 132-byte row), so X'X goes
// to tcgen05.mma kind::tf32 — and because one TF32 product loses 13 mantissa bits it is computed as the classic
// 3-term split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared, lo = x - hi, exact in fp32):
//        X'X = HH + LH + LH^T (+ LL ~ 2^-22, dropped),     HH = hi^T hi,  LH = lo^T hi.
// Both products come out of ONE instruction stream by stacking hi and lo along the MMA's M = 128 dimension:
//        A  (TMEM, 128 lanes x 8 columns per MMA): hi and lo of the feature columns (layout below)
//        B  (SMEM, N x 8, K-major, 128B swizzle)  : the feature rows of the TMA tile as they landed, N = p padded to 16
// The hi operand is the RAW data: the tensor core reads an fp32 operand as TF32 by ignoring the low 13 mantissa bits
// (measured on B200: tests/test_gpu_moments.py compares against exact-product f64-accumulated moments; assuming
// round-to-nearest instead gives 7e-4 relative error), so B needs no conversion at all and the hi lanes of A are a copy.
//
// Only the FEATURES go through the tensor core.  Round 1 sent Z~ = [X | y | 1] (34 columns at p = 32 -> N = 48, three
// TMEM quadrants); the timeline of that kernel (profiles/r02/k2b_trace_16p16_p32.json) shows the MMA warp as the serial
// bottleneck — 550 of 965 cycles per 128-row stage blocked in tcgen05.mma issue, i.e. the tensor pipe's per-instruction
// time at N = 48 — with the TMEM stores of the converters and the FP64 drain of a 128 x 48 accumulator competing for
// the same TMEM.  With A = X and B = X:  N = 32 (a third less tensor time per MMA), two quadrants instead of three
// (a third fewer TMEM stores), and a 64 x 32 accumulator (a third of the FP64 drain).  What the targets and the ones
// column contribute is O(p t) per row and runs on the CUDA cores of warps that hold the data anyway:
//        X'y_j, sum x      : the converter lanes (a lane owns one feature column and has its 32 rows in registers;
//                            y_j arrives by broadcast loads), packed f32 FMAs per stage -> f64
//        sum y, y_j.y_k, n : 8 side lanes per target on an otherwise idle warp, f32 per box -> f64
//
// A-lane layout ("16 + 16"): a warp can only write the 32 TMEM lanes of its own quadrant (warp id mod 4).  Quadrant k
// holds feature columns 16k .. 16k+15 TWICE: lanes 0..15 = hi (raw), lanes 16..31 = lo.  Lane L and lane L + 16 read
// the SAME shared-memory address (a broadcast), so one LDS.128 of a converter warp touches 16 rows x 16 bytes = 2
// wavefronts and every element of the tile is read from shared memory once.  Quadrants without columns do nothing.
//
// Data flow per CTA (persistent, one CTA per SM, 128-row stages dealt round-robin):
//   warp 0             TMA producer  : cp.async.bulk.tensor, per 32-row box one {32 x p} load of X and one {32 x t} of Y
//   warp 1             MMA issuer    : 16 x tcgen05.mma (K = 8) per stage, fp32 accumulators in TMEM (double-buffered)
//   warps 2..9         converters    : 2 sets x 4 quadrant warps; tile row -> registers -> x - (x & mask) -> tcgen05.st
//   warps 10..17       epilogue      : every 256 rows the accumulator is drained with tcgen05.ld into f64 registers
//                                      (fp32 accumulation inside the tensor core rounds toward zero: bias 1.6e-6)
// A second tiny kernel sums the per-CTA partials in a fixed order (bit-reproducible) and applies the symmetrisation.
// Roofline: HBM-bound, algorithmic bytes = 4 (p + t) per row (+4 with a mask).
#include "../common.h"
#include "kernels.h"
#include <cuda.h>
#include <cstdlib>

namespace pdsb {

namespace {

constexpr int BOX_ROWS = 32;            // K extent of one TMA box = 128 bytes of f32 = one swizzle row
constexpr int BPS = 4;                  // boxes per pipeline stage  (stage = 128 rows)
constexpr int STAGE_ROWS = BOX_ROWS * BPS;
constexpr int MAX_RING = 10;            // TMA landing ring (stages)
constexpr int AB = 3;                   // TMEM A ring (slots of 128 columns): 2 x 64 accumulator + 3 x 128 = 512
constexpr int D_COLS = 64;              // TMEM columns per accumulator buffer (N <= 64)
constexpr int A_COL0 = 2 * D_COLS;
constexpr int FLUSH_STAGES = 2;         // accumulate 2 stages = 256 rows in fp32 (RZ accumulation) before draining to f64
constexpr int NCONV = 2;                // converter warp sets, alternating stages
constexpr int EPI_SETS = 2;             // epilogue warp sets, each draining half of the accumulator columns
constexpr int NUM_WARPS = 4 * NCONV + 4 * EPI_SETS + 4;     // warps 16, 17 idle; 18 = MMA issuer, 19 = TMA producer
constexpr int NUM_THREADS = NUM_WARPS * 32;     // 640 (96 registers per thread)
constexpr int TMEM_COLS = 512;
constexpr int A_SLOT_COLS = BPS * BOX_ROWS;
constexpr int Y_ROWS = 8;               // tile rows reserved for the targets (t <= 4; one 8-row swizzle group)
constexpr int MAX_T = 4;
constexpr uint32_t HI_MASK = 0xFFFFE000u;   // TF32 keeps 10 mantissa bits
constexpr int SIDE_SLOTS = 4;             // partial slots per CTA for the side sums (4 side warps, or the 2 converter sets; zero-filled)
constexpr int XSIDE_COLS = 64;          // doubles per (CTA, converter set, slot): slot j < T = x . y_j, slot T = sum x
constexpr int YSIDE_STRIDE = 32;        // doubles per (CTA, converter set): [3u+0] sum y_u, [3u+1] sum y_u^2, [2] count, [12 + 4j + k] y_j.y_k

template <int NB>
struct Shape {
  static constexpr int N = NB * 16;                                  // MMA N = feature rows of the tile, padded to 16
  static constexpr int NH = N / EPI_SETS;                            // accumulator columns per epilogue set
  static constexpr int TILE_ROWS = N + Y_ROWS;                       // feature rows, then the target rows
  static constexpr uint32_t TILE_BYTES = TILE_ROWS * 128;            // one box-tile
  // ring depth: as many stages as fit in ~200 KB (the kernel is latency-bound on bytes in flight per SM)
  static constexpr int RING = (NB == 1) ? 10 : (NB == 2 ? 10 : (NB == 3 ? 7 : 5));
  static constexpr size_t SMEM = (size_t)RING * BPS * TILE_BYTES;
};

// ---------------------------------------------------------------- PTX helpers ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  // bounded wait (4 s of %globaltimer): a protocol bug must surface as a trapped kernel (an error the host reports),
  // never as a hung GPU.
  // First a plain non-blocking test (the phase has usually flipped long ago for the warp that is the bottleneck);
  // then try_wait with a suspend-time hint, which lets the hardware park the warp until the phase flips instead of
  // re-issuing the poll: ncu counted ~195 barrier polls per 128-row stage without it, all of them wavefronts on the
  // shared-memory pipe.
  uint32_t done = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  if (done) return;
  uint64_t t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!done) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    if (!done) {
      uint64_t t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();          // 4 s: no legitimate wait of this kernel comes near it
    }
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::tf32, M = 128
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (SM100 descriptor v1)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address   bits [0,14)
  d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // layout: SWIZZLE_128B
  return d;
}

struct alignas(8) Barriers {
  uint64_t raw_full[MAX_RING], raw_empty[MAX_RING];   // TMA landed / tile free again (MMA done with B + side lanes done)
  uint64_t a_full[AB], a_empty[AB];                   // TMEM A slot written / consumed
  uint64_t d_full[2], d_empty[2];                     // accumulator buffer ready to drain / drained
  uint32_t tmem_base;
};

// feature column c lives in TMEM lanes hi_lane(c) (raw) and hi_lane(c) + 16 (lo)
__host__ __device__ __forceinline__ int hi_lane(int c) { return (c >> 4) * 32 + (c & 15); }

#ifdef PDSB_TC_ABLATION
// timeline of CTA 0 (ablation builds): clock64 stamps, [stage][event]
constexpr int TRACE_STAGES = 2048, TRACE_EVENTS = 16;
__device__ unsigned long long g_trace[TRACE_STAGES * TRACE_EVENTS];
#define PDSB_TRACE(stage, ev)                                                                                   \
  do {                                                                                                          \
    if (DBG == 16 && blockIdx.x == 0 && lane == 0 && (stage) < (uint32_t)TRACE_STAGES)                          \
      g_trace[(stage) * TRACE_EVENTS + (ev)] = (unsigned long long)clock64();                                   \
  } while (0)
#else
#define PDSB_TRACE(stage, ev) do {} while (0)
#endif

struct GramArgs {
  int64_t n, stages_total;
  int p, t;
  int xcol, ycol;        // row-blocked frame: first feature / target column of the frame
  int blocked;           // row-blocked frame (3-D tensor maps) or column-major matrices (2-D)
  int explicit_hi;       // cross-check: the hi lanes clear the low 13 bits themselves
  int joined;            // Y follows X in memory and p is a multiple of 16: ONE box {32 x (p + t)} per load lands Y at tile row N
};

// T = targets the side lanes are compiled for (1, or 4 for t = 2..4).  DBG = timing ablations (never in production:
// results are garbage): 1 no TMEM store, 2 no lo arithmetic, 4 no shared-memory loads, 8 no MMA, 16 timeline trace.
template <int NB, int T, int DBG>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_y,
                    const float* __restrict__ mask, const GramArgs g, double* __restrict__ partials /* [grid][128][N] */,
                    double* __restrict__ xside /* [grid][SIDE_SLOTS][T+1][64] */, double* __restrict__ yside /* [grid][SIDE_SLOTS][32] */  /* both zero-filled */) {
  using S = Shape<NB>;
  constexpr int N = S::N, NH = S::NH, RING = S::RING;
  constexpr uint32_t TILE_BYTES = S::TILE_BYTES;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;                              // RING * BPS * TILE_BYTES
  Barriers* bars = reinterpret_cast<Barriers*>(raw + S::SMEM);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = g.p, t = g.t;
  const int nact = (p + 15) >> 4;                         // TMEM quadrants that hold feature columns
  // stages are dealt round-robin: at any moment the CTAs stream ADJACENT rows of every column (DRAM page locality:
  // with one contiguous range per CTA the chip ran 148 x q far-apart 128-byte streams and topped out at 4.3 TB/s
  // even with all arithmetic removed)
  const uint32_t my_stages = g.stages_total > (int64_t)blockIdx.x
                                 ? (uint32_t)((g.stages_total - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
  auto stage_row0 = [&](uint32_t it) { return ((int64_t)it * gridDim.x + blockIdx.x) * STAGE_ROWS; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&bars->raw_full[i], 1); mbar_init(&bars->raw_empty[i], NB <= 2 ? 1 + BPS : 1); }   // the MMA's commit (+ the side warps, which read the tile on their own)
    for (int i = 0; i < AB; ++i) { mbar_init(&bars->a_full[i], nact); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], nact * EPI_SETS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows the TMA never writes (feature rows p .. N-1 feed accumulator columns nobody reads, target rows t .. 7 feed
  // side lanes whose results nobody reads): zero them once so that nothing uninitialised is ever multiplied
  for (int i = threadIdx.x; i < RING * BPS * (S::TILE_ROWS - p - t) * 8; i += NUM_THREADS) {
    const int per = (S::TILE_ROWS - p - t) * 8;
    const int tile = i / per, rem = i % per;
    int r = rem / 8;
    const int c = rem % 8;
    r = (r < N - p) ? p + r : N + t + (r - (N - p));
    *reinterpret_cast<uint4*>(raw + (size_t)tile * TILE_BYTES + (size_t)r * 128 + c * 16) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  // Roles.  A warp runs on scheduler (warp id & 3) and may only touch TMEM quadrant (warp id & 3), so converters and
  // epilogue warps come two per residue (warps 0..15); warps 18 / 19 are the MMA issuer and the TMA producer (the MMA
  // warp's serial loop is the kernel's critical path: it should not wait for issue slots behind converters).
  enum { R_TMA, R_MMA, R_CONV, R_SIDE, R_EPI, R_NONE };
  const int wr = warp & 3, wk = warp >> 2;
  // x . y_j / sum x / y sums: up to 32 features the four converter warps of quadrants 2 and 3 hold no A lanes and take one
  // box of every stage each (the two busy converter warps per set stay at ~57 instructions per box: with the sums inside
  // them the p = 32 stage took 2400 instead of 650 cycles, profiles/r02).  Above 32 features every converter warp is busy
  // and has twice the time per stage: there the sums ride on the values the converter lanes hold anyway.
  constexpr bool SIDE_WARPS = (NB <= 2);
  int role = R_NONE, sw_id = 0;
  if (wk == 4) {
    // the two serial warps sit on the schedulers of quadrants 2 / 3, which hold no A lanes up to 32 features
    if (wr == 3) role = R_TMA;
    else if (wr == 2) role = R_MMA;
  } else if (wk < NCONV) {
    if (wr < nact) role = R_CONV;
    else if (wk == 0) {   // (side warps included: they zero their quadrant first)
      // lanes nobody feeds: zero them once so the MMA never multiplies uninitialised TMEM (their accumulator rows are
      // not read either way)
      const uint32_t lane_addr = (uint32_t)(wr * 32) << 16;
      uint32_t z[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) z[k] = 0u;
      for (int c = 0; c < AB * BPS; ++c) tmem_st32(tmem + lane_addr + (uint32_t)(A_COL0 + c * BOX_ROWS), z);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    if (SIDE_WARPS && wr >= 2) { role = R_SIDE; sw_id = wk * 2 + (wr - 2); }
  } else {
    role = R_EPI;
  }

  if (role == R_TMA) {
    // =============================== TMA producer (warp-uniform loop, one elected lane issues) ===============
    uint32_t rs = 0, ph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      mbar_wait(&bars->raw_empty[rs], ph ^ 1);
      PDSB_TRACE(it, 0);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->raw_full[rs], (uint32_t)(BPS * (p + t) * 128));
        const int64_t row0 = stage_row0(it);
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
          unsigned char* dst = raw + ((size_t)rs * BPS + b) * TILE_BYTES;
          if (g.blocked) {
            tma_load_3d(dst, &tmap_x, &bars->raw_full[rs], b * BOX_ROWS, g.xcol, (int)(row0 / STAGE_ROWS));
            if (!g.joined) tma_load_3d(dst + N * 128, &tmap_y, &bars->raw_full[rs], b * BOX_ROWS, g.ycol, (int)(row0 / STAGE_ROWS));
          } else {
            tma_load_2d(dst, &tmap_x, &bars->raw_full[rs], (int)(row0 + b * BOX_ROWS), 0);
            if (!g.joined) tma_load_2d(dst + N * 128, &tmap_y, &bars->raw_full[rs], (int)(row0 + b * BOX_ROWS), 0);
          }
        }
      }
      __syncwarp();
      PDSB_TRACE(it, 14);
      if (++rs == RING) { rs = 0; ph ^= 1; }
    }
  } else if (role == R_MMA) {
    // =============================== MMA issuer (warp-uniform loop, one elected lane issues) ===============
    // instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t raw_addr = smem_u32(raw);
    // One loop trip = one flush group (FLUSH_STAGES stages into one accumulator buffer), unrolled: the timeline of the
    // per-stage loop (profiles/r02) showed ~400 of its ~800 cycles per stage outside tcgen05.mma issue — every barrier
    // test, fence and commit of this warp queues behind the converters' shared-memory traffic — so the bookkeeping is
    // paid once per group and the a_full test of the second stage runs while the first stage's MMAs execute.  (Testing
    // both barriers first and issuing the group's 32 MMAs back to back measured 2 % slower: profiles/r02/k2b notes.)
    uint32_t s = 0, ph = 0, rs = 0, buf = 0, dph = 0;   // A slot / phase, ring slot, D buffer / phase
    for (uint32_t it0 = 0; it0 < my_stages; it0 += FLUSH_STAGES) {
      mbar_wait(&bars->d_empty[buf], dph ^ 1);
      PDSB_TRACE(it0, 1);
      const uint32_t d_addr = tmem + buf * D_COLS;
#pragma unroll
      for (int fl = 0; fl < FLUSH_STAGES; ++fl) {
        const uint32_t it = it0 + fl;
        if (it < my_stages) {                 // warp-uniform
          mbar_wait(&bars->a_full[s], ph);    // converters only signal after raw_full: B (the tile) has landed too
          PDSB_TRACE(it, 11);
          tc_fence_after();
          PDSB_TRACE(it, 6);
          if (elect_one()) {
            const uint32_t a_base = tmem + A_COL0 + s * A_SLOT_COLS;
            const uint64_t bd0 = make_b_desc(raw_addr + rs * (BPS * TILE_BYTES));
            if (!(DBG & 8)) {
#pragma unroll
              for (int b = 0; b < BPS; ++b) {
#pragma unroll
                for (int k = 0; k < BOX_ROWS / 8; ++k) {
                  // descriptor start address advances in 16-byte units: +TILE_BYTES per box, +32 bytes per K = 8 step
                  const uint64_t bd = bd0 + (uint64_t)((b * TILE_BYTES + k * 32) >> 4);
                  tc_mma_tf32_ts(d_addr, a_base + b * BOX_ROWS + k * 8, bd, idesc, (fl == 0 && b == 0 && k == 0) ? 0u : 1u);
                }
              }
            }
            tc_commit(&bars->a_empty[s]);       // TMEM A slot reusable
            tc_commit(&bars->raw_empty[rs]);    // tile (B operand) reusable
            if (fl == FLUSH_STAGES - 1 || it == my_stages - 1) tc_commit(&bars->d_full[buf]);
          }
          __syncwarp();
          PDSB_TRACE(it, 7);
          if (++s == AB) { s = 0; ph ^= 1; }
          if (++rs == RING) rs = 0;
        }
      }
      if (buf) dph ^= 1;
      buf ^= 1;
    }
  } else if (role == R_CONV) {
    // =============================== converters: NCONV sets x 4 quadrant warps; set j owns stages it = j (mod NCONV) ===
    // Per 32-row box a lane loads the 32 rows of its feature column (8 x LDS.128), then
    //   * adds x . y_j and sum x of those rows to packed f32 accumulators: the registers are NATURAL pairs (rows 2k, 2k+1),
    //     the target arrives by broadcast loads in the same pairs, so this is one FFMA2 + one FADD2 per two rows and no
    //     second pass over the tile (a separate side warp re-reading x cost 160 shared-memory wavefronts per stage and
    //     ran at 1000 cycles per stage, profiles/r02);  lanes 16..31 hold the same x as lanes 0..15 and repeat the sums,
    //     which costs nothing extra per warp instruction;
    //   * turns the values into hi / lo and stores them to TMEM.
    // The quadrant-0 warp also keeps sum y, y_j . y_k and the masked row count (8 lanes per target).
    const int quad = wr;                       // TMEM lane quadrant this warp may touch
    const uint32_t set = (uint32_t)wk;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int m = quad * 16 + (lane & 15);     // feature column of this lane (lanes L and L + 16 share it)
    // padding lanes feed accumulator rows nobody reads: they load the same address as a real lane (a broadcast)
    const int trow = m < p ? m : 0;
    const uint32_t sw = (uint32_t)(trow & 7);
    // hi lanes keep the raw value (x - 0), lo lanes hold x - (x & HI_MASK): one LOP3 per element, one FADD2 per two
    // (cross-check build: the hi lanes subtract their own low 13 bits, i.e. hold trunc(x) explicitly)
    const uint32_t sub_mask = (lane & 16) ? HI_MASK : (g.explicit_hi ? ~HI_MASK : 0u);
    // f32 partial sums live for SIDE_FLUSH of this warp's stages (chains of at most 64 terms, round-to-nearest) before
    // they are added to the f64 totals: an FP64 instruction costs a warp 20-30 cycles here
    constexpr uint32_t SIDE_FLUSH = 4;
    float2 axy[T][2], asx[2];
#pragma unroll
    for (int j = 0; j < T; ++j) axy[j][0] = axy[j][1] = make_float2(0.0f, 0.0f);
    asx[0] = asx[1] = make_float2(0.0f, 0.0f);
    double dxs[T + 1];                                          // f64: x . y_j (j < T), sum x
#pragma unroll
    for (int j = 0; j <= T; ++j) dxs[j] = 0.0;
    float sy = 0.0f, syy = 0.0f, sxy[3] = {0.0f, 0.0f, 0.0f};   // quadrant 0, lanes 8j + c: chunk c of target j; sxy[d-1] = y_j . y_{j+d}
    double dsy = 0.0, dsyy = 0.0, dcnt = 0.0, dxy[3] = {0.0, 0.0, 0.0};
    uint32_t nst = 0;
    for (uint32_t it = set; it < my_stages; it += NCONV, ++nst) {
      const uint32_t rs = it % RING, rph = (it / RING) & 1;
      const uint32_t s = it % AB, sph = (it / AB) & 1;
      mbar_wait(&bars->raw_full[rs], rph);
      if (quad == 0) PDSB_TRACE(it, 2);
      mbar_wait(&bars->a_empty[s], sph ^ 1);
      tc_fence_after();
      if (quad == 0) PDSB_TRACE(it, 3);
#pragma unroll
      for (int b = 0; b < BPS; ++b) {
        const unsigned char* tile = raw + ((size_t)rs * BPS + b) * TILE_BYTES;
        uint32_t v[32];
        const unsigned char* rowp = tile + (size_t)trow * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 x = make_uint4(c, c + 1, c + 2, c + 3);
          if (!(DBG & 4)) x = *reinterpret_cast<const uint4*>(rowp + ((c ^ sw) << 4));
          v[4 * c + 0] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        if constexpr (!SIDE_WARPS) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
          uint4 yv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) yv[c] = *reinterpret_cast<const uint4*>(tile + (size_t)(N + j) * 128 + ((c ^ j) << 4));   // same address in every lane
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            axy[j][0] = __ffma2_rn(make_float2(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1])),
                                   make_float2(__uint_as_float(yv[c].x), __uint_as_float(yv[c].y)), axy[j][0]);
            axy[j][1] = __ffma2_rn(make_float2(__uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])),
                                   make_float2(__uint_as_float(yv[c].z), __uint_as_float(yv[c].w)), axy[j][1]);
          }
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          asx[0] = __fadd2_rn(asx[0], make_float2(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1])));
          asx[1] = __fadd2_rn(asx[1], make_float2(__uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])));
        }
        }
        if (!(DBG & 2)) {
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            // -(x & mask) as one LOP3: (x & mask) ^ sign   (hi lanes: x + (-0) = x; lo lanes: exact in fp32)
            const float2 r = __fadd2_rn(make_float2(__uint_as_float(v[k]), __uint_as_float(v[k + 1])),
                                        make_float2(__uint_as_float((v[k] & sub_mask) ^ 0x80000000u),
                                                    __uint_as_float((v[k + 1] & sub_mask) ^ 0x80000000u)));
            v[k] = __float_as_uint(r.x); v[k + 1] = __float_as_uint(r.y);
          }
        }
        if (!(DBG & 1)) tmem_st32(tmem + lane_addr + (uint32_t)(A_COL0 + s * A_SLOT_COLS + b * BOX_ROWS), v);
        else if (v[0] == 0x7fc12345u && v[31] == 0x12345u) bars->tmem_base = v[5];   // ablation build: keep v alive
        if (!SIDE_WARPS && quad == 0) {               // warp-uniform
          if (lane < 8 * t) {
            const int j = lane >> 3, c = lane & 7;
            const uint4 yv = *reinterpret_cast<const uint4*>(tile + (size_t)(N + j) * 128 + ((c ^ j) << 4));
            const float y0 = __uint_as_float(yv.x), y1 = __uint_as_float(yv.y), y2 = __uint_as_float(yv.z), y3 = __uint_as_float(yv.w);
            sy += (y0 + y1) + (y2 + y3);
            syy = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, syy))));
            if (T > 1) {
#pragma unroll
              for (int d = 1; d < 4; ++d)
                if (j + d < t) {
                  const int kr = j + d;
                  const uint4 kv = *reinterpret_cast<const uint4*>(tile + (size_t)(N + kr) * 128 + ((c ^ kr) << 4));
                  sxy[d - 1] = fmaf(y0, __uint_as_float(kv.x), fmaf(y1, __uint_as_float(kv.y),
                               fmaf(y2, __uint_as_float(kv.z), fmaf(y3, __uint_as_float(kv.w), sxy[d - 1]))));
                }
            }
          }
          if (mask != nullptr && lane < 8) {
            // row count of a masked frame: the packer zeroes masked rows, so only the count needs the mask
            const int64_t r0 = stage_row0(it) + b * BOX_ROWS + lane * 4;
            float o = 0.0f;
#pragma unroll
            for (int e = 0; e < 4; ++e) if (r0 + e < g.n) o += __ldg(mask + r0 + e);
            dcnt += (double)o;
          }
        }
      }
      if (quad == 0) PDSB_TRACE(it, 4);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_full[s]);
      if (quad == 0) PDSB_TRACE(it, 5);
      if (!SIDE_WARPS && ((nst % SIDE_FLUSH) == SIDE_FLUSH - 1 || it + NCONV >= my_stages)) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
          const float2 a = __fadd2_rn(axy[j][0], axy[j][1]);
          dxs[j] += (double)(a.x + a.y);
          axy[j][0] = axy[j][1] = make_float2(0.0f, 0.0f);
        }
        const float2 a = __fadd2_rn(asx[0], asx[1]);
        dxs[T] += (double)(a.x + a.y);
        asx[0] = asx[1] = make_float2(0.0f, 0.0f);
        if (quad == 0) {
          dsy += (double)sy; dsyy += (double)syy; sy = 0.0f; syy = 0.0f;
          if (T > 1) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { dxy[d] += (double)sxy[d]; sxy[d] = 0.0f; }
          }
        }
      }
    }
    if (!SIDE_WARPS && lane < 16 && m < p) {
      double* xs = xside + ((size_t)blockIdx.x * SIDE_SLOTS + set) * (T + 1) * XSIDE_COLS;
#pragma unroll
      for (int j = 0; j <= T; ++j) xs[j * XSIDE_COLS + m] = dxs[j];
    }
    if (!SIDE_WARPS && quad == 0) {
      // reduce the 8 chunk-lanes of every target (fixed order -> reproducible) and the masked-row count
      double* ys = yside + ((size_t)blockIdx.x * SIDE_SLOTS + set) * YSIDE_STRIDE;
      for (int off = 4; off; off >>= 1) {
        dsy += __shfl_down_sync(0xffffffffu, dsy, off, 8);
        dsyy += __shfl_down_sync(0xffffffffu, dsyy, off, 8);
        dcnt += __shfl_down_sync(0xffffffffu, dcnt, off, 8);
#pragma unroll
        for (int d = 0; d < 3; ++d) dxy[d] += __shfl_down_sync(0xffffffffu, dxy[d], off, 8);
      }
      if ((lane & 7) == 0 && (lane >> 3) < t) {
        const int j = lane >> 3;
        ys[j * 3 + 0] = dsy; ys[j * 3 + 1] = dsyy;
        for (int d = 1; d < 4; ++d) if (j + d < t) ys[12 + j * 4 + (j + d)] = dxy[d - 1];
      }
      if (lane == 0) ys[2] = dcnt;
    }
  } else if (role == R_SIDE) {
    // =============================== side warps (p <= 32): warp k takes box k of EVERY stage, lane = feature ==========
    // x . y_j and sum x over NATURAL register pairs (a 16-byte load gives rows (4c, 4c+1), (4c+2, 4c+3) of the column, the
    // broadcast load of y_j the same rows of the target: one FFMA2 / FADD2 per two rows); lanes 8j + c also hold chunk c of
    // target j (sum y, y_j . y_k); lanes 0..7 count the rows of a masked frame.  All loads of the box are issued before the
    // first FMA; f32 partial sums (chains of at most 64 terms) go to the f64 totals every SIDE_FLUSH stages.
    if constexpr (SIDE_WARPS) {
    constexpr uint32_t SIDE_FLUSH = 4;
    const int b = sw_id;
    const int frow = lane < p ? lane : 0;
    const uint32_t fsw = (uint32_t)(frow & 7);
    float2 axy[T][2], asx[2];
#pragma unroll
    for (int j = 0; j < T; ++j) axy[j][0] = axy[j][1] = make_float2(0.0f, 0.0f);
    asx[0] = asx[1] = make_float2(0.0f, 0.0f);
    double dxs[T + 1];
#pragma unroll
    for (int j = 0; j <= T; ++j) dxs[j] = 0.0;
    float sy = 0.0f, syy = 0.0f, sxy[3] = {0.0f, 0.0f, 0.0f};
    double dsy = 0.0, dsyy = 0.0, dcnt = 0.0, dxy[3] = {0.0, 0.0, 0.0};
    uint32_t rs = 0, rph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      if (sw_id == 0) PDSB_TRACE(it, 15);
      mbar_wait(&bars->raw_full[rs], rph);
      if (sw_id == 0) PDSB_TRACE(it, 12);
      const unsigned char* tile = raw + ((size_t)rs * BPS + b) * TILE_BYTES;
      uint4 xv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) xv[c] = *reinterpret_cast<const uint4*>(tile + (size_t)frow * 128 + ((c ^ fsw) << 4));
#pragma unroll
      for (int j = 0; j < T; ++j) {
        uint4 yv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) yv[c] = *reinterpret_cast<const uint4*>(tile + (size_t)(N + j) * 128 + ((c ^ j) << 4));   // same address in every lane
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          axy[j][0] = __ffma2_rn(make_float2(__uint_as_float(xv[c].x), __uint_as_float(xv[c].y)),
                                 make_float2(__uint_as_float(yv[c].x), __uint_as_float(yv[c].y)), axy[j][0]);
          axy[j][1] = __ffma2_rn(make_float2(__uint_as_float(xv[c].z), __uint_as_float(xv[c].w)),
                                 make_float2(__uint_as_float(yv[c].z), __uint_as_float(yv[c].w)), axy[j][1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        asx[0] = __fadd2_rn(asx[0], make_float2(__uint_as_float(xv[c].x), __uint_as_float(xv[c].y)));
        asx[1] = __fadd2_rn(asx[1], make_float2(__uint_as_float(xv[c].z), __uint_as_float(xv[c].w)));
      }
      if (lane < 8 * t) {
        const int j = lane >> 3, c = lane & 7;
        const uint4 yv = *reinterpret_cast<const uint4*>(tile + (size_t)(N + j) * 128 + ((c ^ j) << 4));
        const float y0 = __uint_as_float(yv.x), y1 = __uint_as_float(yv.y), y2 = __uint_as_float(yv.z), y3 = __uint_as_float(yv.w);
        sy += (y0 + y1) + (y2 + y3);
        syy = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, syy))));
        if (T > 1) {
#pragma unroll
          for (int d = 1; d < 4; ++d)
            if (j + d < t) {
              const int kr = j + d;
              const uint4 kv = *reinterpret_cast<const uint4*>(tile + (size_t)(N + kr) * 128 + ((c ^ kr) << 4));
              sxy[d - 1] = fmaf(y0, __uint_as_float(kv.x), fmaf(y1, __uint_as_float(kv.y),
                           fmaf(y2, __uint_as_float(kv.z), fmaf(y3, __uint_as_float(kv.w), sxy[d - 1]))));
            }
        }
      }
      if (mask != nullptr && lane < 8) {
        // row count of a masked frame: the packer zeroes masked rows, so only the count needs the mask
        const int64_t r0 = stage_row0(it) + b * BOX_ROWS + lane * 4;
        float o = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) if (r0 + e < g.n) o += __ldg(mask + r0 + e);
        dcnt += (double)o;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->raw_empty[rs]);           // the MMA's commit is the other arrival
      if (sw_id == 0) PDSB_TRACE(it, 13);
      if ((it % SIDE_FLUSH) == SIDE_FLUSH - 1 || it == my_stages - 1) {
#pragma unroll
        for (int j = 0; j < T; ++j) {
          const float2 a = __fadd2_rn(axy[j][0], axy[j][1]);
          dxs[j] += (double)(a.x + a.y);
          axy[j][0] = axy[j][1] = make_float2(0.0f, 0.0f);
        }
        const float2 a = __fadd2_rn(asx[0], asx[1]);
        dxs[T] += (double)(a.x + a.y);
        asx[0] = asx[1] = make_float2(0.0f, 0.0f);
        dsy += (double)sy; dsyy += (double)syy; sy = 0.0f; syy = 0.0f;
        if (T > 1) {
#pragma unroll
          for (int d = 0; d < 3; ++d) { dxy[d] += (double)sxy[d]; sxy[d] = 0.0f; }
        }
      }
      if (++rs == RING) { rs = 0; rph ^= 1; }
    }
    if (lane < p) {
      double* xs = xside + ((size_t)blockIdx.x * SIDE_SLOTS + sw_id) * (T + 1) * XSIDE_COLS;
#pragma unroll
      for (int j = 0; j <= T; ++j) xs[j * XSIDE_COLS + lane] = dxs[j];
    }
    double* ys = yside + ((size_t)blockIdx.x * SIDE_SLOTS + sw_id) * YSIDE_STRIDE;
    for (int off = 4; off; off >>= 1) {
      dsy += __shfl_down_sync(0xffffffffu, dsy, off, 8);
      dsyy += __shfl_down_sync(0xffffffffu, dsyy, off, 8);
      dcnt += __shfl_down_sync(0xffffffffu, dcnt, off, 8);
#pragma unroll
      for (int d = 0; d < 3; ++d) dxy[d] += __shfl_down_sync(0xffffffffu, dxy[d], off, 8);
    }
    if ((lane & 7) == 0 && (lane >> 3) < t) {
      const int j = lane >> 3;
      ys[j * 3 + 0] = dsy; ys[j * 3 + 1] = dsyy;
      for (int d = 1; d < 4; ++d) if (j + d < t) ys[12 + j * 4 + (j + d)] = dxy[d - 1];
    }
    if (lane == 0) ys[2] = dcnt;
    }
  } else if (role == R_EPI) {
    // =============================== epilogue: EPI_SETS x 4 warps, set e drains columns [e*NH, (e+1)*NH) =========
    const int quad = wr;
    const int eset = wk - NCONV;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    if (quad < nact) {
      double acc[NH];
#pragma unroll
      for (int j = 0; j < NH; ++j) acc[j] = 0.0;
      const uint32_t groups = (my_stages + FLUSH_STAGES - 1) / FLUSH_STAGES;
      uint32_t buf = 0, dph = 0;
      for (uint32_t grp = 0; grp < groups; ++grp) {
        mbar_wait(&bars->d_full[buf], dph);
        tc_fence_after();
        if (quad == 0 && eset == 0) PDSB_TRACE(grp * FLUSH_STAGES + FLUSH_STAGES - 1, 8);
        uint32_t v[NH];
#pragma unroll
        for (int c = 0; c < NH / 8; ++c) tmem_ld8(tmem + lane_addr + (uint32_t)(buf * D_COLS + eset * NH + c * 8), v + 8 * c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->d_empty[buf]);      // the buffer is free before the f64 adds run
        if (quad == 0 && eset == 0) PDSB_TRACE(grp * FLUSH_STAGES + FLUSH_STAGES - 1, 9);
#pragma unroll
        for (int j = 0; j < NH; ++j) acc[j] += (double)__uint_as_float(v[j]);
        if (quad == 0 && eset == 0) PDSB_TRACE(grp * FLUSH_STAGES + FLUSH_STAGES - 1, 10);
        if (buf) dph ^= 1;
        buf ^= 1;
      }
      double* out = partials + ((size_t)blockIdx.x * 128 + (size_t)(quad * 32 + lane)) * N + eset * NH;
#pragma unroll
      for (int j = 0; j < NH; ++j) out[j] = acc[j];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  }
}

// Sum the per-CTA partials in a fixed order (bit-reproducible), X'X[a][b] = HH[a][b] + LH[a][b] + LH[b][a], and lay the
// moments out in the order [X | Y | 1].  One WARP per element of the upper triangle (mirrored -> exactly symmetric): lane
// l sums parts l, l + 32, ..., then a fixed-order xor tree.  (One thread per element walking all 148 parts took 0.31 ms —
// 6 % of the C2 step; this takes a few microseconds.)
__global__ void __launch_bounds__(256)
gram_finalize_kernel(const double* __restrict__ partials, const double* __restrict__ xside, const double* __restrict__ yside,
                     int nparts, int N, int T, int p, int t, int64_t n, int masked, double* __restrict__ M) {
  const int q1 = p + t + 1;
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= q1 * q1) return;
  const int i = idx / q1, j = idx % q1;
  if (i > j) return;
  double r = 0.0;
  if (j < p) {                                     // X'X
    const int li = hi_lane(i), lj = hi_lane(j);
    double hh = 0.0, hh_t = 0.0, lh_ij = 0.0, lh_ji = 0.0;
    for (int k = lane; k < nparts; k += 32) {
      const double* P = partials + (size_t)k * 128 * N;
      hh += P[(size_t)li * N + j];
      hh_t += P[(size_t)lj * N + i];
      lh_ij += P[(size_t)(li + 16) * N + j];
      lh_ji += P[(size_t)(lj + 16) * N + i];
    }
    r = 0.5 * (hh + hh_t) + (lh_ij + lh_ji);      // per lane; the lanes are added below
  } else if (i < p) {                              // X'y_k (slot k) and the column sums (slot T) from the side sums
    const int slot = (j < p + t) ? j - p : T;
    for (int k = lane; k < nparts * SIDE_SLOTS; k += 32) r += xside[((size_t)k * (T + 1) + slot) * XSIDE_COLS + i];
  } else {
    // y / ones block from the side lanes: [3u+0] sum y_u, [3u+1] sum y_u^2, [2] count, [12 + 4a + b] y_a . y_b
    const int a = i - p, bb = j - p;
    int off;
    if (j == p + t) off = (i == p + t) ? 2 : a * 3 + 0;
    else off = (i == j) ? a * 3 + 1 : 12 + a * 4 + bb;
    if (!(j == p + t && i == p + t && !masked))
      for (int k = lane; k < nparts * SIDE_SLOTS; k += 32) r += yside[(size_t)k * YSIDE_STRIDE + off];
  }
  for (int off = 16; off; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
  if (j == p + t && i == p + t && !masked) r = (double)n;
  if (lane == 0) {
    M[(size_t)i * q1 + j] = r;
    M[(size_t)j * q1 + i] = r;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

bool shape_ok(int p, int t) { return p >= 1 && p <= 64 && t >= 1 && t <= MAX_T; }

// 1 (default): raw hi operand;  0: the hi lanes clear the low 13 bits themselves (proves the hardware truncation of A)
std::atomic<int> g_tc_mode{-1};
int tc_mode() {
  int m = g_tc_mode.load();
  if (m < 0) {
    const char* e = getenv("PDSB_TC_MODE");
    m = e ? atoi(e) : 1;
    g_tc_mode.store(m);
  }
  return m;
}

template <int NB, int T, int DBG>
int launch_one(const CUtensorMap& tx, const CUtensorMap& ty, const float* mask, const GramArgs& g, int grid,
               double* partials, double* xside, double* yside, cudaStream_t s) {
  const size_t smem = Shape<NB>::SMEM + sizeof(Barriers) + 256;
  auto k = gram_tcgen05_kernel<NB, T, DBG>;
  PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, NUM_THREADS, smem, s>>>(tx, ty, mask, g, partials, xside, yside);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template <int NB, int T>
int launch_nb(const CUtensorMap& tx, const CUtensorMap& ty, const float* mask, const GramArgs& g, int grid,
              double* partials, double* xside, double* yside, cudaStream_t s) {
#ifdef PDSB_TC_ABLATION
  // timing ablations of the bench shape only (results are garbage): PDSB_TC_DBG = 1 no TMEM store, 2 no lo arithmetic,
  // 4 no shared-memory loads, 8 no MMA, 15 all of them; 16 = full kernel + timeline trace of CTA 0
  static int dbg = [] { const char* e = getenv("PDSB_TC_DBG"); return e ? atoi(e) : 0; }();
  if (NB == 2 && T == 1) {
    switch (dbg) {
      case 1: return launch_one<NB, T, 1>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 2: return launch_one<NB, T, 2>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 4: return launch_one<NB, T, 4>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 8: return launch_one<NB, T, 8>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 7: return launch_one<NB, T, 7>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 15: return launch_one<NB, T, 15>(tx, ty, mask, g, grid, partials, xside, yside, s);
      case 16: return launch_one<NB, T, 16>(tx, ty, mask, g, grid, partials, xside, yside, s);
      default: break;
    }
  }
#endif
  return launch_one<NB, T, 0>(tx, ty, mask, g, grid, partials, xside, yside, s);
}

// geometry of one call: where X and Y lie
struct Geometry {
  const float* xbase; const float* ybase;   // column-major: first feature / target column;  blocked: the frame (both)
  int64_t ldx, ldy;
  int ncols, xcol, ycol;                    // blocked frame
  bool blocked;
};

int encode_maps(const Geometry& g, int64_t n, int p, int t, bool joined, CUtensorMap* tx, CUtensorMap* ty) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  CUresult cr = CUDA_SUCCESS;
  for (int which = 0; which < 2 && cr == CUDA_SUCCESS; ++which) {
    CUtensorMap* tm = which ? ty : tx;
    const int cols = which ? t : (joined ? p + t : p);
    if (g.blocked) {
      // row-blocked frame: [block][column][128 rows] -> every 128-row x ncols stage is ONE contiguous run in HBM.
      // (column-major matrices cap this kernel at 4.3 TB/s even with all arithmetic removed; blocked: 6.5+ TB/s)
      cuuint64_t dims3[3] = {(cuuint64_t)STAGE_ROWS, (cuuint64_t)g.ncols, (cuuint64_t)ceil_div(n, (int64_t)STAGE_ROWS)};
      cuuint64_t strides3[2] = {(cuuint64_t)STAGE_ROWS * sizeof(float), (cuuint64_t)STAGE_ROWS * g.ncols * sizeof(float)};
      cuuint32_t box3[3] = {(cuuint32_t)BOX_ROWS, (cuuint32_t)cols, 1};
      cuuint32_t estr3[3] = {1, 1, 1};
      cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.xbase), dims3, strides3, box3, estr3,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
      const int64_t ld = which ? g.ldy : g.ldx;
      cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)cols};      // joined: X's map also spans the t target columns behind it
      cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
      cuuint32_t box[2] = {(cuuint32_t)BOX_ROWS, (cuuint32_t)cols};
      cuuint32_t estr[2] = {1, 1};
      cr = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(which ? g.ybase : g.xbase), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
  }
  if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return 1; }
  return 0;
}

int moments_tcgen05_core(const Geometry& geo, const float* mask, int64_t n, int p, int t, double* M, cudaStream_t s) {
  // Y right behind X and N == p: one box per load carries both (half the TMA requests; the 128-byte target boxes are as
  // expensive to issue as the feature boxes)
  const bool joined = (p % 16 == 0) && (geo.blocked ? geo.ycol == geo.xcol + p
                                                     : (geo.ybase == geo.xbase + (size_t)p * geo.ldx && geo.ldx == geo.ldy));
  CUtensorMap tx, ty;
  if (int rc = encode_maps(geo, n, p, t, joined, &tx, &ty)) return rc;
  GramArgs a;
  a.n = n; a.stages_total = ceil_div(n, (int64_t)STAGE_ROWS); a.p = p; a.t = t; a.xcol = geo.xcol; a.ycol = geo.ycol;
  a.blocked = geo.blocked ? 1 : 0; a.explicit_hi = (tc_mode() == 0) ? 1 : 0; a.joined = joined ? 1 : 0;
  const int NB = (p + 15) / 16, N = NB * 16, T = (t == 1) ? 1 : MAX_T;
  int grid = sm_count();
  if (a.stages_total < grid) grid = (int)a.stages_total;
  const size_t n_part = (size_t)grid * 128 * N, n_xs = (size_t)grid * SIDE_SLOTS * (T + 1) * XSIDE_COLS, n_ys = (size_t)grid * SIDE_SLOTS * YSIDE_STRIDE;
  double* partials = nullptr;
  if (dev_alloc((void**)&partials, (n_part + n_xs + n_ys) * sizeof(double), s)) return 1;
  double* xside = partials + n_part;
  double* yside = xside + n_xs;
  PDSB_CUDA_OK(cudaMemsetAsync(xside, 0, (n_xs + n_ys) * sizeof(double), s));     // not every slot has a writer
  int rc;
#define PDSB_GO(NBV) (t == 1 ? launch_nb<NBV, 1>(tx, ty, mask, a, grid, partials, xside, yside, s) \
                             : launch_nb<NBV, MAX_T>(tx, ty, mask, a, grid, partials, xside, yside, s))
  switch (NB) {
    case 1: rc = PDSB_GO(1); break;
    case 2: rc = PDSB_GO(2); break;
    case 3: rc = PDSB_GO(3); break;
    default: rc = PDSB_GO(4); break;
  }
#undef PDSB_GO
  if (!rc) {
    const int q1 = p + t + 1;
    gram_finalize_kernel<<<(q1 * q1 + 7) / 8, 256, 0, s>>>(partials, xside, yside, grid, N, T, p, t, n, mask ? 1 : 0, M);   // 8 warps = 8 elements per block
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) { set_error("gram finalize launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  dev_free(partials, s);
  return rc;
}

}  // namespace

void set_tc_mode(int m) { g_tc_mode.store(m); }

#ifdef PDSB_TC_ABLATION
extern "C" int pdsb_debug_tc_trace(unsigned long long* host, int max_stages) {
  const size_t nbytes = sizeof(unsigned long long) * TRACE_EVENTS * (size_t)(max_stages < TRACE_STAGES ? max_stages : TRACE_STAGES);
  return cudaMemcpyFromSymbol(host, g_trace, nbytes) == cudaSuccess ? TRACE_EVENTS : -1;
}
#endif

// column-major matrices: X = p columns with leading dimension ldx, Y = t columns with ldy (anywhere in memory)
bool moments_tcgen05_supported(const float* X, int64_t ldx, const float* Y, int64_t ldy, int64_t n, int p, int t) {
  if (getenv("PDSB_DISABLE_TCGEN05")) return false;
  if (n < 4096) return false;                 // latency-bound sizes stay on the SIMT kernel
  if (n >= (int64_t(1) << 31) - STAGE_ROWS) return false;   // the 2-D tensor maps are addressed with int32 row coordinates
  if (!get_encode_fn() || !shape_ok(p, t)) return false;
  if ((ldx % 4) != 0 || (ldy % 4) != 0) return false;       // global strides of a tensor map are multiples of 16 bytes
  return ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) == 0;
}

int moments_tcgen05_f32(const float* X, int64_t ldx, const float* Y, int64_t ldy, const float* mask, int64_t n, int p,
                        int t, double* M, cudaStream_t s) {
  if (!moments_tcgen05_supported(X, ldx, Y, ldy, n, p, t)) return -1;
  Geometry g{X, Y, ldx, ldy, 0, 0, 0, false};
  return moments_tcgen05_core(g, mask, n, p, t, M, s);
}

// row-blocked frame: [block][column][FRAME_ROWS]; X = columns xcol .. xcol+p-1, Y = columns ycol .. ycol+t-1
bool moments_tcgen05_frame_supported(int64_t n, int ncols, int xcol, int p, int ycol, int t) {
  if (getenv("PDSB_DISABLE_TCGEN05") || !get_encode_fn()) return false;
  if (n < 4096 || !shape_ok(p, t)) return false;
  if (xcol < 0 || ycol < 0 || xcol + p > ncols || ycol + t > ncols) return false;
  return ceil_div(n, (int64_t)STAGE_ROWS) < (int64_t(1) << 31);     // int32 block coordinate of the 3-D tensor maps
}

int moments_tcgen05_frame_f32(const float* frame, int64_t n, int ncols, int xcol, int p, int ycol, int t, const float* mask,
                              double* M, cudaStream_t s) {
  if (!moments_tcgen05_frame_supported(n, ncols, xcol, p, ycol, t)) return -1;
  if (reinterpret_cast<uintptr_t>(frame) & 15) return -1;
  Geometry g{frame, frame, 0, 0, ncols, xcol, ycol, true};
  return moments_tcgen05_core(g, mask, n, p, t, M, s);
}

}  // namespace pdsb
