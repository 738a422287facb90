// K2b — the headline kernel: moments  M = [X|Y|1]^T [X|Y|1]  for f32 frames on the 5th-gen tensor cores.
//
// Replaces faer's matmul in get_xtx_with_lambda / build_xty (/root/reference/src/linear/lr/lr_solvers.rs:183-211,
// 262-278) and the column sums of faer_coordinate_descent (:483-484): one pass over the frame instead of three.
//
// Shape of the problem: Z~ = [Z | 1] has q~ = p + t + 1 <= 64 columns and n ~ 1e8 rows, i.e. a GEMM with M = N = q~
// and K = n.  At q~ = 34 the FP32 SIMT pipes cannot keep up with HBM (595 FMA per row vs 132 bytes per row), so the
// Gram goes to tcgen05.mma kind::tf32 — and because one TF32 product loses 13 mantissa bits it is computed as the
// classic 3-term split  x = hi + lo  (hi = x with the low 13 mantissa bits cleared, lo = x - hi, exact in fp32):
//        G = HH + LH + LH^T (+ LL ~ 2^-22, dropped),     HH = hi^T hi,  LH = lo^T hi.
// Both products come out of ONE instruction stream by stacking [hi ; lo] along the MMA's M = 128 dimension (which is
// free: an M = 64 and an M = 128 instruction cost the same N/2 cycles):
//        A  (TMEM, 128 lanes x 8 cols per MMA) : lanes 0..63 = hi of column m, lanes 64..127 = lo of column m
//        B  (SMEM, N x 8, K-major, 128B swizzle): hi of column n            ->  D[0:64] = HH,  D[64:128] = LH
// Data flow per CTA (persistent, one CTA per SM, contiguous range of 32-row boxes):
//   warp 0      TMA producer  : cp.async.bulk.tensor {32 rows x q cols} -> raw ring (128B-swizzled, K-major)
//   warps 2-5   converters    : raw row m -> registers -> hi / lo -> tcgen05.st into the TMEM A ring; the hi half is
//                               also written to the B ring (so B == hi bit-exactly, independent of how the tensor core
//                               would round a raw fp32 operand); the "1" column (or the row mask) is synthesised here
//   warp 1      MMA issuer    : 4 x tcgen05.mma (K = 8) per box, accumulating in TMEM (fp32)
//   warps 6-9   epilogue      : every FLUSH_BOXES boxes the accumulator is drained with tcgen05.ld and added to f64
//                               registers (double-buffered D), so fp32 accumulation error never grows with n
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
constexpr int MAX_RAW_STAGES = 4;       // TMA landing ring (3 when N = 64: shared-memory budget)
constexpr int AB_STAGES = 3;            // TMEM-A / SMEM-B ring (3 x 128 columns + 2 x 64 accumulator columns = 512)
constexpr int FLUSH_STAGES = 2;         // accumulate 2 stages = 256 rows in fp32 (RZ accumulation) before draining to f64
constexpr int CONV_SETS = 2;            // converter warp sets (4 warps each), alternating stages
constexpr int EPI_SETS = 2;             // epilogue warp sets, each draining half of the accumulator columns
constexpr int NUM_WARPS = 2 + 4 * CONV_SETS + 4 * EPI_SETS;   // TMA, MMA, converters, epilogue
constexpr int PF_DIST = 0;              // L2-prefetch distance of the producer warp (stages); measured: prefetching does not help (1.63 -> 1.80 ms)
constexpr int NUM_THREADS = NUM_WARPS * 32;                   // 576
constexpr int TMEM_COLS = 512;
constexpr int D_COLS = 64;              // columns reserved per accumulator buffer
constexpr int A_COL0 = 2 * D_COLS;      // first column of the A ring
constexpr int A_SLOT_COLS = BPS * BOX_ROWS;

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
  // bounded spin: a protocol bug must surface as a trapped kernel (an error the host reports), never as a hung GPU.
  // The suspend-time hint lets the hardware park the warp until the phase flips instead of re-issuing the poll:
  // ncu counted ~195 barrier polls per 128-row stage without it, all of them wavefronts on the shared-memory pipe.
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
    if (!done && spins > (1u << 24)) __trap();
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
  uint64_t raw_full[MAX_RAW_STAGES], raw_empty[MAX_RAW_STAGES];
  uint64_t ab_full[AB_STAGES], ab_empty[AB_STAGES];
  uint64_t d_full[2], d_empty[2];
  uint32_t tmem_base;
};

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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr) : "memory");
}

// NB = N / 16 (N = MMA N dimension = padded number of Z~ columns)
template <int NB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ mask, int64_t n, int q /* Z cols */,
                    int64_t stages_total, double* __restrict__ partials /* [grid][128][N] */) {
  constexpr int N = NB * 16;
  constexpr int NH = N / EPI_SETS;                         // accumulator columns per epilogue set
  constexpr int RAW_STAGES = (NB == 4) ? 3 : MAX_RAW_STAGES;
  constexpr uint32_t TILE_BYTES = N * 128;                 // one box-tile: N rows x 128 bytes
  extern __shared__ __align__(1024) unsigned char smem[];
  // carve: raw ring | B ring | barriers
  unsigned char* raw = smem;                                              // RAW_STAGES * BPS * TILE_BYTES
  unsigned char* bt = raw + (size_t)RAW_STAGES * BPS * TILE_BYTES;        // AB_STAGES  * BPS * TILE_BYTES
  Barriers* bars = reinterpret_cast<Barriers*>(bt + (size_t)AB_STAGES * BPS * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = q + 1;                                    // Z~ columns (with the ones / mask column)

  // this CTA's contiguous range of stages
  // stages are dealt round-robin: at any moment the CTAs stream ADJACENT rows of every column (DRAM page locality:
  // with one contiguous range per CTA the chip ran 148 x q far-apart 128-byte streams and topped out at 4.3 TB/s
  // even with all arithmetic removed)
  const uint32_t my_stages = stages_total > (int64_t)blockIdx.x
                                 ? (uint32_t)((stages_total - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
#define STAGE_ROW0(it) (((int64_t)(it) * gridDim.x + blockIdx.x) * STAGE_ROWS)

  if (threadIdx.x == 0) {
    for (int i = 0; i < RAW_STAGES; ++i) { mbar_init(&bars->raw_full[i], 1); mbar_init(&bars->raw_empty[i], 4); }
    for (int i = 0; i < AB_STAGES; ++i) { mbar_init(&bars->ab_full[i], 4); mbar_init(&bars->ab_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], 4 * EPI_SETS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // zero the B ring once (rows >= qt stay zero forever; rows < qt are rewritten every stage)
  for (int i = threadIdx.x; i < AB_STAGES * BPS * (int)TILE_BYTES / 16; i += NUM_THREADS)
    reinterpret_cast<uint4*>(bt)[i] = make_uint4(0, 0, 0, 0);
  // constant rows of the raw tiles: row q = 1.0f, rows q+1 .. N-1 = 0 (position-independent under the swizzle)
  for (int i = threadIdx.x; i < RAW_STAGES * BPS * (N - q) * 8; i += NUM_THREADS) {
    const int tile = i / ((N - q) * 8), rem = i % ((N - q) * 8);
    const int r = q + rem / 8, c = rem % 8;
    const uint32_t val = (r == q) ? 0x3F800000u : 0u;
    *reinterpret_cast<uint4*>(raw + (size_t)tile * TILE_BYTES + (size_t)r * 128 + c * 16) = make_uint4(val, val, val, val);
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

  if (warp == 0) {
    // =============================== TMA producer (warp-uniform loop, one elected lane issues) ===============
    uint32_t rs = 0, ph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      mbar_wait(&bars->raw_empty[rs], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->raw_full[rs], (uint32_t)(BPS * q * 128));
        const int64_t row0 = STAGE_ROW0(it);
#pragma unroll
        for (int b = 0; b < BPS; ++b)
          tma_load_2d(raw + ((size_t)rs * BPS + b) * TILE_BYTES, &tmap, &bars->raw_full[rs], (int)(row0 + b * BOX_ROWS), 0);
      }
      __syncwarp();
      if (++rs == RAW_STAGES) { rs = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform loop, one elected lane issues) ===============
    // instruction descriptor: D=f32, A=B=tf32, both K-major, M=128, N
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t bt_addr = smem_u32(bt);
    uint32_t s = 0, ph = 0, fl = 0, buf = 0, dph = 0;       // ring slot / phase, position in flush group, D buffer / phase
    for (uint32_t it = 0; it < my_stages; ++it) {
      if (fl == 0) mbar_wait(&bars->d_empty[buf], dph ^ 1);
      mbar_wait(&bars->ab_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_addr = tmem + buf * D_COLS;
        const uint32_t a_base = tmem + A_COL0 + s * A_SLOT_COLS;
        const uint64_t bd0 = make_b_desc(bt_addr + s * (BPS * TILE_BYTES));
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
#pragma unroll
          for (int k = 0; k < BOX_ROWS / 8; ++k) {
            // descriptor start address advances in 16-byte units: +TILE_BYTES per box, +32 bytes per K=8 step
            const uint64_t bd = bd0 + (uint64_t)((b * TILE_BYTES + k * 32) >> 4);
            tc_mma_tf32_ts(d_addr, a_base + b * BOX_ROWS + k * 8, bd, idesc, (fl == 0 && b == 0 && k == 0) ? 0u : 1u);
          }
        }
        tc_commit(&bars->ab_empty[s]);
        if (fl == FLUSH_STAGES - 1 || it == my_stages - 1) tc_commit(&bars->d_full[buf]);
      }
      __syncwarp();
      if (++s == AB_STAGES) { s = 0; ph ^= 1; }
      if (++fl == FLUSH_STAGES) { fl = 0; if (buf) dph ^= 1; buf ^= 1; }
    }
  } else if (warp < 2 + 4 * CONV_SETS) {
    // =============================== converters: CONV_SETS x 4 warps; set j owns stages it = j (mod CONV_SETS) ===
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    const bool is_lo = quad >= 2;
    const int m = (quad & 1) * 32 + lane;      // Z~ column handled by this thread
    // Rows >= q of every raw tile are never touched by the TMA (its box has q rows): row q is preset to 1.0 (the
    // "ones" column) and rows > q to 0, so in the common case every thread runs the same select-free code.
    const int mrow = m < N ? m : N - 1;
    const bool is_data = m < q, is_ones = (m == q);
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(mrow & 7);
    for (uint32_t it = set; it < my_stages; it += CONV_SETS) {
      const uint32_t rs = it % RAW_STAGES, rph = (it / RAW_STAGES) & 1;
      const uint32_t s = it % AB_STAGES, sph = (it / AB_STAGES) & 1;
      mbar_wait(&bars->raw_full[rs], rph);
      mbar_wait(&bars->ab_empty[s], sph ^ 1);
      tc_fence_after();
      const int64_t row0 = STAGE_ROW0(it);
      const int64_t left64 = n - row0;
      const int left = left64 > STAGE_ROWS ? STAGE_ROWS : (int)left64;     // valid rows in this stage (>= 1)
      const bool fast = (mask == nullptr) && (left == STAGE_ROWS);          // warp-uniform
#pragma unroll
      for (int b = 0; b < BPS; ++b) {
        uint32_t v[32];
        const unsigned char* rowp = raw + ((size_t)rs * BPS + b) * TILE_BYTES + (size_t)mrow * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ sw) << 4));
          v[4 * c + 0] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        if (!fast) {
          const int nvalid = left - b * BOX_ROWS;                           // rows of this box that exist (may be <= 0)
          if (mask != nullptr) {   // coalesced mask load, broadcast to the ones thread by shuffles
            float mk = 0.0f;
            if (lane < nvalid) mk = __ldg(mask + row0 + b * BOX_ROWS + lane);
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const uint32_t o = __float_as_uint(__shfl_sync(0xffffffffu, mk, k));
              v[k] = is_data ? v[k] : (is_ones ? o : 0u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const uint32_t o = (k < nvalid) ? 0x3F800000u : 0u;
              v[k] = is_data ? v[k] : (is_ones ? o : 0u);
            }
          }
        }
        if (!is_lo) {
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] &= 0xFFFFE000u;          // hi: 10-bit mantissa, exactly representable in TF32
          if (m < qt) {
            unsigned char* brow = bt + ((size_t)s * BPS + b) * TILE_BYTES + (size_t)m * 128;
            const uint32_t swb = (uint32_t)(m & 7);
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(brow + ((c ^ swb) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = __uint_as_float(v[k]);
            const float hi = __uint_as_float(v[k] & 0xFFFFE000u);
            v[k] = __float_as_uint(x - hi);                              // lo: exact in fp32
          }
        }
        tmem_st32(tmem + lane_addr + (uint32_t)(A_COL0 + s * A_SLOT_COLS + b * BOX_ROWS), v);
      }
      // all reads of the raw stage are done (values are in registers / already consumed)
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->raw_empty[rs]);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_async_smem();          // B-tile stores (generic proxy) -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->ab_full[s]);
    }
  } else if (warp < 2 + 4 * CONV_SETS + 4 * EPI_SETS) {
    // =============================== epilogue: EPI_SETS x 4 warps, set e drains columns [e*NH, (e+1)*NH) =========
    const int quad = warp & 3;
    const int eset = (warp - (2 + 4 * CONV_SETS)) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    double acc[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) acc[j] = 0.0;
    const uint32_t groups = (my_stages + FLUSH_STAGES - 1) / FLUSH_STAGES;
    uint32_t buf = 0, dph = 0;
    for (uint32_t g = 0; g < groups; ++g) {
      mbar_wait(&bars->d_full[buf], dph);
      tc_fence_after();
      uint32_t v[NH];
#pragma unroll
      for (int c = 0; c < NH / 8; ++c) tmem_ld8(tmem + lane_addr + (uint32_t)(buf * D_COLS + eset * NH + c * 8), v + 8 * c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->d_empty[buf]);
#pragma unroll
      for (int j = 0; j < NH; ++j) acc[j] += (double)__uint_as_float(v[j]);
      if (buf) dph ^= 1;
      buf ^= 1;
    }
    double* out = partials + ((size_t)blockIdx.x * 128 + (size_t)(quad * 32 + lane)) * N + eset * NH;
#pragma unroll
    for (int j = 0; j < NH; ++j) out[j] = acc[j];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------------------
// Variant "raw-hi": the tensor core reads an fp32 operand as TF32 by IGNORING the low 13 mantissa bits (verified on
// B200 by tests/test_gpu_moments.py::test_tcgen05_raw_hi_matches_explicit_hi, which compares this kernel bit-for-bit
// against the explicit-hi kernel above).  Then hi needs no arithmetic at all:
//     B = the raw TMA tile (no B ring, no generic->async proxy fence),
//     A lanes 0..63 = raw rows copied smem -> registers -> TMEM,   A lanes 64..127 = lo = x - (x & 0xFFFFE000).
// One ring of RING stages holds the raw tiles until the MMA that reads them as B has completed.
// LO_MODE is 0 in production; non-zero values are timing-ablation builds (see launch()).  The round-to-nearest
// experiment (lo = x - cvt.rna.tf32(x)) gave 7e-4 relative error on B200: the tensor core truncates.
constexpr int V4_RING = 6;

// L2 prefetch of the 128-byte lines of one stage (q columns x BPS boxes), spread over the 32 lanes of the producer warp
__device__ __forceinline__ void prefetch_stage(const float* __restrict__ zbase, int64_t ld, int q, int64_t row0, int64_t n, int lane) {
  const int lines = q * BPS;
  for (int idx = lane; idx < lines; idx += 32) {
    const int r = idx / BPS, b = idx % BPS;
    const int64_t row = row0 + b * BOX_ROWS;
    if (row < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(zbase + (int64_t)r * ld + row));
  }
}

struct alignas(8) BarriersV4 {
  uint64_t raw_full[V4_RING], raw_empty[V4_RING];
  uint64_t a_full[AB_STAGES], a_empty[AB_STAGES];
  uint64_t d_full[2], d_empty[2];
  uint32_t tmem_base;
};

template <int NB, int LO_MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gram_tcgen05_rawhi_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ mask, int64_t n, int q,
                          int64_t stages_total, double* __restrict__ partials /* [grid][128][N] */,
                          const float* __restrict__ zbase, int64_t ld, int pf_dist, int blocked) {
  constexpr int N = NB * 16;
  constexpr int NH = N / EPI_SETS;
  constexpr int RING = (NB == 4) ? 5 : V4_RING;
  constexpr uint32_t TILE_BYTES = N * 128;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;                                              // RING * BPS * TILE_BYTES
  BarriersV4* bars = reinterpret_cast<BarriersV4*>(raw + (size_t)RING * BPS * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // stages are dealt round-robin: at any moment the CTAs stream ADJACENT rows of every column (DRAM page locality:
  // with one contiguous range per CTA the chip ran 148 x q far-apart 128-byte streams and topped out at 4.3 TB/s
  // even with all arithmetic removed)
  const uint32_t my_stages = stages_total > (int64_t)blockIdx.x
                                 ? (uint32_t)((stages_total - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
#define STAGE_ROW0(it) (((int64_t)(it) * gridDim.x + blockIdx.x) * STAGE_ROWS)

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&bars->raw_full[i], 1); mbar_init(&bars->raw_empty[i], 1); }
    for (int i = 0; i < AB_STAGES; ++i) { mbar_init(&bars->a_full[i], 4); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], 4 * EPI_SETS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // constant rows of every tile: row q = 1.0f ("ones"), rows q+1 .. N-1 = 0; the TMA box only writes rows 0..q-1
  for (int i = threadIdx.x; i < RING * BPS * (N - q) * 8; i += NUM_THREADS) {
    const int tile = i / ((N - q) * 8), rem = i % ((N - q) * 8);
    const int r = q + rem / 8, c = rem % 8;
    const uint32_t val = (r == q) ? 0x3F800000u : 0u;
    *reinterpret_cast<uint4*>(raw + (size_t)tile * TILE_BYTES + (size_t)r * 128 + c * 16) = make_uint4(val, val, val, val);
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

  if (warp == 0) {
    // ---------------- TMA producer (+ L2 prefetch pf_dist stages ahead, all 32 lanes) ----------------
    // Measured: when every TMA request has to go to DRAM the unit sustains only ~29 GB/s per SM (4.3 TB/s chip-wide,
    // with all arithmetic removed); prefetching the lines into L2 ahead of time turns the TMA loads into L2 hits.
    for (int pfi = 0; pfi < pf_dist && (uint32_t)pfi < my_stages; ++pfi) prefetch_stage(zbase, ld, q, STAGE_ROW0(pfi), n, lane);
    uint32_t rs = 0, ph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      mbar_wait(&bars->raw_empty[rs], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->raw_full[rs], (uint32_t)(BPS * q * 128));
        const int64_t row0 = STAGE_ROW0(it);
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
          if (blocked) tma_load_3d(raw + ((size_t)rs * BPS + b) * TILE_BYTES, &tmap, &bars->raw_full[rs], b * BOX_ROWS, 0, (int)(row0 / STAGE_ROWS));
          else tma_load_2d(raw + ((size_t)rs * BPS + b) * TILE_BYTES, &tmap, &bars->raw_full[rs], (int)(row0 + b * BOX_ROWS), 0);
        }
      }
      __syncwarp();
      if (pf_dist > 0 && it + (uint32_t)pf_dist < my_stages) prefetch_stage(zbase, ld, q, STAGE_ROW0(it + pf_dist), n, lane);
      if (++rs == RING) { rs = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t raw_addr = smem_u32(raw);
    uint32_t s = 0, ph = 0, rs = 0, fl = 0, buf = 0, dph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      if (fl == 0) mbar_wait(&bars->d_empty[buf], dph ^ 1);
      mbar_wait(&bars->a_full[s], ph);      // converters only signal after raw_full: B (the raw tile) has landed too
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_addr = tmem + buf * D_COLS;
        const uint32_t a_base = tmem + A_COL0 + s * A_SLOT_COLS;
        const uint64_t bd0 = make_b_desc(raw_addr + rs * (BPS * TILE_BYTES));
        if (!(LO_MODE & 8))
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
#pragma unroll
          for (int k = 0; k < BOX_ROWS / 8; ++k) {
            const uint64_t bd = bd0 + (uint64_t)((b * TILE_BYTES + k * 32) >> 4);
            tc_mma_tf32_ts(d_addr, a_base + b * BOX_ROWS + k * 8, bd, idesc, (fl == 0 && b == 0 && k == 0) ? 0u : 1u);
          }
        }
        tc_commit(&bars->a_empty[s]);       // TMEM A slot reusable
        tc_commit(&bars->raw_empty[rs]);    // raw tile (B operand) reusable
        if (fl == FLUSH_STAGES - 1 || it == my_stages - 1) tc_commit(&bars->d_full[buf]);
      }
      __syncwarp();
      if (++s == AB_STAGES) { s = 0; ph ^= 1; }
      if (++rs == RING) rs = 0;
      if (++fl == FLUSH_STAGES) { fl = 0; if (buf) dph ^= 1; buf ^= 1; }
    }
  } else if (warp < 2 + 4 * CONV_SETS) {
    // ---------------- converters ----------------
    const int quad = warp & 3;
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    const bool is_lo = quad >= 2;
    const int m = (quad & 1) * 32 + lane;
    const int mrow = m < N ? m : N - 1;
    const bool is_data = m < q, is_ones = (m == q);
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(mrow & 7);
    for (uint32_t it = set; it < my_stages; it += CONV_SETS) {
      const uint32_t rs = it % RING, rph = (it / RING) & 1;
      const uint32_t s = it % AB_STAGES, sph = (it / AB_STAGES) & 1;
      mbar_wait(&bars->raw_full[rs], rph);
      mbar_wait(&bars->a_empty[s], sph ^ 1);
      tc_fence_after();
      const int64_t row0 = STAGE_ROW0(it);
      const int64_t left64 = n - row0;
      const int left = left64 > STAGE_ROWS ? STAGE_ROWS : (int)left64;
      const bool fast = (mask == nullptr) && (left == STAGE_ROWS);
#pragma unroll
      for (int b = 0; b < BPS; ++b) {
        uint32_t v[32];
        unsigned char* rowp = raw + ((size_t)rs * BPS + b) * TILE_BYTES + (size_t)mrow * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 x = make_uint4(c, c + 1, c + 2, c + 3);
          if (!(LO_MODE & 4)) x = *reinterpret_cast<const uint4*>(rowp + ((c ^ sw) << 4));
          v[4 * c + 0] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
        }
        if (!fast) {
          // masked / ragged stage: the ones column differs from the preset constant.  The hi "ones" thread writes the
          // actual values to its A lane and to row q of the tile (the B operand); tiles are reused, so a kernel with a
          // mask takes this path for every stage and always rewrites row q.
          const int nvalid = left - b * BOX_ROWS;
          float mk = 0.0f;
          if (lane < nvalid) mk = mask ? __ldg(mask + row0 + b * BOX_ROWS + lane) : 1.0f;
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const uint32_t o = __float_as_uint(__shfl_sync(0xffffffffu, mk, k));
            v[k] = is_data ? v[k] : (is_ones ? o : 0u);
          }
          if (!is_lo && is_ones) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(rowp + ((c ^ sw) << 4)) = make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          }
        }
        if (is_lo && !(LO_MODE & 2)) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x = __uint_as_float(v[k]);
            v[k] = __float_as_uint(x - __uint_as_float(v[k] & 0xFFFFE000u));
          }
        }
        if (!(LO_MODE & 1)) tmem_st32(tmem + lane_addr + (uint32_t)(A_COL0 + s * A_SLOT_COLS + b * BOX_ROWS), v);
        else if (v[0] == 0x7fc12345u && v[31] == 0x12345u) bars->tmem_base = v[5];   // debug build: keep v alive
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      if (!fast) fence_async_smem();        // the rewritten ones row must be visible to the tensor core
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->a_full[s]);
    }
  } else {
    // ---------------- epilogue ----------------
    const int quad = warp & 3;
    const int eset = (warp - (2 + 4 * CONV_SETS)) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    double acc[NH];
#pragma unroll
    for (int j = 0; j < NH; ++j) acc[j] = 0.0;
    const uint32_t groups = (my_stages + FLUSH_STAGES - 1) / FLUSH_STAGES;
    uint32_t buf = 0, dph = 0;
    for (uint32_t g = 0; g < groups; ++g) {
      mbar_wait(&bars->d_full[buf], dph);
      tc_fence_after();
      uint32_t v[NH];
#pragma unroll
      for (int c = 0; c < NH / 8; ++c) tmem_ld8(tmem + lane_addr + (uint32_t)(buf * D_COLS + eset * NH + c * 8), v + 8 * c);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->d_empty[buf]);
#pragma unroll
      for (int j = 0; j < NH; ++j) acc[j] += (double)__uint_as_float(v[j]);
      if (buf) dph ^= 1;
      buf ^= 1;
    }
    double* out = partials + ((size_t)blockIdx.x * 128 + (size_t)(quad * 32 + lane)) * N + eset * NH;
#pragma unroll
    for (int j = 0; j < NH; ++j) out[j] = acc[j];
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------------------
// Variant "x-only A" (default).  The LSU / shared-memory data pipe is what limits the raw-hi kernel (ncu: 61 % busy
// at 4.0 TB/s): every one of the 128 TMEM lanes costs a shared-memory read and a TMEM store per row, and with
// Z~ = [X | y | 1] (34 columns at p = 32) almost half of the lanes are padding.  Here only the FEATURES go through A:
//     A lanes, per group of 32 feature columns g:  quadrant 2g = hi (raw), quadrant 2g+1 = lo     (p <= 32 -> 2 quadrants)
//     B rows : the raw tile [Z columns as they lie in memory] + lo(y_j) rows + the ones/mask row   (N = q + t + 1 -> pad 16)
//     D      : X'X from hi/lo rows vs X columns (3-term split as before); X'y = hiX.hi_y + loX.hi_y + hiX.lo_y (+ loX.lo_y);
//              column sums from the ones row.   sum(y), sum(y^2) and the row count never touch the tensor core: the
//              lane that builds lo(y) accumulates them (fp32 per box, f64 across boxes).
// Same ring / barrier protocol as the raw-hi kernel.
constexpr int YSIDE_STRIDE = 32;   // doubles per (CTA, converter set): [3u+0] sum y_u, [3u+1] sum y_u^2, [2] count, [12 + 4j + k] y_j.y_k

template <int NB, int NCONV>
__global__ void __launch_bounds__((2 + 4 * NCONV + 4 * EPI_SETS) * 32, 1)
gram_tcgen05_xonly_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ mask, int64_t n, int p, int t,
                          int zx, int zy, int64_t stages_total, double* __restrict__ partials /* [grid][128][N] */,
                          double* __restrict__ yside /* [grid][NCONV][4][3] */, const float* __restrict__ zbase, int64_t ld,
                          int pf_dist, int blocked) {
  constexpr int N = NB * 16;
  constexpr int NH = N / EPI_SETS;
  constexpr int RING = (NB == 4) ? 5 : (NB == 5 ? 4 : V4_RING);
  // TMEM budget: two accumulator buffers of XD_COLS columns + XAB slots of 128 A columns (N = 80 leaves room for 2)
  constexpr int XD_COLS = (NB <= 4) ? 64 : 80;
  constexpr int XA_COL0 = 2 * XD_COLS;
  constexpr int XAB = (NB <= 4) ? AB_STAGES : 2;
  constexpr int NTHREADS = (2 + 4 * NCONV + 4 * EPI_SETS) * 32;
  constexpr uint32_t TILE_BYTES = N * 128;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* raw = smem;
  BarriersV4* bars = reinterpret_cast<BarriersV4*>(raw + (size_t)RING * BPS * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = p + t;                         // rows the TMA box writes
  const int row_loy = q, row_ones = q + t;     // B rows built here
  const int nquad = p > 32 ? 4 : 2;            // active TMEM quadrants
  // stages are dealt round-robin: at any moment the CTAs stream ADJACENT rows of every column (DRAM page locality:
  // with one contiguous range per CTA the chip ran 148 x q far-apart 128-byte streams and topped out at 4.3 TB/s
  // even with all arithmetic removed)
  const uint32_t my_stages = stages_total > (int64_t)blockIdx.x
                                 ? (uint32_t)((stages_total - 1 - blockIdx.x) / gridDim.x + 1) : 0u;
#define STAGE_ROW0(it) (((int64_t)(it) * gridDim.x + blockIdx.x) * STAGE_ROWS)

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&bars->raw_full[i], 1); mbar_init(&bars->raw_empty[i], 1); }
    for (int i = 0; i < XAB; ++i) { mbar_init(&bars->a_full[i], nquad + (nquad == 2 ? 1 : 0)); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], nquad * EPI_SETS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows >= q of every tile are never written by the TMA: lo(y) rows start as 0, the ones row as 1.0, the rest 0
  for (int i = threadIdx.x; i < RING * BPS * (N - q) * 8; i += NTHREADS) {
    const int tile = i / ((N - q) * 8), rem = i % ((N - q) * 8);
    const int r = q + rem / 8, c = rem % 8;
    const uint32_t val = (r == row_ones) ? 0x3F800000u : 0u;
    *reinterpret_cast<uint4*>(raw + (size_t)tile * TILE_BYTES + (size_t)r * 128 + c * 16) = make_uint4(val, val, val, val);
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

  if (warp == 0) {
    // ---------------- TMA producer (+ L2 prefetch pf_dist stages ahead, all 32 lanes) ----------------
    // Measured: when every TMA request has to go to DRAM the unit sustains only ~29 GB/s per SM (4.3 TB/s chip-wide,
    // with all arithmetic removed); prefetching the lines into L2 ahead of time turns the TMA loads into L2 hits.
    for (int pfi = 0; pfi < pf_dist && (uint32_t)pfi < my_stages; ++pfi) prefetch_stage(zbase, ld, q, STAGE_ROW0(pfi), n, lane);
    uint32_t rs = 0, ph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      mbar_wait(&bars->raw_empty[rs], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars->raw_full[rs], (uint32_t)(BPS * q * 128));
        const int64_t row0 = STAGE_ROW0(it);
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
          if (blocked) tma_load_3d(raw + ((size_t)rs * BPS + b) * TILE_BYTES, &tmap, &bars->raw_full[rs], b * BOX_ROWS, 0, (int)(row0 / STAGE_ROWS));
          else tma_load_2d(raw + ((size_t)rs * BPS + b) * TILE_BYTES, &tmap, &bars->raw_full[rs], (int)(row0 + b * BOX_ROWS), 0);
        }
      }
      __syncwarp();
      if (pf_dist > 0 && it + (uint32_t)pf_dist < my_stages) prefetch_stage(zbase, ld, q, STAGE_ROW0(it + pf_dist), n, lane);
      if (++rs == RING) { rs = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t raw_addr = smem_u32(raw);
    uint32_t s = 0, ph = 0, rs = 0, fl = 0, buf = 0, dph = 0;
    for (uint32_t it = 0; it < my_stages; ++it) {
      if (fl == 0) mbar_wait(&bars->d_empty[buf], dph ^ 1);
      mbar_wait(&bars->a_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_addr = tmem + buf * XD_COLS;
        const uint32_t a_base = tmem + XA_COL0 + s * A_SLOT_COLS;
        const uint64_t bd0 = make_b_desc(raw_addr + rs * (BPS * TILE_BYTES));
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
#pragma unroll
          for (int k = 0; k < BOX_ROWS / 8; ++k) {
            const uint64_t bd = bd0 + (uint64_t)((b * TILE_BYTES + k * 32) >> 4);
            tc_mma_tf32_ts(d_addr, a_base + b * BOX_ROWS + k * 8, bd, idesc, (fl == 0 && b == 0 && k == 0) ? 0u : 1u);
          }
        }
        tc_commit(&bars->a_empty[s]);
        tc_commit(&bars->raw_empty[rs]);
        if (fl == FLUSH_STAGES - 1 || it == my_stages - 1) tc_commit(&bars->d_full[buf]);
      }
      __syncwarp();
      if (++s == XAB) { s = 0; ph ^= 1; }
      if (++rs == RING) rs = 0;
      if (++fl == FLUSH_STAGES) { fl = 0; if (buf) dph ^= 1; buf ^= 1; }
    }
  } else if (warp < 2 + 4 * NCONV) {
    // ---------------- converters: set j owns stages it = j (mod NCONV) ----------------
    // p <= 32: quadrant 0 = hi(X), quadrant 1 = lo(X), and the otherwise idle quadrant-2 warp of the set does the y / ones
    // side work (lo(y) rows and the ones row of the B tile, sum y, sum y^2, y_i.y_j, row count), so no warp carries both a
    // full LDS + STTM stream and the side work.  p > 32: quadrants 0..3 = hi/lo of two feature groups, quadrant 0 also
    // does the side work.
    const int quad = warp & 3;
    const uint32_t set = (uint32_t)(warp - 2) >> 2;
    const bool do_x = quad < nquad;
    const bool do_y = (nquad == 2) ? (quad == 2) : (quad == 0);
    if (do_x || do_y) {
      const bool is_lo = quad & 1;
      const int m = (quad >> 1) * 32 + lane;             // feature column
      const bool is_data = do_x && m < p;
      const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
      const int xrow = zx + m;
      const uint32_t sw = (uint32_t)(xrow & 7);
      float sy = 0.0f, syy = 0.0f, sxy[3] = {0.0f, 0.0f, 0.0f};   // sxy[d-1]: y_j . y_{j+d} (multi-target cross moments)
      double dsy = 0.0, dsyy = 0.0, dcnt = 0.0, dxy[3] = {0.0, 0.0, 0.0};
      for (uint32_t it = set; it < my_stages; it += NCONV) {
        const uint32_t rs = it % RING, rph = (it / RING) & 1;
        const uint32_t s = it % XAB, sph = (it / XAB) & 1;
        mbar_wait(&bars->raw_full[rs], rph);
        mbar_wait(&bars->a_empty[s], sph ^ 1);
        tc_fence_after();
        const int64_t row0 = STAGE_ROW0(it);
        const int64_t left64 = n - row0;
        const int left = left64 > STAGE_ROWS ? STAGE_ROWS : (int)left64;
        const bool fast = (mask == nullptr) && (left == STAGE_ROWS);
#pragma unroll
        for (int b = 0; b < BPS; ++b) {
          unsigned char* tile = raw + ((size_t)rs * BPS + b) * TILE_BYTES;
          uint32_t v[32];
          if (is_data) {
            const unsigned char* rowp = tile + (size_t)xrow * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint4 x = *reinterpret_cast<const uint4*>(rowp + ((c ^ sw) << 4));
              v[4 * c + 0] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0u;
          }
          if (do_y) {               // warp-uniform
            // 8 lanes per target: lane 8j + c handles the c-th 16-byte chunk of y_j's row (lo(y) row, sum y, sum y^2)
            if (lane < 8 * t) {
              const int j = lane >> 3, c = lane & 7;
              const int yr = zy + j, lr = row_loy + j;
              const uint4 yv = *reinterpret_cast<const uint4*>(tile + (size_t)yr * 128 + ((c ^ (yr & 7)) << 4));
              const float y0 = __uint_as_float(yv.x), y1 = __uint_as_float(yv.y), y2 = __uint_as_float(yv.z), y3 = __uint_as_float(yv.w);
              uint4 lo;
              lo.x = __float_as_uint(y0 - __uint_as_float(yv.x & 0xFFFFE000u));
              lo.y = __float_as_uint(y1 - __uint_as_float(yv.y & 0xFFFFE000u));
              lo.z = __float_as_uint(y2 - __uint_as_float(yv.z & 0xFFFFE000u));
              lo.w = __float_as_uint(y3 - __uint_as_float(yv.w & 0xFFFFE000u));
              *reinterpret_cast<uint4*>(tile + (size_t)lr * 128 + ((c ^ (lr & 7)) << 4)) = lo;
              sy += (y0 + y1) + (y2 + y3);
              syy = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, syy))));
#pragma unroll
              for (int d = 1; d < 4; ++d)
                if (j + d < t) {
                  const int kr = zy + j + d;
                  const uint4 kv = *reinterpret_cast<const uint4*>(tile + (size_t)kr * 128 + ((c ^ (kr & 7)) << 4));
                  sxy[d - 1] = fmaf(y0, __uint_as_float(kv.x), fmaf(y1, __uint_as_float(kv.y),
                               fmaf(y2, __uint_as_float(kv.z), fmaf(y3, __uint_as_float(kv.w), sxy[d - 1]))));
                }
            }
            if (!fast) {
              // masked / ragged stage: lanes 0..7 rewrite the ones row with the 32 mask values of this box
              const int nvalid = left - b * BOX_ROWS;
              if (lane < 8) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int k = lane * 4 + e;
                  o[e] = (k < nvalid) ? (mask ? __ldg(mask + row0 + b * BOX_ROWS + k) : 1.0f) : 0.0f;
                }
                dcnt += (double)((o[0] + o[1]) + (o[2] + o[3]));
                *reinterpret_cast<uint4*>(tile + (size_t)row_ones * 128 + ((lane ^ (row_ones & 7)) << 4)) =
                    make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
              }
            }
          }
          if (is_lo) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const float x = __uint_as_float(v[k]);
              v[k] = __float_as_uint(x - __uint_as_float(v[k] & 0xFFFFE000u));
            }
          }
          if (do_x) tmem_st32(tmem + lane_addr + (uint32_t)(XA_COL0 + s * A_SLOT_COLS + b * BOX_ROWS), v);
        }
        dsy += (double)sy; dsyy += (double)syy; sy = 0.0f; syy = 0.0f;
#pragma unroll
        for (int d = 0; d < 3; ++d) { dxy[d] += (double)sxy[d]; sxy[d] = 0.0f; }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        if (do_y) fence_async_smem();         // lo(y) / ones rows written through the generic proxy
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->a_full[s]);
      }
      if (do_y) {
        // reduce the 8 chunk-lanes of every target (fixed order -> reproducible) and the masked-row count
        double* ys = yside + ((size_t)blockIdx.x * NCONV + set) * YSIDE_STRIDE;
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
    }
  } else {
    // ---------------- epilogue ----------------
    const int quad = warp & 3;
    const int eset = (warp - (2 + 4 * NCONV)) >> 2;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    if (quad < nquad) {
      double acc[NH];
#pragma unroll
      for (int j = 0; j < NH; ++j) acc[j] = 0.0;
      const uint32_t groups = (my_stages + FLUSH_STAGES - 1) / FLUSH_STAGES;
      uint32_t buf = 0, dph = 0;
      for (uint32_t g = 0; g < groups; ++g) {
        mbar_wait(&bars->d_full[buf], dph);
        tc_fence_after();
        // drain in chunks of 8 columns (keeps the register footprint flat for N = 80)
#pragma unroll
        for (int c = 0; c < NH / 8; ++c) {
          uint32_t v[8];
          tmem_ld8(tmem + lane_addr + (uint32_t)(buf * XD_COLS + eset * NH + c * 8), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c * 8 + j] += (double)__uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->d_empty[buf]);
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

// finalize for the x-only variant: moments order [X | Y | 1]
__global__ void gram_finalize_xonly_kernel(const double* __restrict__ partials, const double* __restrict__ yside, int nparts,
                                           int nconv, int N, int p, int t, int zx, int zy, int64_t n, int masked,
                                           double* __restrict__ M) {
  const int q1 = p + t + 1;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= q1 * q1) return;
  int i = idx / q1, j = idx % q1;
  if (i > j) { int x = i; i = j; j = x; }          // evaluate the upper triangle, mirror -> exactly symmetric
  const int q = p + t;
  auto hi_lane = [&](int a) { return (a >> 5) * 64 + (a & 31); };
  auto sum_d = [&](int lanei, int col) { double s = 0.0; for (int k = 0; k < nparts; ++k) s += partials[((size_t)k * 128 + lanei) * N + col]; return s; };
  double r;
  if (j < p) {                                     // X'X
    const int a = i, b = j;
    const double hh = 0.5 * (sum_d(hi_lane(a), zx + b) + sum_d(hi_lane(b), zx + a));
    r = hh + sum_d(hi_lane(a) + 32, zx + b) + sum_d(hi_lane(b) + 32, zx + a);
  } else if (i < p && j < p + t) {                 // X'y
    const int a = i, k = j - p;
    r = sum_d(hi_lane(a), zy + k) + sum_d(hi_lane(a) + 32, zy + k) + sum_d(hi_lane(a), q + k) + sum_d(hi_lane(a) + 32, q + k);
  } else if (i < p) {                              // column sums (ones / mask row)
    r = sum_d(hi_lane(i), q + t) + sum_d(hi_lane(i) + 32, q + t);
  } else {
    // y / ones block from the side accumulators
    double sy[4] = {0, 0, 0, 0}, syy[4] = {0, 0, 0, 0}, cnt = 0.0, cross = 0.0;
    const bool want_cross = (j < p + t) && (i != j);
    for (int k = 0; k < nparts * nconv; ++k) {
      const double* ys = yside + (size_t)k * YSIDE_STRIDE;
      for (int u = 0; u < t; ++u) { sy[u] += ys[u * 3 + 0]; syy[u] += ys[u * 3 + 1]; }
      cnt += ys[2];
      if (want_cross) cross += ys[12 + (i - p) * 4 + (j - p)];
    }
    const double count = masked ? cnt : (double)n;
    if (j == p + t) r = (i == p + t) ? count : sy[i - p];
    else r = (i == j) ? syy[i - p] : cross;         // y_i . y_j from the side lanes (exact products, f64 across stages)
  }
  M[(size_t)i * q1 + j] = r;
  M[(size_t)j * q1 + i] = r;
}

// Sum the per-CTA partials in a fixed order, then  G~[a][b] = HH[a][b] + LH[a][b] + LH[b][a]  and permute the Z~
// columns (targets may precede the features in memory) into the moments order [X | Y | 1].
__global__ void gram_finalize_kernel(const double* __restrict__ partials, int nparts, int N, int p, int t, int zx, int zy,
                                     double* __restrict__ M) {
  // one warp per output element of the upper triangle: lane l sums parts l, l+32, ... then a fixed-order xor tree
  // (bit-reproducible); the mirrored element gets the same value -> exactly symmetric
  const int q1 = p + t + 1;
  const int lane = threadIdx.x & 31;
  const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (idx >= q1 * q1) return;
  const int i = idx / q1, j = idx % q1;
  if (i > j) return;
  auto zcol = [&](int c) { return c < p ? zx + c : (c < p + t ? zy + (c - p) : p + t); };
  const int a = zcol(i), b = zcol(j);
  double hh = 0.0, hh_t = 0.0, lh_ab = 0.0, lh_ba = 0.0;
  for (int k = lane; k < nparts; k += 32) {
    const double* P = partials + (size_t)k * 128 * N;
    hh += P[(size_t)a * N + b];
    hh_t += P[(size_t)b * N + a];
    lh_ab += P[(size_t)(64 + a) * N + b];
    lh_ba += P[(size_t)(64 + b) * N + a];
  }
  for (int off = 16; off; off >>= 1) {
    hh += __shfl_xor_sync(0xffffffffu, hh, off);
    hh_t += __shfl_xor_sync(0xffffffffu, hh_t, off);
    lh_ab += __shfl_xor_sync(0xffffffffu, lh_ab, off);
    lh_ba += __shfl_xor_sync(0xffffffffu, lh_ba, off);
  }
  if (lane == 0) {
    const double r = 0.5 * (hh + hh_t) + (lh_ab + lh_ba);
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

// geometry shared by the support check and the launcher
struct Geometry { const float* base; int q; int zx, zy; bool ok; bool blocked; };

Geometry analyse(const float* X, int64_t ldx, const float* Y, int64_t ldy, int p, int t) {
  Geometry g{nullptr, p + t, 0, 0, false, false};
  if (p < 1 || t < 1 || p > 64 || (p + t + 1 > 64 && p + 2 * t + 1 > 80) || t > 8) return g;
  if (ldx != ldy || (ldx % 4) != 0) return g;
  if (Y == X + (size_t)p * ldx) { g.base = X; g.zx = 0; g.zy = p; g.ok = true; }          // [X | Y]
  else if (X == Y + (size_t)t * ldy) { g.base = Y; g.zx = t; g.zy = 0; g.ok = true; }     // [Y | X]
  if (g.ok && (reinterpret_cast<uintptr_t>(g.base) & 15)) g.ok = false;
  return g;
}

// 1 (default): raw-hi, hardware truncation (measured on B200: bit-identical to the explicit-hi kernel);
// 3: x-only A operand (half the LSU traffic, but measured slower: 1.83 ms vs 1.63 ms per 5e7 x 33 rows);
// 0: explicit hi (B ring) — kept as the cross-check; 2: raw-hi assuming round-to-nearest (WRONG on B200: 7e-4
// relative error, kept only to document the experiment in profiles/tc_modes.py).
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

template <int NB, int LO_MODE>
int launch_rawhi(const CUtensorMap& tmap, const float* mask, int64_t n, int q, int64_t stages_total, int grid,
                 double* partials, const float* zbase, int64_t ld, int blocked, cudaStream_t s) {
  constexpr int N = NB * 16;
  constexpr int RING = (NB == 4) ? 5 : V4_RING;
  const size_t smem = (size_t)RING * BPS * N * 128 + sizeof(BarriersV4) + 256;
  auto k = gram_tcgen05_rawhi_kernel<NB, LO_MODE>;
  PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static int pf = [] { const char* e = getenv("PDSB_TC_PF"); return e ? atoi(e) : PF_DIST; }();
  k<<<grid, NUM_THREADS, smem, s>>>(tmap, mask, n, q, stages_total, partials, zbase, ld, pf, blocked);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

template <int NB>
int launch(const CUtensorMap& tmap, const float* mask, int64_t n, int q, int64_t stages_total, int grid, double* partials,
           const float* zbase, int64_t ld, int blocked, cudaStream_t s) {
  if (tc_mode() == 1 || blocked) {
    // PDSB_TC_DBG (timing ablations only, results are garbage): 1 no TMEM store, 2 no lo arithmetic, 4 no smem loads,
    // 8 no MMA, 15 all of them
    static int dbg = [] { const char* e = getenv("PDSB_TC_DBG"); return e ? atoi(e) : 0; }();
    if (NB == 3) {
      switch (dbg) {
        case 1: return launch_rawhi<NB, 1>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        case 2: return launch_rawhi<NB, 2>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        case 4: return launch_rawhi<NB, 4>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        case 8: return launch_rawhi<NB, 8>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        case 7: return launch_rawhi<NB, 7>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        case 15: return launch_rawhi<NB, 15>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
        default: break;
      }
    }
    return launch_rawhi<NB, 0>(tmap, mask, n, q, stages_total, grid, partials, zbase, ld, blocked, s);
  }
  constexpr int N = NB * 16;
  constexpr int RAW_STAGES = (NB == 4) ? 3 : MAX_RAW_STAGES;
  const size_t smem = (size_t)(RAW_STAGES + AB_STAGES) * BPS * N * 128 + sizeof(Barriers) + 256;
  auto k = gram_tcgen05_kernel<NB>;
  PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k<<<grid, NUM_THREADS, smem, s>>>(tmap, mask, n, q, stages_total, partials);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

}  // namespace

void set_tc_mode(int m) { g_tc_mode.store(m); }

template <int NB, int NCONV>
int launch_xonly(const CUtensorMap& tmap, const float* mask, int64_t n, int p, int t, int zx, int zy, int64_t stages_total,
                 int grid, double* partials, double* yside, const float* zbase, int64_t ld, int blocked, cudaStream_t s) {
  constexpr int N = NB * 16;
  constexpr int RING = (NB == 4) ? 5 : (NB == 5 ? 4 : V4_RING);
  const size_t smem = (size_t)RING * BPS * N * 128 + sizeof(BarriersV4) + 256;
  auto k = gram_tcgen05_xonly_kernel<NB, NCONV>;
  PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static int pf = [] { const char* e = getenv("PDSB_TC_PF"); return e ? atoi(e) : PF_DIST; }();
  k<<<grid, (2 + 4 * NCONV + 4 * EPI_SETS) * 32, smem, s>>>(tmap, mask, n, p, t, zx, zy, stages_total, partials, yside, zbase, ld, pf, blocked);
  PDSB_LAUNCH_OK();
  count_launch();
  return 0;
}

int xonly_nconv() {
  static int v = [] { const char* e = getenv("PDSB_TC_NCONV"); int x = e ? atoi(e) : 2; return (x == 3) ? 3 : 2; }();
  return v;
}

bool moments_tcgen05_supported(const float* X, int64_t ldx, const float* Y, int64_t ldy, int64_t n, int p, int t) {
  if (getenv("PDSB_DISABLE_TCGEN05")) return false;
  if (n < 4096) return false;                 // latency-bound sizes stay on the SIMT kernel
  if (n >= (int64_t(1) << 31) - STAGE_ROWS) return false;   // the 2-D tensor map is addressed with int32 row coordinates
  if (!get_encode_fn()) return false;
  return analyse(X, ldx, Y, ldy, p, t).ok;
}

static int moments_tcgen05_core(const Geometry& g, int64_t ldx, const float* mask, int64_t n, int p, int t, double* M,
                                cudaStream_t s);

int moments_tcgen05_f32(const float* X, int64_t ldx, const float* Y, int64_t ldy, const float* mask, int64_t n, int p,
                        int t, double* M, cudaStream_t s) {
  const Geometry g = analyse(X, ldx, Y, ldy, p, t);
  if (!g.ok) return -1;
  return moments_tcgen05_core(g, ldx, mask, n, p, t, M, s);
}

// row-blocked frame: [block][column][FRAME_ROWS]; the frame holds exactly the p + t columns, X at xcol, Y at ycol
bool moments_tcgen05_frame_supported(int64_t n, int ncols, int xcol, int p, int ycol, int t) {
  if (getenv("PDSB_DISABLE_TCGEN05") || !get_encode_fn()) return false;
  if (n < 4096 || p < 1 || t < 1 || p > 64 || (p + t + 1 > 64 && (p + 2 * t + 1 > 80 || t > 4)) || ncols != p + t) return false;
  return (xcol == 0 && ycol == p) || (ycol == 0 && xcol == t);
}

int moments_tcgen05_frame_f32(const float* frame, int64_t n, int ncols, int xcol, int p, int ycol, int t, const float* mask,
                              double* M, cudaStream_t s) {
  if (!moments_tcgen05_frame_supported(n, ncols, xcol, p, ycol, t)) return -1;
  if (reinterpret_cast<uintptr_t>(frame) & 15) return -1;
  Geometry g{frame, p + t, xcol, ycol, true, true};
  return moments_tcgen05_core(g, 0, mask, n, p, t, M, s);
}

static int moments_tcgen05_core(const Geometry& g, int64_t ldx, const float* mask, int64_t n, int p, int t, double* M,
                                cudaStream_t s) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -1;
  const int q = g.q, qt = q + 1;
  // features-only A operand: chosen explicitly (variant 3) or whenever Z~ has more than 64 columns (p up to 64)
  const bool xonly = (tc_mode() == 3 || p + t + 1 > 64) && (p + 2 * t + 1 <= 80) && p <= 64 && t <= 4;
  const int N = xonly ? ((p + 2 * t + 1 + 15) / 16) * 16 : ((qt + 15) / 16) * 16;
  CUtensorMap tmap;
  CUresult cr;
  if (g.blocked) {
    // row-blocked frame: [block][column][128 rows] -> every 128-row x q stage is ONE contiguous 512*q-byte run in HBM.
    // (column-major frames cap this kernel at 4.3 TB/s even with all arithmetic removed; blocked: 6.5 TB/s)
    cuuint64_t dims3[3] = {(cuuint64_t)STAGE_ROWS, (cuuint64_t)q, (cuuint64_t)ceil_div(n, STAGE_ROWS)};
    cuuint64_t strides3[2] = {(cuuint64_t)STAGE_ROWS * sizeof(float), (cuuint64_t)STAGE_ROWS * q * sizeof(float)};
    cuuint32_t box3[3] = {(cuuint32_t)BOX_ROWS, (cuuint32_t)q, 1};
    cuuint32_t estr3[3] = {1, 1, 1};
    cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(g.base), dims3, strides3, box3, estr3,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)q};
    cuuint64_t strides[1] = {(cuuint64_t)ldx * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BOX_ROWS, (cuuint32_t)q};
    cuuint32_t estr[2] = {1, 1};
    cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(g.base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (cr != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)cr); return 1; }
  const int64_t stages_total = ceil_div(n, STAGE_ROWS);
  int grid = sm_count();
  if (stages_total < grid) grid = (int)stages_total;
  double* partials = nullptr;
  if (dev_alloc((void**)&partials, ((size_t)grid * 128 * N + (size_t)grid * 3 * YSIDE_STRIDE) * sizeof(double), s)) return 1;
  double* yside = partials + (size_t)grid * 128 * N;
  int rc;
  const int q1 = p + t + 1;
  if (xonly) {
    const int nconv = xonly_nconv();
#define PDSB_XO(NBV) (nconv == 3 ? launch_xonly<NBV, 3>(tmap, mask, n, p, t, g.zx, g.zy, stages_total, grid, partials, yside, g.base, ldx, g.blocked ? 1 : 0, s) \
                                 : launch_xonly<NBV, 2>(tmap, mask, n, p, t, g.zx, g.zy, stages_total, grid, partials, yside, g.base, ldx, g.blocked ? 1 : 0, s))
    switch (N / 16) {
      case 1: rc = PDSB_XO(1); break;
      case 2: rc = PDSB_XO(2); break;
      case 3: rc = PDSB_XO(3); break;
      case 4: rc = PDSB_XO(4); break;
      default: rc = PDSB_XO(5); break;
    }
#undef PDSB_XO
    if (!rc) {
      gram_finalize_xonly_kernel<<<(q1 * q1 + 127) / 128, 128, 0, s>>>(partials, yside, grid, nconv, N, p, t, g.zx, g.zy, n,
                                                                    mask ? 1 : 0, M);
      cudaError_t e = cudaGetLastError();
      count_launch();
      if (e != cudaSuccess) { set_error("gram finalize launch failed: %s", cudaGetErrorString(e)); rc = 1; }
    }
    dev_free(partials, s);
    return rc;
  }
  switch (N / 16) {
    case 1: rc = launch<1>(tmap, mask, n, q, stages_total, grid, partials, g.base, ldx, g.blocked ? 1 : 0, s); break;
    case 2: rc = launch<2>(tmap, mask, n, q, stages_total, grid, partials, g.base, ldx, g.blocked ? 1 : 0, s); break;
    case 3: rc = launch<3>(tmap, mask, n, q, stages_total, grid, partials, g.base, ldx, g.blocked ? 1 : 0, s); break;
    default: rc = launch<4>(tmap, mask, n, q, stages_total, grid, partials, g.base, ldx, g.blocked ? 1 : 0, s); break;
  }
  if (!rc) {
    gram_finalize_kernel<<<(q1 * q1 + 7) / 8, 256, 0, s>>>(partials, grid, N, p, t, g.zx, g.zy, M);   // 8 warps = 8 elements per block
    cudaError_t e = cudaGetLastError();
    count_launch();
    if (e != cudaSuccess) { set_error("gram finalize launch failed: %s", cudaGetErrorString(e)); rc = 1; }
  }
  dev_free(partials, s);
  return rc;
}

}  // namespace pdsb
