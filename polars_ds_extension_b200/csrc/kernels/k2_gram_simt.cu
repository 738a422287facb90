// K2a — generic SIMT moments kernel:  M = [X|Y|1]^T diag(w) [X|Y|1]   (q1 x q1, f64, row-major)
//
// Replaces the reference's get_xtx_with_lambda + build_xty + column sums
// (/root/reference/src/linear/lr/lr_solvers.rs:183-211, 262-278, 483-484) and x^T w x
// (src/num_ext/linear_regression.rs:1026-1027) with ONE pass over the data.
//
// This is the path for f64 data, weighted fits, tiny inputs and shapes the tcgen05 kernel does not take
// (k2_gram_tcgen05.cu is the f32 headline path).  Layout: X col-major [n x p] (ldx), Y col-major [n x t] (ldy).
// Each CTA walks row tiles of TILE_R rows: tile -> shared memory (row-major, row stride S), every thread owns
// up to MAXT 4x4 blocks of the upper triangle; per tile the block is accumulated in T (a TILE_R-long FMA chain)
// and then added to f64 accumulators, so f32 rounding never grows with n.  Per-CTA partials are reduced in a
// fixed order by a second kernel -> bit-reproducible results.
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

  // tile-pair decode for this thread
  int ti[MAXT], tj[MAXT];
  bool act[MAXT];
#pragma unroll
  for (int m = 0; m < MAXT; ++m) {
    int idx = tid + m * 256;
    act[m] = idx < ntp;
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
    // ---- load tile (coalesced along rows) ----
    for (int idx = tid; idx < tile_r * q1; idx += 256) {
      int c = idx / tile_r, r = idx - c * tile_r;
      int64_t row = row0 + r;
      T v = T(0);
      if (row < n) {
        if (bstride) {   // row-blocked frame: [block][column][FRAME_ROWS]
          const int64_t o = (row >> 7) * bstride + (row & 127);
          if (c < p) v = X[o + ((int64_t)c << 7)];
          else if (c < p + t) v = Y[o + ((int64_t)(c - p) << 7)];
          else v = mask ? mask[row] : T(1);
        } else
        if (c < p) v = X[(int64_t)c * ldx + row];
        else if (c < p + t) v = Y[(int64_t)(c - p) * ldy + row];
        else v = mask ? mask[row] : T(1);
      }
      Zs[r * S + c] = v;
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
      for (int r = 0; r < tile_r; ++r) {
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
  double* out = partials + (size_t)blockIdx.x * q1 * q1;
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

__global__ void reduce_partials_kernel(const double* __restrict__ partials, int nparts, int len,
                                       double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int k = 0; k < nparts; ++k) s += partials[(size_t)k * len + i];
  out[i] = s;
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
  const int nt = (q1 + 3) / 4, ntp = nt * (nt + 1) / 2;
  const int maxt = (ntp + 255) / 256;
  const int S = nt * 4;  // row stride, multiple of 4 elements
  // tile rows: keep the tile under ~64 KB
  int tile_r = 128;
  while (tile_r > 16 && (size_t)tile_r * (S + 1) * sizeof(T) > 64 * 1024) tile_r >>= 1;
  size_t smem = (size_t)tile_r * (S + 1) * sizeof(T);
  int64_t ntiles = ceil_div(n > 0 ? n : 1, tile_r);
  int grid = (int)std::min<int64_t>(ntiles, (int64_t)sm_count() * 2);
  if (grid < 1) grid = 1;
  double* partials = nullptr;
  if (dev_alloc((void**)&partials, (size_t)grid * q1 * q1 * sizeof(double), s)) return 1;
  int rc;
  switch (maxt) {
    case 1: rc = launch_gram<T, 1>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    case 2: rc = launch_gram<T, 2>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    case 3: case 4: rc = launch_gram<T, 4>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
    default: rc = launch_gram<T, 9>(X, ldx, Y, ldy, w, mask, n, p, t, tile_r, S, grid, smem, partials, bstride, s); break;
  }
  if (rc) { dev_free(partials, s); return rc; }
  int len = q1 * q1;
  reduce_partials_kernel<<<(len + 255) / 256, 256, 0, s>>>(partials, grid, len, M);
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
