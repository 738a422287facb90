// K4 — streaming predict / residual / SSR:  pred = [X|1] beta, resid = Y - pred.
//
// Reference: `let pred = x * &coeffs; let resid = y - &pred;` (/root/reference/src/num_ext/linear_regression.rs:
// 782-785, 635-636) plus the mask re-insertion loop (:790-812) and the residual sums the report needs (:863-875,
// 1036-1038).  One coalesced, vectorised pass: each thread owns VEC consecutive rows (16-byte loads per column
// because X is column-major), beta lives in shared memory, pred and resid are written once; the optional SSR is
// reduced in f64 per block and summed in a fixed order by a second tiny kernel (bit-reproducible).
// HBM-bound: algorithmic bytes per row = (p + t) * s read + 2 t * s written.
#include "../common.h"
#include "kernels.h"

namespace pdsb {

namespace {

template <typename T, int VEC> struct VecT;
template <> struct VecT<float, 4> { using type = float4; };
template <> struct VecT<double, 2> { using type = double2; };
template <> struct VecT<float, 1> { using type = float; };
template <> struct VecT<double, 1> { using type = double; };

template <typename T, int VEC>
__global__ void __launch_bounds__(256)
predict_kernel(const T* __restrict__ X, int64_t ldx, const T* __restrict__ Y, int64_t ldy,
               const T* __restrict__ w, const T* __restrict__ mask, int64_t n, int p, int t, int add_bias,
               const double* __restrict__ beta, const int* __restrict__ status, T* __restrict__ pred,
               T* __restrict__ resid, int64_t ldo, uint8_t* __restrict__ valid, double* __restrict__ ssr_part,
               int64_t bstride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sb = reinterpret_cast<T*>(smem_raw);  // [(p+1) x t]
  const int q = p + (add_bias ? 1 : 0);
  const bool gated = status && (*status != 0);
  for (int i = threadIdx.x; i < (p + 1) * t; i += blockDim.x) {
    int r = i % (p + 1), k = i / (p + 1);
    double v = 0.0;
    if (!gated) {
      if (r < p) v = beta[(size_t)k * q + r];
      else if (add_bias) v = beta[(size_t)k * q + p];
    }
    sb[i] = (T)v;
  }
  __syncthreads();
  using V = typename VecT<T, VEC>::type;
  double ssr_loc[4] = {0.0, 0.0, 0.0, 0.0};   // sum e^2   (up to 4 targets)
  double wssr_loc[4] = {0.0, 0.0, 0.0, 0.0};  // sum w e^2
  const int64_t nvec = (n + VEC - 1) / VEC;
  for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec; iv += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = iv * VEC;
    const bool full = (row + VEC <= n);
    for (int k0 = 0; k0 < t; ++k0) {
      T acc[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[v] = sb[k0 * (p + 1) + p];
      for (int c = 0; c < p; ++c) {
        const T b = sb[k0 * (p + 1) + c];
        const T* src = bstride ? X + (row >> 7) * bstride + ((int64_t)c << 7) + (row & 127) : X + (int64_t)c * ldx + row;
        T xv[VEC];
        if (full) { V tmp = *reinterpret_cast<const V*>(src); memcpy(xv, &tmp, sizeof(V)); }
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) xv[v] = (row + v < n) ? src[v] : T(0);
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = fma(xv[v], b, acc[v]);
      }
      T yv[VEC], rv[VEC], mv[VEC], wv[VEC];
      const T* ys = bstride ? Y + (row >> 7) * bstride + ((int64_t)k0 << 7) + (row & 127) : Y + (int64_t)k0 * ldy + row;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        bool in = row + v < n;
        yv[v] = in ? ys[v] : T(0);
        mv[v] = (in && mask) ? mask[row + v] : T(1);
        wv[v] = (in && w) ? w[row + v] : T(1);
        rv[v] = yv[v] - acc[v];
      }
      double sl = 0.0, swl = 0.0;
#pragma unroll
      for (int v = 0; v < VEC; ++v)
        if (row + v < n && mv[v] != T(0)) {
          double e2 = (double)rv[v] * (double)rv[v];
          sl += e2; swl += (double)wv[v] * e2;
        }
      if (k0 < 4) { ssr_loc[k0] += sl; wssr_loc[k0] += swl; }
      T* pd = pred + (int64_t)k0 * ldo + row;
      T* rd = resid + (int64_t)k0 * ldo + row;
      if (full) {
        V tp, tr; memcpy(&tp, acc, sizeof(V)); memcpy(&tr, rv, sizeof(V));
        *reinterpret_cast<V*>(pd) = tp; *reinterpret_cast<V*>(rd) = tr;
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) if (row + v < n) { pd[v] = acc[v]; rd[v] = rv[v]; }
      }
      if (valid && k0 == 0) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (row + v < n) valid[row + v] = (!gated && mv[v] != T(0)) ? 1 : 0;
      }
    }
  }
  if (ssr_part) {
    __shared__ double red[8][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int k = 0; k < 8; ++k) {
      double v = (k < 4) ? ssr_loc[k] : wssr_loc[k - 4];
      for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      double v = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[threadIdx.x][i];
      ssr_part[(size_t)blockIdx.x * 8 + threadIdx.x] = v;
    }
  }
}

__global__ void ssr_reduce_kernel(const double* __restrict__ part, int nblocks, double* __restrict__ ssr) {
  int k = threadIdx.x;
  if (k >= 8) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += part[(size_t)b * 8 + k];
  ssr[k] = s;
}

}  // namespace

template <typename T>
int predict_resid(const T* X, int64_t ldx, const T* Y, int64_t ldy, const T* w, const T* mask, int64_t n,
                  int p, int t, int add_bias, const double* beta, const int* status, T* pred, T* resid,
                  int64_t ldo, uint8_t* valid, double* ssr, cudaStream_t s, int64_t bstride) {
  if (n <= 0) return 0;
  if (ssr && t > 4) { set_error("predict: ssr supports at most 4 targets"); return 1; }
  constexpr int VEC = sizeof(T) == 4 ? 4 : 2;
  auto aligned = [&](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
  bool vec_ok = aligned(X) && aligned(pred) && aligned(resid) && (bstride || ldx % VEC == 0) && (ldo % VEC == 0);
  const int64_t work = vec_ok ? ceil_div(n, VEC) : n;
  int grid = (int)std::min<int64_t>(ceil_div(work, 256), (int64_t)sm_count() * 8);
  if (grid < 1) grid = 1;
  size_t smem = (size_t)(p + 1) * t * sizeof(T);
  double* part = nullptr;
  if (ssr) { if (dev_alloc((void**)&part, (size_t)grid * 8 * sizeof(double), s)) return 1; }
  if (vec_ok) {
    auto k = predict_kernel<T, VEC>;
    if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, add_bias, beta, status, pred, resid, ldo, valid, part, bstride);
  } else {
    auto k = predict_kernel<T, 1>;
    if (smem > 48 * 1024) PDSB_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(X, ldx, Y, ldy, w, mask, n, p, t, add_bias, beta, status, pred, resid, ldo, valid, part, bstride);
  }
  cudaError_t e = cudaGetLastError();
  count_launch();
  if (e == cudaSuccess && ssr) {
    ssr_reduce_kernel<<<1, 32, 0, s>>>(part, grid, ssr);
    e = cudaGetLastError();
    count_launch();
  }
  if (part) dev_free(part, s);
  if (e != cudaSuccess) { set_error("predict launch failed: %s", cudaGetErrorString(e)); return 1; }
  return 0;
}

template int predict_resid<float>(const float*, int64_t, const float*, int64_t, const float*, const float*,
                                  int64_t, int, int, int, const double*, const int*, float*, float*, int64_t,
                                  uint8_t*, double*, cudaStream_t, int64_t);
template int predict_resid<double>(const double*, int64_t, const double*, int64_t, const double*, const double*,
                                   int64_t, int, int, int, const double*, const int*, double*, double*, int64_t,
                                   uint8_t*, double*, cudaStream_t, int64_t);

}  // namespace pdsb
