// FP64 issue-rate probe for sm_100a: how many m8n8k4 / m16n8k8 DMMAs and DFMAs an SM retires per cycle.
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/_ab/fp64_peak profiles/fp64_peak.cu
// Every warp keeps NACC independent accumulators; ITERS rounds; grid = SMs x ctas, block = warps x 32.
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, {%12, %13, %14, %15}, {%0, %1, %2, %3};"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a, double b) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 12345.678) out[0] = s;
}

template <int NACC>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  if (s == 12345.678f) out[0] = s;
}

template <typename F>
static float time_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double* out; cudaMalloc(&out, 64);
  const int iters = 4096;
  printf("device %s, %d SMs, clock attr %.0f MHz\n", pr.name, sms, khz / 1e3);
  for (int warps : {4, 8, 16, 32}) {
    const int ctas = 1, th = warps * 32;
    constexpr int NA = 8;
    float ms;
    double n_inst = (double)sms * ctas * warps * NA * iters;
    ms = time_ms([&] { k_dmma884<NA><<<sms * ctas, th>>>(out, iters, 1.0000001, 0.9999999); });
    printf("warps/SM %2d  m8n8k4   %8.3f ms  %7.2f TFLOP/s  (%.2f ns per warp-inst per SM)\n", warps, ms, n_inst * 512 / ms / 1e9, ms * 1e6 / (n_inst / sms));
    ms = time_ms([&] { k_dmma1688<NA><<<sms * ctas, th>>>(out, iters, 1.0000001, 0.9999999); });
    printf("warps/SM %2d  m16n8k8  %8.3f ms  %7.2f TFLOP/s  (%.2f ns)\n", warps, ms, n_inst * 2048 / ms / 1e9, ms * 1e6 / (n_inst / sms));
    ms = time_ms([&] { k_dmma16816<NA><<<sms * ctas, th>>>(out, iters, 1.0000001, 0.9999999); });
    printf("warps/SM %2d  m16n8k16 %8.3f ms  %7.2f TFLOP/s  (%.2f ns)\n", warps, ms, n_inst * 4096 / ms / 1e9, ms * 1e6 / (n_inst / sms));
    ms = time_ms([&] { k_dfma<NA><<<sms * ctas, th>>>(out, iters, 1.0000001, 0.9999999); });
    printf("warps/SM %2d  DFMA     %8.3f ms  %7.2f TFLOP/s  (%.2f ns)\n", warps, ms, n_inst * 64 / ms / 1e9, ms * 1e6 / (n_inst / sms));
    ms = time_ms([&] { k_ffma<NA><<<sms * ctas, th>>>((float*)out, iters, 1.0000001f, 0.9999999f); });
    printf("warps/SM %2d  FFMA     %8.3f ms  %7.2f TFLOP/s  (%.2f ns)\n", warps, ms, n_inst * 64 / ms / 1e9, ms * 1e6 / (n_inst / sms));
  }
  return 0;
}
