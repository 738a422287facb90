// Shared host-side helpers of libpds_b200: error channel, launch accounting, CUDA checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <atomic>

namespace pdsb {

// thread-local last error (pdsb_last_error / _polars_plugin_get_last_error_message)
void set_error(const char* fmt, ...);
const char* get_error();

extern std::atomic<int64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

#define PDSB_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ::pdsb::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                        cudaGetErrorString(_e));                                             \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

#define PDSB_LAUNCH_OK()                                                                     \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      ::pdsb::set_error("CUDA launch error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                        __LINE__, cudaGetErrorString(_e));                                   \
      return 1;                                                                              \
    }                                                                                        \
  } while (0)

// number of SMs on the current device (cached per device)
int sm_count();
// make sure a usable CUDA device exists; loud failure otherwise (no CPU fallback anywhere in this library)
int require_device();

// stream-ordered scratch from the driver's default memory pool (release threshold raised once)
int dev_alloc(void** p, size_t bytes, cudaStream_t s);
void dev_free(void* p, cudaStream_t s);

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pdsb
