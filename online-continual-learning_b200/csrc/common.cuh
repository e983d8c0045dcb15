// common.cuh -- shared helpers for libb200ocl (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/b200ocl.h"

namespace b200ocl {

void set_error(const char* fmt, ...);
// Optional per-launch timing (bench.py roofline): when enabled, every launch is bracketed by CUDA
// events on its own stream and accumulated per kernel class.  Off by default (no overhead).
extern bool g_prof_on;
void prof_begin(const char* kernel_class, double work, cudaStream_t stream);
void prof_end();
extern std::atomic<uint64_t> g_launches;
int sm_count();
// Function attributes (cudaFuncSetAttribute) are per device: every launcher keeps one flag per device.
#define B200OCL_MAX_DEVICES 64
inline int device_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= B200OCL_MAX_DEVICES) dev = 0;
  return dev;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

#define B200OCL_CHECK_ARG(cond, msg)                      \
  do {                                                    \
    if (!(cond)) {                                        \
      b200ocl::set_error("%s: %s", __func__, msg);        \
      return B200OCL_EINVAL;                              \
    }                                                     \
  } while (0)

#define B200OCL_CUDA(call)                                                                    \
  do {                                                                                        \
    cudaError_t err__ = (call);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      (void)cudaGetLastError(); /* clear the non-sticky error state */                        \
      b200ocl::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(err__));    \
      return B200OCL_ECUDA;                                                                   \
    }                                                                                         \
  } while (0)

// Count the launch and surface launch-configuration errors immediately.
// Declare the kernel class and its algorithmic work (FLOPs or bytes) right before a launch.
#define B200OCL_PROF(kernel_class, work, stream)                                              \
  do {                                                                                        \
    if (b200ocl::g_prof_on) b200ocl::prof_begin(kernel_class, (double)(work), stream);        \
  } while (0)

#define B200OCL_LAUNCHED()                                                                    \
  do {                                                                                        \
    if (b200ocl::g_prof_on) b200ocl::prof_end();                                              \
    b200ocl::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
    cudaError_t err__ = cudaGetLastError();                                                   \
    if (err__ != cudaSuccess) {                                                               \
      b200ocl::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(err__)); \
      return B200OCL_ECUDA;                                                                   \
    }                                                                                         \
  } while (0)

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
  return v;
}

}  // namespace b200ocl
