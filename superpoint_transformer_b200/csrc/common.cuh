// common.cuh — shared helpers for libspt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <float.h>
#include "../../include/spt_b200.h"

#ifndef __CUDA_ARCH__
// host pass
#else
#if __CUDA_ARCH__ < 1000
#error "libspt_b200 is written for sm_100a (B200) only"
#endif
#endif

namespace spt {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return SPT_OK;
}

#define SPT_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      ::spt::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Per-device one-time setup.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the
// CURRENT device only, and the SM count differs between parts: both are keyed by device id
// (a process may drive several GPUs).
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
inline int device_sm_count() {
  static int cache[64] = {0};
  const int dev = current_device() & 63;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}
template <typename Kernel>
inline void ensure_dynamic_smem(Kernel kernel, int bytes, unsigned long long* done_mask) {
  const unsigned long long bit = 1ull << (current_device() & 63);
  if (!(*done_mask & bit)) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    *done_mask |= bit;
  }
}
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// streaming (read-once) loads: keep L1/L2 for the gathered node rows
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}

}  // namespace spt
