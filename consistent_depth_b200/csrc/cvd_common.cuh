// Shared helpers for libcvd_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/cvd.h"

extern thread_local char g_cvd_err[512];
extern long long g_cvd_launches;

static inline int cvd_fail(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt);
  vsnprintf(g_cvd_err, sizeof(g_cvd_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CVD_CHECK_ARG(cond, ...) do { if (!(cond)) return cvd_fail(__VA_ARGS__); } while (0)

// after a launch: count it and surface launch-configuration errors
#define CVD_LAUNCH_OK(name) do {                                              \
    ++g_cvd_launches;                                                         \
    cudaError_t e__ = cudaGetLastError();                                     \
    if (e__ != cudaSuccess) return cvd_fail("%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit load that does not pollute L1 (read-once inputs)
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

static inline int cvd_num_sms() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
  return n;
}
