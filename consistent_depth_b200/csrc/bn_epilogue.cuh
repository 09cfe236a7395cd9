// BatchNorm2d(train) batch statistics fused into a conv epilogue (shared by conv_tc.cu and conv2.cu).
//
// Each of the 4 epilogue warps holds 32 accumulator rows (pixels) x 16 columns (channels) in registers per chunk.
// accumulate16: per-column sum and sum of squares over the warp's 32 rows by a transposing shuffle butterfly
//   (8+4+2+1+1 shuffles per quantity), added into the warp's private shared-memory row  ws[0..cout) | ws[cout..2cout).
// finalize: CTA partials -> fp64 atomics in global scratch; the LAST CTA (ticket) turns them into
//   a = gamma*rstd, b = beta - mean*a, rstd, mean and updates the running statistics exactly as nn.BatchNorm2d
//   does in train mode (hourglass.py:28,40,43,165: eps, momentum, unbiased running variance), then re-zeroes the scratch.
#pragma once
#include <cuda_runtime.h>

namespace bnepi {

struct Stats {
  double* scratch; const float* gamma; const float* beta; float* rm; float* rv;
  float* a; float* b; float* rstd; float* mean; float eps, mom; long long count;
};

__device__ __forceinline__ void accumulate16(const float* v, bool inside, int lane, float* ws, int cout, int c16)
{
  float sv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) sv[i] = inside ? v[i] : 0.f;
  float tot[2];
#pragma unroll
  for (int qq = 0; qq < 2; ++qq) {
    float t8[8], t4[4], t2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float lo_ = qq ? sv[i] * sv[i] : sv[i], hi_ = qq ? sv[8 + i] * sv[8 + i] : sv[8 + i];
      t8[i] = (b4 ? hi_ : lo_) + __shfl_xor_sync(0xffffffffu, b4 ? lo_ : hi_, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) t4[i] = (b3 ? t8[4 + i] : t8[i]) + __shfl_xor_sync(0xffffffffu, b3 ? t8[i] : t8[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) t2[i] = (b2 ? t4[2 + i] : t4[i]) + __shfl_xor_sync(0xffffffffu, b2 ? t4[i] : t4[2 + i], 4);
    const float t1 = (b1 ? t2[1] : t2[0]) + __shfl_xor_sync(0xffffffffu, b1 ? t2[0] : t2[1], 2);
    tot[qq] = t1 + __shfl_xor_sync(0xffffffffu, t1, 1);
  }
  if (!(lane & 1)) {                           // 16 lanes hold the 16 distinct columns of this chunk
    const int col = c16 + ((lane >> 1) & 1) + ((lane >> 2) & 1) * 2 + ((lane >> 3) & 1) * 4 + ((lane >> 4) & 1) * 8;
    ws[col] += tot[0]; ws[cout + col] += tot[1];
  }
}

// Called by ALL epilogue threads (et = 0 .. 32*NW-1) after their last accumulate16.
// sstat: [NW warps][2][cout] floats; last_flag: a shared-memory int; uses named barrier BAR (32*NW threads).
template <int NW = 4, int BAR = 1>
__device__ __forceinline__ void finalize(const Stats& f, const float* sstat, int cout, int cout_valid, int et,
                                         volatile int* last_flag)
{
  constexpr int NT = 32 * NW;
  auto sync = [] { asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(NT) : "memory"); };
  sync();
  for (int c = et; c < cout_valid; c += NT) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) { s1 += sstat[(size_t)w * 2 * cout + c]; s2 += sstat[(size_t)w * 2 * cout + cout + c]; }
    atomicAdd(f.scratch + 2 * c, (double)s1);
    atomicAdd(f.scratch + 2 * c + 1, (double)s2);
  }
  __threadfence();
  sync();
  if (et == 0) {
    unsigned int* ticket = reinterpret_cast<unsigned int*>(f.scratch + 2 * 256);
    const int last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    *last_flag = last;
    if (last) *ticket = 0u;
  }
  sync();
  if (*last_flag) {
    __threadfence();
    for (int c = et; c < cout_valid; c += NT) {
      const double sum = __ldcg(f.scratch + 2 * c), sq = __ldcg(f.scratch + 2 * c + 1);
      f.scratch[2 * c] = 0.0; f.scratch[2 * c + 1] = 0.0;
      const double mean = sum / (double)f.count;
      double var = sq / (double)f.count - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rs = (float)(1.0 / sqrt(var + (double)f.eps));
      const float g = f.gamma ? f.gamma[c] : 1.f, be = f.beta ? f.beta[c] : 0.f;
      const float av = g * rs;
      f.a[c] = av; f.b[c] = be - (float)mean * av; f.rstd[c] = rs; f.mean[c] = (float)mean;
      if (f.rm) {
        const double unb = f.count > 1 ? var * (double)f.count / (double)(f.count - 1) : var;
        f.rm[c] = (1.f - f.mom) * f.rm[c] + f.mom * (float)mean;
        f.rv[c] = (1.f - f.mom) * f.rv[c] + f.mom * (float)unb;
      }
    }
  }
}

}  // namespace bnepi
