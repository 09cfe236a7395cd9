// Per-element arithmetic of the three FlowNet2 custom ops (the reference's only native code,
// third_party/flownet2/networks/{correlation,resample2d,channelnorm}_package), shared by the CUDA kernels
// (flownet_ops.cu) and the gcc-compiled host check of the CPU test suite.  Forward only: the pipeline runs FlowNet2 for
// inference (flow.py), never through autograd.  NCHW fp32 tensors, as the reference ops take them.
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define CVD_HD __host__ __device__ __forceinline__
#else
#define CVD_HD static inline
#endif

// correlation_cuda_kernel.cu:51-128 (correlation_forward) on zero-padded inputs, written on the unpadded tensors:
//   out[n, (tj+R)*D + (ti+R), y, x] = 1/(K*K*C) * sum_{j,i in [-kr,kr]} sum_c
//        in1[n, c, y*s1 + md - pad + j, x*s1 + md - pad + i] * in2[n, c, (same) + tj*s2, (same) + ti*s2]     (0 outside)
//   R = md / s2, D = 2R + 1, kr = (K - 1) / 2.   FlowNetC: pad = md = 20, K = 1, s1 = 1, s2 = 2  (FlowNetC.py:28-31)
CVD_HD float cvd_correlation_element(const float* in1, const float* in2, long long idx, int C, int H, int W, int Ho, int Wo,
                                     int pad, int K, int md, int s1, int s2)
{
  const int R = md / s2, D = 2 * R + 1, kr = (K - 1) / 2;
  const int x = (int)(idx % Wo); long long t = idx / Wo;
  const int y = (int)(t % Ho); t /= Ho;
  const int tc = (int)(t % (D * D)); const long long n = t / (D * D);
  const int tj = tc / D - R, ti = tc % D - R;
  const int y1 = y * s1 + md - pad, x1 = x * s1 + md - pad;         // centre in UNPADDED coordinates
  const int y2 = y1 + tj * s2, x2 = x1 + ti * s2;
  const size_t hw = (size_t)H * W;
  const float* a = in1 + (size_t)n * C * hw;
  const float* b = in2 + (size_t)n * C * hw;
  float acc = 0.f;
  for (int j = -kr; j <= kr; ++j)
    for (int i = -kr; i <= kr; ++i) {
      const int ya = y1 + j, xa = x1 + i, yb = y2 + j, xb = x2 + i;
      if (ya < 0 || ya >= H || xa < 0 || xa >= W || yb < 0 || yb >= H || xb < 0 || xb >= W) continue;   // zero padding
      const float* pa = a + (size_t)ya * W + xa;
      const float* pb = b + (size_t)yb * W + xb;
      for (int c = 0; c < C; ++c) acc += pa[(size_t)c * hw] * pb[(size_t)c * hw];
    }
  return acc / (float)(K * K * C);
}

// resample2d_kernel.cu:17-73 (kernel_size = 1, bilinear): taps floor / floor+1 clamped to the image, weights from the
// fractional part of the UNclamped coordinate (so out-of-image samples replicate the border with the same weights).
CVD_HD float cvd_resample2d_element(const float* in1, const float* flow, long long idx, int C, int H, int W)
{
  const int x = (int)(idx % W); long long t = idx / W;
  const int y = (int)(t % H); t /= H;
  const int c = (int)(t % C); const long long n = t / C;
  const size_t hw = (size_t)H * W;
  const float dx = flow[((size_t)n * 2 + 0) * hw + (size_t)y * W + x], dy = flow[((size_t)n * 2 + 1) * hw + (size_t)y * W + x];
  const float xf = (float)x + dx, yf = (float)y + dy;
  const float fx = floorf(xf), fy = floorf(yf);
  const float alpha = xf - fx, beta = yf - fy;
  int xL = (int)fx, xR = (int)fx + 1, yT = (int)fy, yB = (int)fy + 1;
  xL = xL < 0 ? 0 : (xL > W - 1 ? W - 1 : xL); xR = xR < 0 ? 0 : (xR > W - 1 ? W - 1 : xR);
  yT = yT < 0 ? 0 : (yT > H - 1 ? H - 1 : yT); yB = yB < 0 ? 0 : (yB > H - 1 ? H - 1 : yB);
  const float* p = in1 + ((size_t)n * C + c) * hw;
  return (1.f - alpha) * (1.f - beta) * p[(size_t)yT * W + xL] + alpha * (1.f - beta) * p[(size_t)yT * W + xR]
       + (1.f - alpha) * beta * p[(size_t)yB * W + xL] + alpha * beta * p[(size_t)yB * W + xR];
}

// channelnorm_kernel.cu:16-60 (norm_deg = 2): out[n, 0, y, x] = sqrt(sum_c in[n, c, y, x]^2)
CVD_HD float cvd_channelnorm_element(const float* in, long long idx, int C, int H, int W)
{
  const size_t hw = (size_t)H * W;
  const long long n = idx / (long long)hw;
  const float* p = in + (size_t)n * C * hw + (size_t)(idx % (long long)hw);
  float acc = 0.f;
  for (int c = 0; c < C; ++c) { const float v = p[(size_t)c * hw]; acc += v * v; }
  return sqrtf(acc);
}
