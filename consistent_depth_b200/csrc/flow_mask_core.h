// Per-element arithmetic of cvd_flow_consistency_masks (flow_mask.cu), shared by the CUDA kernel and the host-compiled
// check of the CPU test suite.  Element i = ((b * 2 + k) * H + y) * W + x of masks (B, 2 directions, H, W).
#pragma once
#include <math.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define CVD_HD __host__ __device__ __forceinline__
#else
#define CVD_HD static inline
#endif

CVD_HD float cvd_flow_mask_element(const float* flows, const float* colors, long long i, int H, int W, float ft2, float ct2)
{
  const long long hw = (long long)H * W;
  const long long pix = i % hw;
  const int k = (int)((i / hw) & 1);
  const long long b = i / (2 * hw);
  const int x = (int)(pix % W), y = (int)(pix / W);
  const float* fr = flows + ((size_t)(b * 2 + k) * 2) * hw;           // flow of direction k (ref -> tgt)
  const float* ft = flows + ((size_t)(b * 2 + (1 - k)) * 2) * hw;     // flow of the opposite direction
  const float* cr = colors + ((size_t)(b * 2 + k) * 3) * hw;
  const float* ct = colors + ((size_t)(b * 2 + (1 - k)) * 3) * hw;
  const float u = fr[pix], v = fr[hw + pix];
  const float ix = u + (float)x, iy = v + (float)y;
  const int inside = ix >= 0.f && ix <= (float)(W - 1) && iy >= 0.f && iy <= (float)(H - 1);
  // F.grid_sample(align_corners=False, padding_mode="border") of grid = 2 uv / (W, H) - 1  ==  bilinear at uv - 0.5, clamped
  const float sx = fminf(fmaxf(ix - 0.5f, 0.f), (float)(W - 1)), sy = fminf(fmaxf(iy - 0.5f, 0.f), (float)(H - 1));
  const int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
  const int x1 = x0 + 1 < W ? x0 + 1 : W - 1, y1 = y0 + 1 < H ? y0 + 1 : H - 1;
  const float tx = sx - (float)x0, ty = sy - (float)y0;
  const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  const long long p00 = (long long)y0 * W + x0, p01 = (long long)y0 * W + x1, p10 = (long long)y1 * W + x0, p11 = (long long)y1 * W + x1;
  float fsse = 0.f;
  for (int c = 0; c < 2; ++c) {
    const float* pl = ft + (size_t)c * hw;
    const float s = -(w00 * pl[p00] + w01 * pl[p01] + w10 * pl[p10] + w11 * pl[p11]);
    const float d = fr[(size_t)c * hw + pix] - s;
    fsse += d * d;
  }
  float csse = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float* pl = ct + (size_t)c * hw;
    const float s = w00 * pl[p00] + w01 * pl[p01] + w10 * pl[p10] + w11 * pl[p11];
    const float d = cr[(size_t)c * hw + pix] - s;
    csse += d * d;
  }
  return (inside && fsse < ft2 && csse < ct2) ? 1.f : 0.f;
}
