// Flow / photometric consistency masks: utils/consistency.py:8-67 (called by
// flow.py:199-228 mask_valid_correspondences) as ONE kernel per frame pair, both directions.
//   mask_k = inside(x + u_k, y + v_k)  AND  |flow_k - sample(-flow_{1-k})|^2 < flow_thresh^2
//                                      AND  |color_k - sample(color_{1-k})|^2 < 3 color_thresh^2
// sample = bilinear at (x + u - 0.5, y + v - 0.5), border-clamped (grid = 2 uv / (W, H) - 1 and F.grid_sample with
// align_corners=False, padding_mode="border": utils/consistency.py:8-24 -- note the (W, H) normalisation, unlike
// utils/geometry.py:201-208 which divides by (W-1, H-1)).
// NOT YET VALIDATED ON HARDWARE (written after the round's GPU budget was spent): tests/test_flowmask_gpu.py is opt-in.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"

namespace {

// planar tensors: flows (B, 2 dirs, 2, H, W), colors (B, 2 frames, 3, H, W), masks (B, 2 dirs, H, W) float {0,1}
__global__ void flow_mask_kernel(const float* __restrict__ flows, const float* __restrict__ colors, float* __restrict__ masks,
                                 int B, int H, int W, float ft2, float ct2)
{
  const long long hw = (long long)H * W;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 2 * hw) return;
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
  bool ok = ix >= 0.f && ix <= (float)(W - 1) && iy >= 0.f && iy <= (float)(H - 1);
  const float sx = fminf(fmaxf(ix - 0.5f, 0.f), (float)(W - 1)), sy = fminf(fmaxf(iy - 0.5f, 0.f), (float)(H - 1));
  const int x0 = (int)floorf(sx), y0 = (int)floorf(sy);
  const int x1 = min(x0 + 1, W - 1), y1 = min(y0 + 1, H - 1);
  const float tx = sx - (float)x0, ty = sy - (float)y0;
  const float w00 = (1.f - tx) * (1.f - ty), w01 = tx * (1.f - ty), w10 = (1.f - tx) * ty, w11 = tx * ty;
  const long long p00 = (long long)y0 * W + x0, p01 = (long long)y0 * W + x1, p10 = (long long)y1 * W + x0, p11 = (long long)y1 * W + x1;
  float fsse = 0.f;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const float* pl = ft + (size_t)c * hw;
    const float s = -(w00 * __ldg(pl + p00) + w01 * __ldg(pl + p01) + w10 * __ldg(pl + p10) + w11 * __ldg(pl + p11));
    const float d = fr[(size_t)c * hw + pix] - s;
    fsse += d * d;
  }
  float csse = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* pl = ct + (size_t)c * hw;
    const float s = w00 * __ldg(pl + p00) + w01 * __ldg(pl + p01) + w10 * __ldg(pl + p10) + w11 * __ldg(pl + p11);
    const float d = cr[(size_t)c * hw + pix] - s;
    csse += d * d;
  }
  ok = ok && fsse < ft2 && csse < ct2;
  masks[i] = ok ? 1.f : 0.f;
}

}  // namespace

extern "C" int cvd_flow_consistency_masks(const float* flows, const float* colors, float* masks, int B, int H, int W,
                                          float flow_thresh, float color_thresh, void* stream)
{
  CVD_CHECK_ARG(flows && colors && masks && B > 0 && H > 0 && W > 0, "cvd_flow_consistency_masks: bad arguments");
  const long long n = (long long)B * 2 * H * W;
  flow_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(flows, colors, masks, B, H, W,
                                                                                  flow_thresh * flow_thresh,
                                                                                  3.f * color_thresh * color_thresh);
  CVD_LAUNCH_OK("flow_mask_kernel");
  return 0;
}
