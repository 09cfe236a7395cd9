// Flow / photometric consistency masks: utils/consistency.py:8-67 (called by
// flow.py:199-228 mask_valid_correspondences) as ONE kernel per frame pair, both directions.
//   mask_k = inside(x + u_k, y + v_k)  AND  |flow_k - sample(-flow_{1-k})|^2 < flow_thresh^2
//                                      AND  |color_k - sample(color_{1-k})|^2 < 3 color_thresh^2
// sample = bilinear at (x + u - 0.5, y + v - 0.5), border-clamped (grid = 2 uv / (W, H) - 1 and F.grid_sample with
// align_corners=False, padding_mode="border": utils/consistency.py:8-24 -- note the (W, H) normalisation, unlike
// utils/geometry.py:201-208 which divides by (W-1, H-1)).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cvd.h"
#include "cvd_common.cuh"

// The per-pixel arithmetic lives in flow_mask_core.h as a __host__ __device__ function so that the CPU test suite can
// compile it with gcc and check it against the oracle without a GPU (tests/test_flowmask_core_cpu.py).
#include "flow_mask_core.h"

namespace {

// planar tensors: flows (B, 2 dirs, 2, H, W), colors (B, 2 frames, 3, H, W), masks (B, 2 dirs, H, W) float {0,1}
__global__ void flow_mask_kernel(const float* __restrict__ flows, const float* __restrict__ colors, float* __restrict__ masks,
                                 int B, int H, int W, float ft2, float ct2)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * 2 * H * W) return;
  masks[i] = cvd_flow_mask_element(flows, colors, i, H, W, ft2, ct2);
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
