/* TEST INFRASTRUCTURE: the per-element arithmetic of the CUDA kernel flow_mask.cu (csrc/flow_mask_core.h), compiled for
 * the host by tests/test_flowmask_core_cpu.py so that it can be checked against the oracle without a GPU.  Never linked
 * into the product library. */
#include <stddef.h>
#include "flow_mask_core.h"

void flow_mask_host(const float* flows, const float* colors, float* masks, int B, int H, int W, float flow_thresh, float color_thresh)
{
  const long long n = (long long)B * 2 * H * W;
  for (long long i = 0; i < n; ++i)
    masks[i] = cvd_flow_mask_element(flows, colors, i, H, W, flow_thresh * flow_thresh, 3.f * color_thresh * color_thresh);
}
