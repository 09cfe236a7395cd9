"""Flow / photometric consistency masks — same surface as the reference's utils/consistency.py:53-67
(`consistent_flow_masks(flows, colors, flow_thresh, color_thresh)` on numpy (H,W,2) / (H,W,3) arrays, list of two
bool masks back), computed by one CUDA kernel (cvd_flow_consistency_masks) instead of numpy + F.grid_sample on the CPU.
`consistent_flow_masks_batched` is the tensor-level entry for a whole clip resident on the GPU.
Validated on a B200 against the oracle and the reference-generated golden masks (tests/test_flowmask_gpu.py)."""
import ctypes as C

import numpy as np
import torch

from .. import _lib


def consistent_flow_masks_batched(flows, colors, flow_thresh, color_thresh):
    """flows (B,2,2,H,W), colors (B,2,3,H,W) CUDA fp32 -> masks (B,2,H,W) fp32 {0,1}."""
    B, _, _, H, W = flows.shape
    assert colors.shape == (B, 2, 3, H, W)
    masks = torch.empty(B, 2, H, W, device=flows.device)
    # keep the contiguous copies alive across the launch: a temporary created inside the argument list is returned to the
    # caching allocator before the next argument is evaluated, and `colors.contiguous()` then reuses the block `flows` was
    # copied to (found on hardware: masks computed from overwritten flows)
    flows, colors = flows.contiguous(), colors.contiguous()
    _lib.check(_lib.lib().cvd_flow_consistency_masks(_lib.ptr(flows), _lib.ptr(colors), _lib.ptr(masks),
                                                     B, H, W, C.c_float(flow_thresh), C.c_float(color_thresh), _lib.stream()),
               "cvd_flow_consistency_masks")
    return masks


def consistent_flow_masks(flows, colors, flow_thresh, color_thresh):
    if not torch.cuda.is_available():
        raise RuntimeError("consistent_flow_masks (consistent_depth_b200) needs a CUDA device: there is no CPU path")
    dev = torch.device("cuda", torch.cuda.current_device())
    f = torch.tensor(np.stack([np.asarray(x, np.float32).transpose(2, 0, 1) for x in flows])[None], device=dev)
    c = torch.tensor(np.stack([np.asarray(x, np.float32).reshape(x.shape[0], x.shape[1], -1).transpose(2, 0, 1) for x in colors])[None], device=dev)
    m = consistent_flow_masks_batched(f, c, flow_thresh, color_thresh)[0]
    return [m[0].cpu().numpy() > 0.5, m[1].cpu().numpy() > 0.5]
