"""Fused stand-in for the geometry primitives the fine-tuning loss uses.

The reference chains utils/geometry.py pixel_grid:9, pixels_to_rays:38,
pixels_to_points:86, reproject_points:103, project:64 and sample:201 as ~60
separate ATen kernels; here they exist only fused inside
`cvd_consistency_fwd_bwd` (csrc/consistency_loss.cu).  This module exposes that
fused op at the tensor level; there is deliberately no per-primitive GPU op and
no CPU implementation (the CPU restatement lives in oracle/, test-only).
"""
import ctypes as C

import torch

from .. import _lib


class _Workspace:
    """Per-(device,B) scratch reused across calls so the hot loop never allocates."""
    _cache = {}

    @classmethod
    def get(cls, dev, B):
        key = (dev.index, B)
        ws = cls._cache.get(key)
        if ws is None:
            ws = {
                "acc": torch.empty(int(_lib.lib().cvd_consistency_workspace_bytes(B)) // 8 + 2, dtype=torch.float64, device=dev),
                "msum": torch.empty(B * 2, dtype=torch.float32, device=dev),
            }
            cls._cache[key] = ws
        return ws


def fused_consistency(depth, flows, masks, extrinsics, intrinsics, lambda_reprojection, lambda_view_baseline,
                      want_grad=True, B_global=None, f_dir=None, msum=None):
    """One launch of the fused reproject+consistency kernel.

    depth (B,2,H,W); flows = [flow0, flow1] each (B,2,H,W); masks = [m0, m1] each (B,1,H,W);
    extrinsics (B,2,3,4); intrinsics (B,2,4).  All CUDA fp32 contiguous.
    Returns (loss (1,), pair_losses (2,B) [reprojection, disparity], grad_depth (B,2,H,W) or None).
    """
    L = _lib.lib()
    B, N, H, W = depth.shape
    if N != 2:
        raise _lib.CvdError("fused_consistency expects pairs (N == 2)")
    dev = depth.device
    ws = _Workspace.get(dev, B)
    st = _lib.stream()
    f32 = lambda t: t if (t.dtype == torch.float32 and t.is_contiguous()) else t.contiguous().float()
    depth, extrinsics, intrinsics = f32(depth), f32(extrinsics), f32(intrinsics)
    f0, f1, m0, m1 = f32(flows[0]), f32(flows[1]), f32(masks[0]), f32(masks[1])
    if msum is None:
        msum = ws["msum"]
        _lib.check(L.cvd_mask_sums(_lib.ptr(m0), _lib.ptr(m1), B, H, W, _lib.ptr(msum), st), "cvd_mask_sums")
    out_pair = torch.empty((2, B), dtype=torch.float32, device=dev)
    out_loss = torch.empty((1,), dtype=torch.float32, device=dev)
    grad = torch.empty_like(depth) if want_grad else None
    fptr = fdev = None
    if torch.is_tensor(f_dir):                # device-resident (2,) tensor: read by the kernels at run time
        fdev = f_dir
    elif f_dir is not None:
        fptr = (C.c_float * 2)(float(f_dir[0]), float(f_dir[1]))
    _lib.check(L.cvd_consistency_fwd_bwd(
        _lib.ptr(depth), _lib.ptr(f0), _lib.ptr(f1), _lib.ptr(m0), _lib.ptr(m1),
        _lib.ptr(extrinsics), _lib.ptr(intrinsics), _lib.ptr(msum), fptr, _lib.ptr(fdev),
        C.c_float(lambda_reprojection), C.c_float(lambda_view_baseline),
        B, B_global or B, H, W, _lib.ptr(ws["acc"]), _lib.ptr(out_pair), _lib.ptr(out_loss),
        _lib.ptr(grad), st), "cvd_consistency_fwd_bwd")
    return out_loss, out_pair, grad
