"""ConsistencyLoss with the reference's call signature, backed by ONE fused CUDA kernel.

Mirrors loss/consistency_loss.py:91-253: `ConsistencyLoss(opt)(depths, metadata)`
returns `(mean-over-pairs loss, {"reprojection": (B,), "disparity": (B,)})`.
Forward and backward are computed together by `cvd_consistency_fwd_bwd`; the
autograd.Function just hands the precomputed dL/d depth back to autograd.
"""
import torch

from ..utils.geometry import fused_consistency


class _FusedConsistencyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, depths, flows0, flows1, masks0, masks1, extrinsics, intrinsics, lam_r, lam_b, B_global, f_dir):
        need = depths.requires_grad
        loss, pair, grad = fused_consistency(depths.detach(), (flows0, flows1), (masks0, masks1), extrinsics,
                                             intrinsics, lam_r, lam_b, want_grad=need, B_global=B_global, f_dir=f_dir)
        ctx.grad = grad
        ctx.mark_non_differentiable(pair)
        return loss, pair

    @staticmethod
    def backward(ctx, g_loss, _g_pair):
        g = ctx.grad * g_loss.reshape(()) if ctx.grad is not None else None
        return (g,) + (None,) * 10


class ConsistencyLoss(torch.nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.opt = opt

    def __call__(self, depths, metadata):
        """depths (B,2,H,W); metadata as produced by loaders/video_dataset.py:131-207 (collated)."""
        geom = metadata["geometry_consistency"]
        lam_r = float(self.opt.lambda_reprojection)
        lam_b = float(self.opt.lambda_view_baseline)
        loss, pair = _FusedConsistencyFn.apply(
            depths, geom["flows"][0], geom["flows"][1], geom["masks"][0], geom["masks"][1],
            metadata["extrinsics"], metadata["intrinsics"], lam_r, lam_b,
            geom.get("B_global"), geom.get("f_global"))
        batch_losses = {"reprojection": pair[0], "disparity": pair[1]}
        return loss.reshape(()), batch_losses
