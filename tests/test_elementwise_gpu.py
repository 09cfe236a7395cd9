"""GPU: BatchNorm statistics / backward reductions, pool, upsample-merge glue kernels vs torch (fp64)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rnd(seed, shape, lo=-1.0, hi=1.0):
    return torch.tensor(synth.uniform(seed, 1, shape, lo, hi), device=DEV)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2)


@pytest.mark.parametrize("C,ct,off,affine", [(32, 32, 0, False), (208, 256, 0, False), (48, 96, 32, False), (128, 128, 0, True),
                                                 (512, 512, 0, True)])      # > 256 channels: one chunked launch (blockIdx.y)
def test_bn_stats_and_backward_match_torch_batchnorm(C, ct, off, affine):
    from consistent_depth_b200 import ops
    N, H, W = 2, 24, 20
    x = rnd(1, (N, C, H, W), -2, 3)
    dy = rnd(2, (N, C, H, W))
    gamma = rnd(3, (C,), 0.5, 1.5) if affine else None
    beta = rnd(4, (C,), -0.5, 0.5) if affine else None
    rm = torch.zeros(C, device=DEV); rv = torch.ones(C, device=DEV)
    xb = torch.zeros(N, H, W, ct, device=DEV); xb[..., off:off + C] = nhwc(x)
    dyb = torch.zeros(N, H, W, ct, device=DEV); dyb[..., off:off + C] = nhwc(dy)
    a = torch.zeros(ct, device=DEV); b = torch.zeros(ct, device=DEV); rstd = torch.zeros(ct, device=DEV); mean = torch.zeros(ct, device=DEV)
    scratch = ops.bn_scratch(DEV, C)
    npix = N * H * W
    ops.bn_stats(xb, off, C, npix, scratch, a, b, rstd, mean, gamma, beta, rm, rv)
    # torch reference (fp64), train-mode BatchNorm2d + ReLU
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True) if affine else None
    bd = beta.double().requires_grad_(True) if affine else None
    rmd = torch.zeros(C, device=DEV, dtype=torch.float64); rvd = torch.ones(C, device=DEV, dtype=torch.float64)
    y = F.relu(F.batch_norm(xd, rmd, rvd, gd, bd, True, 0.1, 1e-5))
    y.backward(dy.double())
    torch.cuda.synchronize()
    yk = F.relu(x.double() * a[off:off + C].double().view(1, -1, 1, 1) + b[off:off + C].double().view(1, -1, 1, 1))
    assert (yk - y.detach()).abs().max().item() <= 1e-5
    torch.testing.assert_close(rm.double(), rmd, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rv.double(), rvd, rtol=1e-5, atol=1e-6)
    bw = torch.zeros(ct, 4, device=DEV)
    dg = torch.zeros(C, device=DEV) if affine else None
    db = torch.zeros(C, device=DEV) if affine else None
    dbias = torch.zeros(C, device=DEV)
    ops.bn_bwd_reduce(xb, off, C, dyb, npix, scratch, a, b, rstd, mean, bw, True, gamma, beta, dg, db, dbias)
    torch.cuda.synchronize()
    yv = x.double() * a[off:off + C].double().view(1, -1, 1, 1) + b[off:off + C].double().view(1, -1, 1, 1)
    g = torch.where(yv > 0, dy.double(), torch.zeros_like(yv))
    c = bw[off:off + C].double()
    dx = c[:, 0].view(1, -1, 1, 1) * g - c[:, 1].view(1, -1, 1, 1) - c[:, 2].view(1, -1, 1, 1) * yv
    assert (dx - xd.grad).abs().max().item() <= 2e-5 * xd.grad.abs().max().item()
    if affine:
        torch.testing.assert_close(dg.double(), gd.grad, rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(db.double(), bd.grad, rtol=1e-4, atol=1e-4)
    assert dbias.abs().max().item() <= 1e-3 * xd.grad.abs().sum().item() + 1e-3   # mathematically zero


def test_pool_and_merge_up_forward_backward():
    from consistent_depth_b200 import ops
    N, H, W, C = 2, 16, 24, 32
    x1 = rnd(1, (N, C, H, W)); x2 = rnd(2, (N, C, H // 2, W // 2))
    a1 = rnd(3, (48,), 0.5, 1.5); b1 = rnd(4, (48,), -0.5, 0.5)
    a2 = rnd(5, (C,), 0.5, 1.5); b2 = rnd(6, (C,), -0.5, 0.5)
    # x1 lives in a 48-channel buffer through a gapped view: logical [0,16) -> phys [0,16), [16,32) -> phys [32,48)
    x1b = torch.zeros(N, H, W, 48, device=DEV)
    x1b[..., :16] = nhwc(x1)[..., :16]; x1b[..., 32:] = nhwc(x1)[..., 16:]
    v1 = ops.View(x1b, 0, 16, 16)
    idx = torch.cat([torch.arange(16), torch.arange(32, 48)]).to(DEV)
    z = torch.empty(N, H, W, C, device=DEV)
    ops.merge_up_fwd(v1, a1, b1, ops.View(nhwc(x2)), a2, b2, z, N, H, W, C)
    x1d = x1.double().requires_grad_(True); x2d = x2.double().requires_grad_(True)
    y1 = F.relu(x1d * a1[idx].double().view(1, -1, 1, 1) + b1[idx].double().view(1, -1, 1, 1))
    y2 = F.relu(x2d * a2.double().view(1, -1, 1, 1) + b2.double().view(1, -1, 1, 1))
    ref = y1 + F.interpolate(y2, scale_factor=2, mode="bilinear", align_corners=True)
    torch.cuda.synchronize()
    assert (nchw(z).double() - ref.detach()).abs().max().item() <= 1e-5
    dz = rnd(7, (N, C, H, W))
    up = F.interpolate(y2.detach().requires_grad_(True), scale_factor=2, mode="bilinear", align_corners=True)
    y2r = y2.detach().requires_grad_(True)
    F.interpolate(y2r, scale_factor=2, mode="bilinear", align_corners=True).backward(dz.double())
    dy2 = torch.empty(N, H // 2, W // 2, C, device=DEV)
    dzb = nhwc(dz)
    dy1b = torch.full((N, H, W, 48), 3.0, device=DEV)
    ops.merge_up_bwd(dzb, ops.View(dy2), ops.View(dy1b, 0, 16, 16), True, N, H, W, C)
    torch.cuda.synchronize()
    assert (nchw(dy2).double() - y2r.grad).abs().max().item() <= 1e-5
    assert torch.equal(dy1b[..., idx], dzb + 3.0) and (dy1b[..., 16:32] == 3.0).all()
    # pool
    p = torch.empty(N, H // 2, W // 2, C, device=DEV)
    ops.pool_fwd(v1, a1, b1, True, p, N, H, W, C)
    refp = F.avg_pool2d(y1.detach(), 2)
    torch.cuda.synchronize()
    assert (nchw(p).double() - refp).abs().max().item() <= 1e-5
    dp = rnd(8, (N, C, H // 2, W // 2))
    dx = torch.full((N, H, W, C), 2.0, device=DEV)
    ops.pool_bwd(nhwc(dp), ops.View(dx), True, N, H, W, C)
    yy = y1.detach().requires_grad_(True)
    F.avg_pool2d(yy, 2).backward(dp.double())
    torch.cuda.synchronize()
    assert (nchw(dx).double() - 2.0 - yy.grad).abs().max().item() <= 1e-6


def test_image_layout_and_dlogdepth():
    from consistent_depth_b200 import ops
    N, H, W = 2, 8, 12
    img = rnd(1, (N, 3, H, W), 0, 1)
    out = torch.empty(N, H, W, 4, device=DEV)
    ops.image_to_nhwc4(img, out, N, H, W)
    torch.cuda.synchronize()
    assert torch.equal(out[..., :3], nhwc(img)) and (out[..., 3] == 0).all()
    gd = rnd(2, (N, H, W)); d = rnd(3, (N, H, W), 0.5, 2.0)
    o4 = torch.empty(N, H, W, 4, device=DEV); dbias = torch.zeros(1, device=DEV)
    ops.dlogdepth(gd, d, o4, dbias)
    torch.cuda.synchronize()
    torch.testing.assert_close(o4[..., 0], gd * d)
    torch.testing.assert_close(dbias[0], (gd * d).sum(), rtol=1e-4, atol=1e-4)
