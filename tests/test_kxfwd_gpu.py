"""GPU: the experimental kx-fused forward convolution of DESIGN.md §8 -- column conv with
N = k*Cout GEMM columns (conv_col.cu) + shifted sum (kx_epilogue.cu) -- against torch's conv2d, and the MC engine with
CVD_KXFWD=1 against the default engine.  Not part of the default suite: the path is off by default and has not been
validated on hardware yet."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rnd(seed, shape, lo=-1.0, hi=1.0):
    return torch.tensor(synth.uniform(seed, 1, shape, lo, hi), device=DEV)


@pytest.mark.parametrize("cin,cout,k,N,H,W", [(64, 16, 11, 2, 32, 48), (32, 32, 7, 1, 20, 28), (64, 32, 3, 2, 16, 24),
                                               (32, 16, 5, 1, 14, 24), (64, 16, 7, 1, 33, 50)])
def test_kx_fused_forward_matches_torch(cin, cout, k, N, H, W):
    from consistent_depth_b200 import ops
    x = rnd(1 + cin, (N, cin, H, W))
    w = rnd(2 + cout, (cout, cin, k, k), -0.1, 0.1)
    bias = rnd(3, (cout,))
    xb = x.permute(0, 2, 3, 1).contiguous()
    yb = torch.full((N, H, W, cout + 8), 7.0, device=DEV)
    bufs = ops.kxfwd_buffers(cin, cout, k, N, H, W, 3, DEV)
    ops.conv_kxfwd(ops.make_src(ops.View(xb, 0)), w, bias, ops.View(yb, 4), N, H, W, cin, cout, k, 3, bufs)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
    got = yb[..., 4:4 + cout].double()
    assert (got - ref).abs().max() <= 6e-5 * ref.abs().max()
    assert (yb[..., :4] == 7).all() and (yb[..., 4 + cout:] == 7).all()


def test_mc_engine_with_kx_fused_forward(monkeypatch):
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel, default_init_state
    sd = default_init_state(3)
    img = torch.tensor(synth.uniform(9, 1, (1, 2, 3, 32, 48), 0, 1), device=DEV)
    base = MannequinChallengeModel(state_dict=sd)
    with torch.no_grad():
        d0 = base(img).cpu().numpy()
    monkeypatch.setenv("CVD_KXFWD", "1")
    kx = MannequinChallengeModel(state_dict=sd)
    with torch.no_grad():
        d1 = kx(img).cpu().numpy()
    assert kx.engine(2, 32, 48).kxfwd
    np.testing.assert_allclose(d1, d0, rtol=1e-3)
