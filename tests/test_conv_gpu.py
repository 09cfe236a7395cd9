"""GPU: tcgen05 implicit-GEMM conv (fwd + dgrad) vs torch fp64 conv on the same inputs.

Tolerances: precision 3 (bf16x3 split, the parity mode) max-abs error <= 2e-5 of the output's
max magnitude; precision 1 (plain bf16) <= 2e-2.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def rnd(seed, shape, lo=-1.0, hi=1.0):
    return torch.tensor(synth.uniform(seed, 1, shape, lo, hi), device=DEV)


def run_conv(x_nchw, w, bias, k, precision=3, relu=False, a=None, b=None, flags=0, ctot_pad=0, out_pad=0, y_init=None):
    from consistent_depth_b200 import ops
    N, cin, H, W = x_nchw.shape
    cout = w.shape[0]
    xb = torch.zeros(N, H, W, cin + ctot_pad, device=DEV)
    xb[..., ctot_pad // 2: ctot_pad // 2 + cin] = nhwc(x_nchw) if ctot_pad % 8 == 0 else 0
    xv = ops.View(xb, ctot_pad // 2)
    yb = torch.full((N, H, W, cout + out_pad), 7.0, device=DEV) if y_init is None else y_init
    yv = ops.View(yb, out_pad // 2)
    A = B = None
    if a is not None:
        A = torch.zeros(cin + ctot_pad, device=DEV); A[ctot_pad // 2: ctot_pad // 2 + cin] = a
        B = torch.zeros(cin + ctot_pad, device=DEV); B[ctot_pad // 2: ctot_pad // 2 + cin] = b
    pk = ops.pack_weights(w, False, precision)
    ops.conv(ops.make_src(xv, A, B, relu), pk, bias, ops.make_dst(yv), N, H, W, cin, cout, k, precision, flags)
    torch.cuda.synchronize()
    return yb


SHAPES = [
    # cin, cout, k, N, H, W
    (64, 16, 11, 1, 32, 48),
    (32, 32, 3, 2, 32, 48),
    (64, 64, 7, 1, 16, 24),
    (32, 64, 5, 1, 20, 28),          # H, W not multiples of the 16x8 M-tile
    (128, 208, 1, 1, 32, 48),        # fused inception 1x1 (16+64*3), two channel groups
    (256, 160, 1, 2, 16, 24),
    (64, 32, 11, 1, 14, 24),         # smallest hourglass level
]


@pytest.mark.parametrize("cin,cout,k,N,H,W", SHAPES)
@pytest.mark.parametrize("precision", [3, 1])
def test_conv_forward_matches_torch(cin, cout, k, N, H, W, precision):
    x = rnd(10 + cin + k, (N, cin, H, W))
    w = rnd(20 + cout + k, (cout, cin, k, k), -0.1, 0.1)
    bias = rnd(30 + cout, (cout,))
    y = run_conv(x, w, bias, k, precision)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=(k - 1) // 2)
    err = (y.permute(0, 3, 1, 2).double() - ref).abs().max().item()
    tol = (6e-5 if precision == 3 else 2e-2) * ref.abs().max().item()   # fp32 TMEM accumulation truncates: ~K/16*3 steps x 2^-24
    assert err <= tol, (err, tol)


def test_conv_affine_relu_on_load_and_views():
    cin, cout, k, N, H, W = 32, 48, 3, 1, 32, 40
    x = rnd(1, (N, cin, H, W)); w = rnd(2, (cout, cin, k, k), -0.1, 0.1); bias = rnd(3, (cout,))
    a = rnd(4, (cin,), 0.5, 1.5); b = rnd(5, (cin,), -0.5, 0.5)
    y = run_conv(x, w, bias, k, 3, relu=True, a=a, b=b, ctot_pad=16, out_pad=8)
    xin = F.relu(x.double() * a.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1))
    ref = F.conv2d(xin, w.double(), bias.double(), padding=1)
    got = y[..., 4:4 + cout].permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max().item() <= 6e-5 * ref.abs().max().item()
    assert (y[..., :4] == 7.0).all() and (y[..., 4 + cout:] == 7.0).all()      # padding channels untouched


def test_conv_image_first_layer_and_pred_layer_exp():
    from consistent_depth_b200 import ops
    N, H, W = 2, 32, 48
    img = rnd(7, (N, 3, H, W), 0.0, 1.0)
    w = rnd(8, (128, 3, 7, 7), -0.08, 0.08); bias = rnd(9, (128,))
    xb = torch.zeros(N, H, W, 4, device=DEV); xb[..., :3] = nhwc(img)
    yb = torch.empty(N, H, W, 128, device=DEV)
    ops.conv(ops.make_src(ops.View(xb)), ops.pack_weights(w, False, 3), bias, ops.make_dst(ops.View(yb)), N, H, W, 3, 128, 7, 3)
    ref = F.conv2d(img.double(), w.double(), bias.double(), padding=3)
    assert (yb.permute(0, 3, 1, 2).double() - ref).abs().max().item() <= 6e-5 * ref.abs().max().item()
    # pred layer: 64 -> 1, exp epilogue, C=1 destination
    f = rnd(10, (N, 64, H, W)); wp = rnd(11, (1, 64, 3, 3), -0.05, 0.05); bp = rnd(12, (1,))
    fb = nhwc(f); db = torch.empty(N, H, W, 1, device=DEV)
    ops.conv(ops.make_src(ops.View(fb)), ops.pack_weights(wp, False, 3), bp, ops.make_dst(ops.View(db)), N, H, W, 64, 1, 3, 3, ops.FLAG_EXP)
    refd = torch.exp(F.conv2d(f.double(), wp.double(), bp.double(), padding=1))
    assert (db.permute(0, 3, 1, 2).double() - refd).abs().max().item() <= 6e-5 * refd.abs().max().item()


def test_conv_dgrad_matches_autograd_and_accumulates():
    from consistent_depth_b200 import ops
    cin, cout, k, N, H, W = 32, 64, 5, 1, 32, 24
    x = rnd(1, (N, cin, H, W)).double().requires_grad_(True)
    w = rnd(2, (cout, cin, k, k), -0.1, 0.1)
    gy = rnd(3, (N, cout, H, W))
    F.conv2d(x, w.double(), None, padding=2).backward(gy.double())
    gyb = nhwc(gy)
    dxb = torch.full((N, H, W, cin), 1.0, device=DEV)
    pk = ops.pack_weights(w, True, 3)
    ops.conv(ops.make_src(ops.View(gyb)), pk, None, ops.make_dst(ops.View(dxb)), N, H, W, cout, cin, k, 3, ops.FLAG_ACCUM)
    torch.cuda.synchronize()
    got = dxb.permute(0, 3, 1, 2).double() - 1.0
    assert (got - x.grad).abs().max().item() <= 6e-5 * x.grad.abs().max().item()


def test_conv_bn_backward_on_load():
    """dgrad with the BatchNorm+ReLU backward of the conv OUTPUT applied while staging."""
    from consistent_depth_b200 import ops
    cin, cout, k, N, H, W = 32, 32, 3, 1, 16, 24
    w = rnd(2, (cout, cin, k, k), -0.1, 0.1)
    xraw = rnd(3, (N, cout, H, W))            # the conv's raw output (pre-BN)
    dy = rnd(4, (N, cout, H, W))              # gradient wrt post BN+ReLU activation
    a = rnd(5, (cout,), 0.5, 1.5); b = rnd(6, (cout,), -0.3, 0.3)
    c0 = rnd(7, (cout,), 0.5, 1.5); c1 = rnd(8, (cout,), -0.1, 0.1); c2 = rnd(9, (cout,), -0.1, 0.1)
    bw = torch.stack([c0, c1, c2, torch.zeros_like(c0)], 1).contiguous()
    yv = xraw.double() * a.double().view(1, -1, 1, 1) + b.double().view(1, -1, 1, 1)
    g = torch.where(yv > 0, dy.double(), torch.zeros_like(yv))
    dxraw = c0.double().view(1, -1, 1, 1) * g - c1.double().view(1, -1, 1, 1) - c2.double().view(1, -1, 1, 1) * yv
    ref = F.conv_transpose2d(dxraw, w.double(), padding=1)
    dxb = torch.empty(N, H, W, cin, device=DEV)
    src = ops.make_src(ops.View(nhwc(xraw)), a, b, True, dy=ops.View(nhwc(dy)), bw=bw)
    ops.conv(src, ops.pack_weights(w, True, 3), None, ops.make_dst(ops.View(dxb)), N, H, W, cout, cin, k, 3)
    torch.cuda.synchronize()
    assert (dxb.permute(0, 3, 1, 2).double() - ref).abs().max().item() <= 6e-5 * ref.abs().max().item()


def test_conv_full_size_layer_linearity():
    """BASELINE config-2 size (8 frames 224x384, the heaviest layer 64->16 11x11): conv(x1+x2) == conv(x1)+conv(x2)
    and agreement with torch (fp64) on a top-left probe crop."""
    cin, cout, k, N, H, W = 64, 16, 11, 8, 224, 384
    g = torch.Generator(device=DEV).manual_seed(0)
    x1 = torch.rand(N, cin, H, W, device=DEV, generator=g) - 0.5
    x2 = torch.rand(N, cin, H, W, device=DEV, generator=g) - 0.5
    w = (torch.rand(cout, cin, k, k, device=DEV, generator=g) - 0.5) * 0.05
    y1 = run_conv(x1, w, None, k); y2 = run_conv(x2, w, None, k); y12 = run_conv(x1 + x2, w, None, k)
    assert (y12 - (y1 + y2)).abs().max().item() <= 1e-4 * y12.abs().max().item()
    ref = F.conv2d(x1[7:, :, :40, :56].double(), w.double(), None, padding=5)[:, :, :32, :48]
    got = y1[7:, :32, :48].permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max().item() <= 6e-5 * ref.abs().max().item()


WG_SHAPES = [
    # cin, cout, k, N, H, W
    (64, 16, 11, 1, 32, 48),       # M = Cin(64), N = Cout(16), 6 tap groups
    (32, 32, 3, 2, 32, 32),        # M=64 padded from 32
    (32, 64, 5, 1, 20, 28),        # M = Cout, ragged H/W
    (64, 64, 7, 1, 16, 32),        # 8 taps per group
    (128, 208, 1, 1, 32, 48),      # fused 1x1: M = Cout(208 -> 2 x 128), N = Cin
    (256, 160, 1, 1, 16, 32),      # M = Cin (2 x 128)
    (3, 128, 7, 1, 32, 48),        # first layer (image, 4-channel tensor)
    (64, 1, 3, 1, 32, 48),         # pred layer (G has 1 channel in a 4-channel tensor)
]


@pytest.mark.parametrize("cin,cout,k,N,H,W", WG_SHAPES)
def test_conv_wgrad_matches_autograd(cin, cout, k, N, H, W):
    from consistent_depth_b200 import ops
    x = rnd(1 + cin, (N, cin, H, W))
    gy = rnd(2 + cout, (N, cout, H, W))
    w = torch.zeros(cout, cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, None, padding=(k - 1) // 2).backward(gy.double())
    cx, cg = (cin + 3) // 4 * 4, (cout + 3) // 4 * 4
    xb = torch.zeros(N, H, W, cx, device=DEV); xb[..., :cin] = nhwc(x)
    gb = torch.zeros(N, H, W, cg, device=DEV); gb[..., :cout] = nhwc(gy)
    dw = torch.zeros(cout, cin, k, k, device=DEV)
    ops.conv_wgrad(ops.make_src(ops.View(gb)), ops.make_src(ops.View(xb)), dw, N, H, W, cin, cout, k, 3)
    torch.cuda.synchronize()
    err = (dw.double() - w.grad).abs().max().item()
    assert err <= 6e-5 * w.grad.abs().max().item(), (err, w.grad.abs().max().item())


def test_conv_wgrad_with_transforms_on_load():
    from consistent_depth_b200 import ops
    cin, cout, k, N, H, W = 32, 32, 3, 1, 16, 32
    xin_raw = rnd(3, (N, cin, H, W)); ax = rnd(4, (cin,), 0.5, 1.5); bx = rnd(5, (cin,), -0.5, 0.5)
    xraw = rnd(6, (N, cout, H, W)); dy = rnd(7, (N, cout, H, W))
    a = rnd(8, (cout,), 0.5, 1.5); b = rnd(9, (cout,), -0.3, 0.3)
    c0 = rnd(10, (cout,), 0.5, 1.5); c1 = rnd(11, (cout,), -0.1, 0.1); c2 = rnd(12, (cout,), -0.1, 0.1)
    bw = torch.stack([c0, c1, c2, torch.zeros_like(c0)], 1).contiguous()
    v = lambda t: t.double().view(1, -1, 1, 1)
    xin = F.relu(xin_raw.double() * v(ax) + v(bx))
    yv = xraw.double() * v(a) + v(b)
    g = torch.where(yv > 0, dy.double(), torch.zeros_like(yv))
    dxraw = v(c0) * g - v(c1) - v(c2) * yv
    w = torch.zeros(cout, cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, w, None, padding=1).backward(dxraw)
    dw = torch.zeros(cout, cin, k, k, device=DEV)
    gsrc = ops.make_src(ops.View(nhwc(xraw)), a, b, True, dy=ops.View(nhwc(dy)), bw=bw)
    xsrc = ops.make_src(ops.View(nhwc(xin_raw)), ax, bx, True)
    ops.conv_wgrad(gsrc, xsrc, dw, N, H, W, cin, cout, k, 3)
    torch.cuda.synchronize()
    assert (dw.double() - w.grad).abs().max().item() <= 6e-5 * w.grad.abs().max().item()


@pytest.mark.parametrize("cin,cout,k,N,H,W,off", [
    (128, 208, 1, 2, 32, 48, 0),       # fused inception 1x1: statistics of all 208 columns
    (64, 16, 11, 2, 32, 48, 24),       # k x k conv writing a channel slice at offset 24 of a wider buffer
    (32, 64, 5, 1, 20, 28, 8),         # ragged tiles: out-of-image accumulator rows must not count
    (3, 128, 7, 1, 32, 48, 0),         # conv1 with affine gamma / beta
])
def test_conv_fused_batchnorm_statistics(cin, cout, k, N, H, W, off):
    """cvd_conv_fwd_bn == cvd_conv_fwd followed by nn.BatchNorm2d(train) statistics (hourglass.py:28,40,43,165):
    a = gamma*rstd, b = beta - mean*a, running stats with momentum 0.1 and unbiased variance; scratch self-cleans."""
    from consistent_depth_b200 import ops
    x = rnd(70 + cin, (N, cin, H, W))
    w = rnd(71 + cout, (cout, cin, k, k), -0.1, 0.1)
    bias = rnd(72, (cout,))
    gamma, beta = rnd(73, (cout,), 0.5, 1.5), rnd(74, (cout,))
    ct = cout + off + 8
    xb = nhwc(x) if cin % 8 == 0 else torch.cat([nhwc(x), torch.zeros(N, H, W, 4 - cin, device=DEV)], -1).contiguous()
    yb = torch.zeros(N, H, W, ct, device=DEV)
    a, b, rstd, mean = (torch.full((ct,), 9.0, device=DEV) for _ in range(4))
    rm, rv = rnd(75, (cout,)), rnd(76, (cout,), 0.5, 2.0)
    rm0, rv0 = rm.clone(), rv.clone()
    scratch = ops.bn_scratch(DEV)
    pk = ops.pack_weights(w, False, 3)
    bn = ops.make_bn(scratch, a, b, rstd, mean, gamma, beta, rm, rv)
    for _ in range(2):       # twice: the scratch must come back zeroed
        ops.conv(ops.make_src(ops.View(xb, 0)), pk, bias, ops.make_dst(ops.View(yb, off)), N, H, W, cin, cout, k, 3, 0, bn=bn)
    torch.cuda.synchronize()
    y = yb[..., off:off + cout].double()
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
    assert (y - ref).abs().max() <= 6e-5 * ref.abs().max()
    m, v = y.mean((0, 1, 2)), y.var((0, 1, 2), unbiased=False)
    rs = 1.0 / torch.sqrt(v + 1e-5)
    sl = slice(off, off + cout)
    assert torch.allclose(mean[sl].double(), m, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rstd[sl].double(), rs, rtol=2e-5)
    assert torch.allclose(a[sl].double(), gamma.double() * rs, rtol=2e-5)
    assert torch.allclose(b[sl].double(), beta.double() - m * gamma.double() * rs, rtol=1e-4, atol=1e-5)
    assert (a[:off] == 9).all() and (a[off + cout:] == 9).all()          # neighbours untouched
    n = N * H * W
    rm_ref, rv_ref = rm0.double(), rv0.double()
    for _ in range(2):
        rm_ref = 0.9 * rm_ref + 0.1 * m
        rv_ref = 0.9 * rv_ref + 0.1 * v * n / (n - 1)
    assert torch.allclose(rm.double(), rm_ref, rtol=1e-5, atol=1e-6)
    assert torch.allclose(rv.double(), rv_ref, rtol=2e-5)
    assert (scratch == 0).all()
