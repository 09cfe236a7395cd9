"""GPU: the second-generation conv path -- cvd_prep_operand (per-channel transform + bf16 hi/lo split into chunk-planar
planes) and cvd_conv2_fwd (TMA-fed, kx-fused tcgen05 conv, forward and input-gradient) -- against torch fp64.

Tolerances: operand planes hi + lo reproduce the fp32 value to 2^-16 relative; convolution max-abs error <= 6e-5 of the
output's max magnitude (bf16x3 split products, fp32 TMEM accumulation), the bar of the first-generation kernel."""
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


def z_to_nchw(z, C, H, W):
    """planes [2][N][c8][HW][8] bf16 -> fp32 (N, C, H, W) value hi + lo."""
    v = z[0].float() + z[1].float()                          # (N, c8, HW, 8)
    N, c8, HW, _ = v.shape
    return v.permute(0, 1, 3, 2).reshape(N, c8 * 8, H, W)[:, :C]


def test_prep_operand_affine_and_bnbwd():
    from consistent_depth_b200 import ops
    N, H, W, C, ct, off = 2, 12, 20, 40, 64, 8                # 40 channels: the last k-block is half padding
    xb = rnd(1, (N, H, W, ct)); a = rnd(2, (ct,), 0.5, 1.5); b = rnd(3, (ct,), -0.5, 0.5)
    z = ops.z_alloc(N, C, H, W, DEV)
    assert z.shape == (2, N, 6, H * W, 8)
    ops.prep_operand(ops.make_src(ops.View(xb, off), a, b, True), C, z)
    torch.cuda.synchronize()
    ref = F.relu(xb[..., off:off + C] * a[off:off + C] + b[off:off + C]).permute(0, 3, 1, 2)
    got = z_to_nchw(z, 48, H, W)
    assert (got[:, :C] - ref).abs().max().item() <= 2 ** -16 * ref.abs().max().item()
    assert (got[:, C:] == 0).all()                              # channels beyond C are zero
    # view with a gap + BatchNorm/ReLU backward on load (the CVD_XF_BNBWD formula of include/cvd.h)
    dyb = rnd(4, (N, H, W, ct)); bw = rnd(5, (ct, 4), -1.0, 1.0)
    n0, gap = 16, 16
    src = ops.make_src(ops.View(xb, 0, n0, gap), a, b, True, dy=ops.View(dyb, 0, n0, gap), bw=bw)
    C2 = 32
    z2 = ops.z_alloc(N, C2, H, W, DEV)
    ops.prep_operand(src, C2, z2)
    torch.cuda.synchronize()
    phys = list(range(0, 16)) + list(range(32, 48))
    y = xb[..., phys] * a[phys] + b[phys]
    g = torch.where(y > 0, dyb[..., phys], torch.zeros_like(y))
    ref2 = (bw[phys, 0] * g - bw[phys, 1] - bw[phys, 2] * y).permute(0, 3, 1, 2)
    got2 = z_to_nchw(z2, C2, H, W)
    assert (got2 - ref2).abs().max().item() <= 2 ** -15 * ref2.abs().max().item()


def run_conv2(x_nchw, w, bias, k, flip=False, flags=0, y_init=None, out_pad=0, zc_extra=0, bn=None):
    """x: the already-transformed operand (identity prep); returns the NHWC output buffer."""
    from consistent_depth_b200 import ops
    N, cin, H, W = x_nchw.shape
    cout = w.shape[1] if flip else w.shape[0]
    xb = nhwc(x_nchw)
    z = ops.z_alloc(N, cin + zc_extra, H, W, DEV)
    z.zero_()
    zoff = zc_extra // 8
    ops.prep_operand(ops.make_src(ops.View(xb, 0)), cin, z, zoff)
    yb = torch.full((N, H, W, cout + out_pad), 7.0, device=DEV) if y_init is None else y_init
    pk = ops.conv2_pack(w, flip)
    ops.conv2(z, zoff, pk, bias, ops.make_dst(ops.View(yb, out_pad // 2)), N, H, W, cin, cout, k, flags, bn)
    torch.cuda.synchronize()
    return yb


SHAPES = [
    # cin, cout, k, N, H, W
    (64, 16, 11, 1, 32, 48),         # G = 6 x 2 groups, N = 96
    (64, 16, 11, 2, 40, 136),        # several column tiles (VW = 54) and row tiles, ragged right / bottom edges
    (32, 32, 3, 2, 32, 48),
    (32, 32, 7, 1, 30, 70),          # N = 224, one group
    (64, 64, 7, 1, 16, 24),          # WS = 32 window (W + k - 1 <= 32), two tap groups
    (64, 64, 11, 1, 24, 56),         # three tap groups (4 + 4 + 3), N = 256
    (32, 64, 5, 1, 20, 28),
    (64, 32, 11, 1, 14, 24),         # smallest hourglass level
    (128, 208, 1, 1, 32, 48),        # fused inception 1x1: weights resident in shared memory
    (256, 160, 1, 2, 16, 24),
    (256, 256, 1, 1, 28, 48),        # H*W = 1344 = 21 x 64: flattened rows of 64 slots; weights streamed
    (128, 128, 1, 1, 14, 24),        # H*W = 336 = 21 x 16
]


@pytest.mark.parametrize("cin,cout,k,N,H,W", SHAPES)
def test_conv2_forward_matches_torch(cin, cout, k, N, H, W):
    x = rnd(10 + cin + k, (N, cin, H, W))
    w = rnd(20 + cout + k, (cout, cin, k, k), -0.1, 0.1)
    bias = rnd(30 + cout, (cout,))
    y = run_conv2(x, w, bias, k)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=(k - 1) // 2)
    err = (y.permute(0, 3, 1, 2).double() - ref).abs().max().item()
    assert err <= 6e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("cin,cout,k,N,H,W", [(64, 16, 11, 1, 32, 48), (32, 64, 7, 2, 24, 40), (128, 208, 1, 1, 16, 24), (64, 32, 5, 1, 20, 72)])
def test_conv2_dgrad_matches_autograd(cin, cout, k, N, H, W):
    """Input gradient of conv(cin -> cout): flip-packed weights, GEMM cin = cout; accumulate flag; channel offset in Z."""
    w = rnd(5 + k, (cout, cin, k, k), -0.1, 0.1)
    g = rnd(6 + k, (N, cout, H, W))
    ref = torch.nn.grad.conv2d_input((N, cin, H, W), w.double(), g.double(), padding=(k - 1) // 2)
    y = run_conv2(g, w, None, k, flip=True, zc_extra=16)
    err = (y.permute(0, 3, 1, 2).double() - ref).abs().max().item()
    assert err <= 6e-5 * ref.abs().max().item(), (err, ref.abs().max().item())
    y0 = rnd(9, (N, H, W, cin + 8))
    y1 = run_conv2(g, w, None, k, flip=True, flags=1, y_init=y0.clone(), out_pad=8)
    got = (y1 - y0)[..., 4:4 + cin].permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max().item() <= 6e-5 * ref.abs().max().item() + 1e-6
    assert torch.equal(y1[..., :4], y0[..., :4]) and torch.equal(y1[..., 4 + cin:], y0[..., 4 + cin:])


def test_conv2_fused_bn_statistics_and_views():
    """Epilogue BatchNorm(train) statistics (cvd_bn_t) of the kx-fused kernel + destination view with a gap."""
    from consistent_depth_b200 import ops
    cin, cout, k, N, H, W = 32, 48, 5, 2, 40, 72
    x = rnd(1, (N, cin, H, W)); w = rnd(2, (cout, cin, k, k), -0.1, 0.1); bias = rnd(3, (cout,))
    ct = 96
    yb = torch.full((N, H, W, ct), 7.0, device=DEV)
    z = ops.z_alloc(N, cin, H, W, DEV)
    ops.prep_operand(ops.make_src(ops.View(nhwc(x), 0)), cin, z)
    a, b, rstd, mean = (torch.zeros(ct, device=DEV) for _ in range(4))
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    bn = ops.make_bn(ops.bn_scratch(DEV), a, b, rstd, mean, None, None, rm, rv)
    dstv = ops.View(yb, 16)
    ops.conv2(z, 0, ops.conv2_pack(w), bias, ops.make_dst(dstv), N, H, W, cin, cout, k, 0, bn)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=2)
    got = yb[..., 16:16 + cout].permute(0, 3, 1, 2).double()
    assert (got - ref).abs().max().item() <= 6e-5 * ref.abs().max().item()
    assert (yb[..., :16] == 7.0).all() and (yb[..., 16 + cout:] == 7.0).all()
    m = ref.mean(dim=(0, 2, 3)); v = ref.var(dim=(0, 2, 3), unbiased=False)
    np.testing.assert_allclose(mean[16:16 + cout].cpu().numpy(), m.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(rstd[16:16 + cout].cpu().numpy(), (1.0 / torch.sqrt(v + 1e-5)).cpu().numpy(), rtol=1e-4)
    np.testing.assert_allclose(a[16:16 + cout].cpu().numpy(), rstd[16:16 + cout].cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(rm.cpu().numpy(), 0.1 * m.cpu().numpy(), rtol=1e-4, atol=1e-6)
    n = N * H * W
    np.testing.assert_allclose(rv.cpu().numpy(), 0.9 + 0.1 * (v * n / (n - 1)).cpu().numpy(), rtol=1e-4)


def test_conv2_bench_layer_shapes_cropped():
    """The heaviest bench layers at full width (384 / 192 columns: 8 / 4 column tiles) on a few rows."""
    for cin, cout, k, H, W in ((64, 16, 11, 24, 384), (64, 32, 7, 20, 192), (128, 208, 1, 8, 384)):
        x = rnd(cin + k, (1, cin, H, W)); w = rnd(cout + k, (cout, cin, k, k), -0.05, 0.05)
        y = run_conv2(x, w, None, k)
        ref = F.conv2d(x.double(), w.double(), None, padding=(k - 1) // 2)
        err = (y.permute(0, 3, 1, 2).double() - ref).abs().max().item()
        assert err <= 6e-5 * ref.abs().max().item(), (cin, cout, k, err)


WG_SHAPES = [
    # cin, cout, k, N, H, W
    (64, 16, 11, 1, 20, 40),         # M = 2 ky rows x 64 ch, N = 96, three passes of 2 ky groups
    (64, 16, 11, 2, 33, 130),        # several column tiles (ragged), odd row count
    (32, 32, 7, 2, 24, 40),          # M = 4 ky rows x 32 ch, N = 64: one pass
    (64, 32, 7, 1, 16, 24),
    (32, 64, 5, 1, 20, 28),          # more output than input channels: 8 G chunks
    (64, 64, 11, 1, 14, 24),         # ky group split over two passes (8 chunks x 96 columns > 512)
    (32, 16, 3, 1, 12, 20),
    (128, 208, 1, 1, 32, 48),        # 1x1: K = flattened pixels, N = 208
    (256, 160, 1, 2, 16, 24),        # two M blocks
    (256, 256, 1, 1, 14, 24),        # H*W = 336: ragged last K tile
    (128, 112, 1, 1, 8, 12),
]


@pytest.mark.parametrize("cin,cout,k,N,H,W", WG_SHAPES)
def test_conv2_wgrad_matches_autograd(cin, cout, k, N, H, W):
    from consistent_depth_b200 import ops
    x = rnd(40 + cin + k, (N, cin, H, W))
    g = rnd(50 + cout + k, (N, cout, H, W))
    ref = torch.nn.grad.conv2d_weight(x.double(), (cout, cin, k, k), g.double(), padding=(k - 1) // 2)
    # operands inside wider planes at non-zero chunk offsets, as the engine uses them
    xz = ops.z_alloc(N, cin + 32, H, W, DEV); gz = ops.z_alloc(N, cout + 16, H, W, DEV)
    xz.zero_(); gz.zero_()
    ops.prep_operand(ops.make_src(ops.View(nhwc(x), 0)), cin, xz, 4)
    ops.prep_operand(ops.make_src(ops.View(nhwc(g), 0)), cout, gz, 2)
    dw = torch.zeros(cout, cin, k, k, device=DEV)
    assert ops.conv2_wgrad(xz, 4, gz, 2, dw, N, H, W, cin, cout, k)
    torch.cuda.synchronize()
    err = (dw.double() - ref).abs().max().item()
    assert err <= 6e-5 * ref.abs().max().item(), (err, ref.abs().max().item())
    # accumulation: a second call doubles the result
    assert ops.conv2_wgrad(xz, 4, gz, 2, dw, N, H, W, cin, cout, k)
    torch.cuda.synchronize()
    assert (dw.double() - 2 * ref).abs().max().item() <= 1.2e-4 * ref.abs().max().item()


def test_conv2_wgrad_reports_unsupported_shapes():
    from consistent_depth_b200 import ops
    N, H, W = 1, 8, 8
    xz = ops.z_alloc(N, 48, H, W, DEV); gz = ops.z_alloc(N, 16, H, W, DEV)
    dw = torch.zeros(16, 48, 3, 3, device=DEV)
    assert ops.conv2_wgrad(xz, 0, gz, 0, dw, N, H, W, 48, 16, 3) is False        # 48 input channels: 6 chunks do not tile M = 128
