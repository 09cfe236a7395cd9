"""GPU: monodepth2 path — the passes of mono2_ops.cu against the torch ops they replace, the
>256-channel conv chunking, and the whole Mono2Engine (forward, backward, BN running statistics, fine-tune steps)
against oracle/monodepth2_oracle.py and the golden fixture the real reference modules produced.

Tolerances: element-wise passes 1e-6 relative (same fp32 arithmetic up to summation order / expf); convs as in
test_conv_gpu.py (bf16x3: 6e-5 of the output magnitude); network level, train mode: depth rel 2e-3, gradients by norm / cosine;
eval mode (well conditioned): depth rel 1e-4
(ReLU / max-pool selections make single elements ill-conditioned, see test_mc_gpu.py).
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import consistency_oracle as co
from oracle import monodepth2_oracle as m2
from oracle import synth
from oracle.make_golden import MONO2_CASE

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rnd(seed, shape, lo=-1.0, hi=1.0):
    return torch.tensor(synth.uniform(seed, 1, shape, lo, hi), device=DEV)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.cuda.synchronize()
    np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().cpu().numpy(), rtol=rtol, atol=atol)


# ------------------------------------------------------------------ element-wise passes
@pytest.mark.parametrize("H,W,oh,ow", [(24, 40, 64, 96), (48, 80, 32, 64), (13, 21, 32, 32)])
def test_bicubic_image_resize_and_normalise(H, W, oh, ow):
    from consistent_depth_b200 import ops
    img = rnd(1, (2, 3, H, W), 0.0, 1.0)
    out = torch.full((2, oh, ow, 4), 5.0, device=DEV)
    ops.bicubic_image(img, out)
    ref = (F.interpolate(img, size=(oh, ow), mode="bicubic", align_corners=False) - 0.45) / 0.225
    close(out[..., :3], nhwc(ref), rtol=1e-5, atol=2e-6)
    assert (out[..., 3] == 0).all()


@pytest.mark.parametrize("fh,fw,H,W", [(64, 96, 24, 40), (32, 64, 48, 80), (32, 32, 32, 32)])
def test_disparity_resize_reciprocal_forward_backward(fh, fw, H, W):
    from consistent_depth_b200 import ops
    disp = rnd(2, (2, fh, fw), 0.2, 0.9).requires_grad_(True)
    depth = torch.zeros(2, H, W, device=DEV)
    ops.disp_to_depth(disp.detach(), depth)
    ref = F.interpolate(disp[:, None], size=(H, W), mode="bicubic", align_corners=False)[:, 0].reciprocal()
    close(depth, ref, rtol=1e-5)
    g = rnd(3, (2, H, W))
    ref.backward(g)
    dd = torch.full((2, fh, fw), 3.0, device=DEV)            # must be zeroed by the call
    ops.disp_to_depth_bwd(g, depth, dd)
    close(dd, disp.grad, rtol=1e-4, atol=1e-5 * float(disp.grad.abs().max()))


def test_sigmoid_on_padded_grid():
    from consistent_depth_b200 import ops
    N, fh, fw = 2, 8, 12
    raw = rnd(4, (N, fh + 2, fw + 2, 4), -3, 3)
    disp = torch.zeros(N, fh, fw, device=DEV)
    ops.sigmoid_fwd(raw, disp)
    ref = torch.sigmoid(raw[:, 1:-1, 1:-1, 0])
    close(disp, ref)
    dd = rnd(5, (N, fh, fw))
    draw = torch.zeros_like(raw)
    ops.sigmoid_bwd(dd, disp, draw)
    close(draw[:, 1:-1, 1:-1, 0], dd * ref * (1 - ref))
    draw[:, 1:-1, 1:-1, 0] = 0
    assert (draw == 0).all()                                  # border and the padding channels stay zero


def test_stride2_pick_stuff_and_materialised_bn_backward():
    from consistent_depth_b200 import ops
    N, H, W, C = 2, 10, 14, 16
    x = rnd(6, (N, H, W, C))
    y = torch.zeros(N, H // 2, W // 2, C, device=DEV)
    ops.subsample2(x, y)
    close(y, x[:, ::2, ::2])
    d = rnd(7, (N, H, W, C)); d0 = d.clone()
    ops.stuff2(y, d, True)
    ref = d0.clone(); ref[:, ::2, ::2] += y
    close(d, ref)
    ops.stuff2(y, d, False)
    ref[:, ::2, ::2] = y
    close(d, ref)
    # BN backward materialised: c0 g - c1 - c2 yh, with the ReLU mask
    a, b = rnd(8, (C,), 0.5, 1.5), rnd(9, (C,))
    bw = rnd(10, (C, 4))
    dy = rnd(11, y.shape)
    for relu in (True, False):
        for stride in (1, 2):
            dst = torch.zeros(N, H, W, C, device=DEV) if stride == 2 else torch.zeros_like(y)
            ops.bnbwd_stuff(y, dy, a, b, bw, relu, dst, stride)
            yh = a * y + b
            g = dy * (yh > 0) if relu else dy
            want = bw[:, 0] * g - bw[:, 1] - bw[:, 2] * yh
            if stride == 2:
                close(dst[:, ::2, ::2], want)
                dst[:, ::2, ::2] = 0
                assert (dst == 0).all()
            else:
                close(dst, want)


@pytest.mark.parametrize("H,W", [(16, 24), (10, 14)])
def test_maxpool_3x3_s2_forward_backward(H, W):
    from consistent_depth_b200 import ops
    N, C = 2, 64
    x = rnd(12, (N, H, W, C)).requires_grad_(True)
    a, b = rnd(13, (C,), -1.5, 1.5), rnd(14, (C,))
    oh, ow = (H + 1) // 2, (W + 1) // 2
    out = torch.zeros(N, oh, ow, C, device=DEV)
    am = torch.zeros(N, oh, ow, C, dtype=torch.uint8, device=DEV)
    ops.maxpool_fwd(x.detach(), a, b, True, out, am)
    act = F.relu(nchw(x) * a[None, :, None, None] + b[None, :, None, None])
    act.retain_grad()
    ref = F.max_pool2d(act, 3, 2, 1)
    close(out, nhwc(ref))
    g = rnd(15, (N, oh, ow, C))
    ref.backward(nchw(g))
    dx = rnd(16, (N, H, W, C)); dx0 = dx.clone()
    ops.maxpool_bwd(g, am, dx, True)
    # gradient w.r.t. the POOLED ACTIVATION's input (post-ReLU tensor); ties among zeros die in the ReLU backward,
    # so compare after masking with the activation > 0
    mask = nhwc(act.detach() > 0)
    close((dx - dx0) * mask, nhwc(act.grad) * mask, rtol=1e-5, atol=1e-6)
    ops.maxpool_bwd(g, am, dx, False)
    close(dx * mask, nhwc(act.grad) * mask, rtol=1e-5, atol=1e-6)
    # every output gradient lands exactly once
    close(dx.sum((1, 2)), g.sum((1, 2)), rtol=1e-4, atol=1e-4)


def test_basic_block_tail_forward_backward():
    from consistent_depth_b200 import ops
    N, h, w, C = 2, 6, 10, 32
    y, r = rnd(17, (N, h, w, C)), rnd(18, (N, h, w, C))
    a, b, ra, rb = rnd(19, (C,), 0.5, 1.5), rnd(20, (C,)), rnd(21, (C,), 0.5, 1.5), rnd(22, (C,))
    out = torch.zeros_like(y)
    ops.bn_add_relu(y, a, b, r, None, None, out)
    close(out, F.relu(a * y + b + r))
    ops.bn_add_relu(y, a, b, r, ra, rb, out)
    ref = F.relu(a * y + b + ra * r + rb)
    close(out, ref)
    dout = rnd(23, out.shape); d0 = dout.clone()
    dres = rnd(24, out.shape); r0 = dres.clone()
    ops.relu_bwd_add(dout, out, dres, True)
    close(dout, d0 * (ref > 0)); close(dres, r0 + d0 * (ref > 0))
    dout = d0.clone()
    ops.relu_bwd_add(dout, out, dres, False)
    close(dres, d0 * (ref > 0))
    dout = d0.clone()
    ops.relu_bwd_add(dout, out, None, False)
    close(dout, d0 * (ref > 0))


@pytest.mark.parametrize("up,mode,s_pad,hs,ws", [(0, 0, 0, 2, 3), (0, 1, 1, 6, 10), (1, 1, 1, 5, 7), (0, 2, 0, 8, 6), (1, 0, 0, 1, 2)])
def test_reflect_pad_gather_forward_backward(up, mode, s_pad, hs, ws):
    """ReflectionPad2d(1)(T(src) [nearest x2]) written into a channel slice of a wider buffer, and its transpose."""
    from consistent_depth_b200 import ops
    N, C, Cd, doff = 2, 16, 40, 8
    src = rnd(25, (N, hs + 2 * s_pad, ws + 2 * s_pad, C + 8), -2, 2)
    a, b = rnd(26, (C + 8,), -1.5, 1.5), rnd(27, (C + 8,))
    inner = src[:, s_pad:s_pad + hs, s_pad:s_pad + ws, 4:4 + C].clone().requires_grad_(True)
    t = nchw(inner)
    if mode == 1:
        t = F.elu(t)
    elif mode == 2:
        t = F.relu(t * a[4:4 + C][None, :, None, None] + b[4:4 + C][None, :, None, None])
    if up:
        t = F.interpolate(t, scale_factor=2, mode="nearest")
    ref = F.pad(t, (1, 1, 1, 1), mode="reflect")
    hu, wu = hs << up, ws << up
    dst = torch.full((N, hu + 2, wu + 2, Cd), 7.0, device=DEV)
    ops.gather_pad_fwd(src, 4, s_pad, a if mode == 2 else None, b if mode == 2 else None, dst, doff, C, up, mode)
    close(dst[..., doff:doff + C], nhwc(ref))
    assert (dst[..., :doff] == 7).all() and (dst[..., doff + C:] == 7).all()
    # transpose
    dP = rnd(28, (N, hu + 2, wu + 2, Cd))
    if mode == 2:
        # the engine applies the ReLU / BN backward where the gradient is consumed: transpose of pad+upsample only
        tt = nchw(inner.detach()).clone().requires_grad_(True)
        r2 = F.pad(F.interpolate(tt, scale_factor=2, mode="nearest") if up else tt, (1, 1, 1, 1), mode="reflect")
        r2.backward(nchw(dP[..., doff:doff + C]))
        want = nhwc(tt.grad)
    else:
        ref.backward(nchw(dP[..., doff:doff + C]))
        want = inner.grad
    ds_pad = s_pad
    dsrc = rnd(29, (N, hs + 2 * ds_pad, ws + 2 * ds_pad, C + 8)); d0 = dsrc.clone()
    ops.gather_pad_bwd(dP, doff, src if mode == 1 else None, 4, s_pad, dsrc, 4, ds_pad, C, up, mode, True)
    got = (dsrc - d0)[:, ds_pad:ds_pad + hs, ds_pad:ds_pad + ws, 4:4 + C]
    close(got, want, rtol=1e-5, atol=1e-5)
    ops.gather_pad_bwd(dP, doff, src if mode == 1 else None, 4, s_pad, dsrc, 4, ds_pad, C, up, mode, False)
    close(dsrc[:, ds_pad:ds_pad + hs, ds_pad:ds_pad + ws, 4:4 + C], want, rtol=1e-5, atol=1e-5)
    # nothing outside the interior / channel slice was touched
    chk = dsrc.clone(); chk[:, ds_pad:ds_pad + hs, ds_pad:ds_pad + ws, 4:4 + C] = d0[:, ds_pad:ds_pad + hs, ds_pad:ds_pad + ws, 4:4 + C]
    assert torch.equal(chk, d0)


@pytest.mark.parametrize("C,ct,off", [(256, 256, 0), (16, 24, 8), (1, 4, 0), (64, 64, 0)])
def test_channel_sum(C, ct, off):
    from consistent_depth_b200 import ops
    x = rnd(30, (3, 17, 29, ct))
    out = rnd(31, (C,)); o0 = out.clone()
    ops.channel_sum(x, off, C, out)
    close(out, o0 + x[..., off:off + C].sum((0, 1, 2)), rtol=1e-4, atol=1e-4)


# ------------------------------------------------------------------ convs above 256 channels
@pytest.mark.parametrize("cin,cout,k,H,W", [(512, 512, 3, 6, 10), (256, 512, 1, 8, 12), (512, 256, 3, 12, 34)])
def test_conv_above_256_channels_forward_dgrad_wgrad(cin, cout, k, H, W):
    from consistent_depth_b200 import ops
    N = 2
    x = rnd(40 + cin, (N, cin, H, W))
    w = rnd(41 + cout, (cout, cin, k, k), -0.05, 0.05)
    bias = rnd(42, (cout,))
    xb, yb = nhwc(x), torch.zeros(N, H, W, cout, device=DEV)
    ops.conv(ops.make_src(ops.View(xb, 0)), ops.pack_weights(w, False, 3), bias, ops.make_dst(ops.View(yb, 0)), N, H, W, cin, cout, k, 3, 0)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=k // 2)
    torch.cuda.synchronize()
    assert (nchw(yb).double() - ref).abs().max() <= 6e-5 * ref.abs().max()
    # fused BN statistics through the chunked launch
    a, b, rstd, mean = (torch.zeros(cout, device=DEV) for _ in range(4))
    gamma, beta = rnd(43, (cout,), 0.5, 1.5), rnd(44, (cout,))
    rm, rv = torch.zeros(cout, device=DEV), torch.ones(cout, device=DEV)
    bn = ops.make_bn(ops.bn_scratch(DEV, cout), a, b, rstd, mean, gamma, beta, rm, rv)
    ops.conv(ops.make_src(ops.View(xb, 0)), ops.pack_weights(w, False, 3), bias, ops.make_dst(ops.View(yb, 0)), N, H, W, cin, cout, k, 3, 0, bn=bn)
    m = ref.mean((0, 2, 3)); v = ref.var((0, 2, 3), unbiased=False)
    close(mean.double(), m, rtol=1e-4, atol=1e-5)
    close(a.double(), gamma.double() / torch.sqrt(v + 1e-5), rtol=1e-4)
    # dgrad = conv with flipped / transposed weights, GEMM cin = cout
    g = rnd(45, (N, cout, H, W))
    dxb = torch.zeros(N, H, W, cin, device=DEV)
    ops.conv(ops.make_src(ops.View(nhwc(g), 0)), ops.pack_weights(w, True, 3), None, ops.make_dst(ops.View(dxb, 0)), N, H, W, cout, cin, k, 3, 0)
    dref = F.conv_transpose2d(g.double(), w.double(), padding=k // 2)
    torch.cuda.synchronize()
    assert (nchw(dxb).double() - dref).abs().max() <= 6e-5 * dref.abs().max()
    # wgrad
    dw = torch.zeros_like(w)
    ops.conv_wgrad(ops.make_src(ops.View(nhwc(g), 0)), ops.make_src(ops.View(xb, 0)), dw, N, H, W, cin, cout, k, 3)
    xr = x.double().requires_grad_(False)
    wr = w.double().clone().requires_grad_(True)
    F.conv2d(xr, wr, None, padding=k // 2).backward(g.double())
    torch.cuda.synchronize()
    assert (dw.double() - wr.grad).abs().max() <= 6e-5 * wr.grad.abs().max()


# ------------------------------------------------------------------ the network
def _case():
    c = MONO2_CASE
    batch = synth.make_pair_batch(c["seed"], c["pairs"], c["H"], c["W"])
    sd = m2.mono2_init_state(c["seed"])
    return c, batch, sd


def _oracle_run(batch, sd, feed, dtype=torch.float32):
    P, buffers = m2.to_torch(sd, dtype=dtype, requires_grad=True)
    cap = {}
    images = torch.tensor(batch["images"], dtype=dtype)
    depth = m2.estimate_depth(images, P, buffers, feed, capture=cap)
    t = lambda a: torch.tensor(a, dtype=dtype)
    loss, _ = co.consistency_loss(depth, t(batch["extrinsics"]), t(batch["intrinsics"]),
                                  [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]], 1.0, 1.0)
    depth.retain_grad()
    loss.backward()
    return P, buffers, cap, depth, loss


def test_mono2_forward_matches_oracle_and_reference_golden():
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    c, batch, sd = _case()
    model = Monodepth2Model(state_dict=sd, feed_size=c["feed"])
    model.train()
    images = torch.tensor(batch["images"], device=DEV)
    with torch.no_grad():
        depth = model(images)
    P, buffers, cap, odepth, _ = _oracle_run(batch, sd, c["feed"])
    eng = model.engine(2, c["H"], c["W"])
    # layer by layer: raw conv outputs
    worst = []
    for key, raw in eng.raw_outputs.items():
        want = cap[key]
        got = nchw(raw)
        if key.startswith("decoder."):
            got = got[:, :want.shape[1], 1:-1, 1:-1]          # interior of the padded grid (and the 1 real channel)
        elif got.shape != want.shape:
            got = got[:, :, ::2, ::2]                         # stride-2 conv computed at stride 1
        err = float((got.cpu() - want).abs().max() / want.abs().max())
        worst.append((err, key))
    worst.sort(reverse=True)
    assert worst[0][0] < 5e-3, worst[:5]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "monodepth2_small.npz"))
    np.testing.assert_allclose(depth.cpu().numpy(), odepth.detach().numpy(), rtol=2e-3)
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth"], rtol=2e-3)
    # BN running statistics (momentum 0.1, unbiased variance) of a few layers
    st = model.state_dict()
    for k in g.files:
        if k.startswith("buf::"):
            np.testing.assert_allclose(st[k[5:]].cpu().numpy(), g[k], rtol=3e-3, atol=1e-5)
    assert int(st["encoder.bn1.num_batches_tracked"]) == 1 and (st["height"], st["width"]) == tuple(c["feed"])


def test_mono2_eval_mode_forward_is_tight():
    """model.eval() (running statistics): the composition of every forward kernel without batch-statistics amplification."""
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    c, batch, sd = _case()
    sd = dict(sd)
    for i, k in enumerate(sd):
        if k.endswith("running_mean"):
            sd[k] = synth.uniform(7, 3000 + i, sd[k].shape, -0.2, 0.2)
        elif k.endswith("running_var"):
            sd[k] = synth.uniform(7, 6000 + i, sd[k].shape, 0.5, 1.5)
    model = Monodepth2Model(state_dict=sd, feed_size=c["feed"])
    model.eval()
    with torch.no_grad():
        depth = model(torch.tensor(batch["images"], device=DEV))
    P, buffers = m2.to_torch(sd, dtype=torch.float64)
    with torch.no_grad():
        want = m2.estimate_depth(torch.tensor(batch["images"], dtype=torch.float64), P, buffers, c["feed"], train=False)
    np.testing.assert_allclose(depth.cpu().numpy(), want.numpy(), rtol=1e-4)


def test_mono2_backward_matches_oracle():
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    from consistent_depth_b200.monodepth import mono2_arch
    c, batch, sd = _case()
    model = Monodepth2Model(state_dict=sd, feed_size=c["feed"])
    model.train()
    P, buffers, cap, odepth, oloss = _oracle_run(batch, sd, c["feed"], torch.float64)
    images = torch.tensor(batch["images"], device=DEV)
    depth = model(images)
    model.P.grad_flat.zero_()
    depth.backward(odepth.grad.to(DEV, torch.float32))        # same upstream gradient as the oracle
    torch.cuda.synchronize()
    bad = []
    for k, _ in model.P.named_parameters():
        got = model.P._g(k).double().cpu()
        if mono2_arch.dead_parameter(k):
            assert float(got.abs().max()) == 0.0, k
            continue
        want = P[k].grad
        nw = float(want.norm())
        if nw < 1e-7:
            continue
        cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
        rel = float((got - want).norm() / nw)
        if not (cos > 0.99 and rel < 0.15):
            bad.append((k, cos, rel, float(got.norm()), nw))
    assert not bad, bad[:8]


def test_mono2_fine_tune_steps_follow_oracle():
    """depth_fine_tuning.py:261-283 with model_type monodepth2 (lr 4e-5, lambda_view_baseline 1): 3 steps."""
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    c, batch, sd = _case()
    model = Monodepth2Model(state_dict=sd, feed_size=c["feed"])
    model.train()
    step = FineTuneStep(model, 1, c["H"], c["W"], Monodepth2Model.learning_rate)
    t = lambda a: torch.tensor(a)
    step.load_batch(t(batch["images"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]],
                    t(batch["extrinsics"]), t(batch["intrinsics"]))
    losses = []
    for _ in range(3):
        losses.append(float(step.step()[0]))
    # oracle loop (fp32)
    P, buffers = m2.to_torch(sd, requires_grad=True)
    keys = m2.trainable_keys()
    opt = torch.optim.Adam([P[k] for k in keys], Monodepth2Model.learning_rate, betas=(0.9, 0.999))
    args = (t(batch["extrinsics"]), t(batch["intrinsics"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]])
    ol = []
    for _ in range(3):
        depth = m2.estimate_depth(t(batch["images"]), P, buffers, c["feed"])
        opt.zero_grad()
        loss, _ = co.consistency_loss(depth, *args, 1.0, 1.0)
        loss.backward()
        opt.step()
        ol.append(float(loss[0]))
    np.testing.assert_allclose(losses, ol, rtol=3e-2)
    assert losses[0] == pytest.approx(ol[0], rel=1e-3)
    w = model.state_dict()["decoder.10.conv.bias"].cpu().numpy()
    np.testing.assert_allclose(w, P["decoder.10.conv.bias"].detach().numpy(), rtol=5e-2, atol=1e-4)
    assert step.launches_per_step > 100


def test_mono2_registry_and_adapter_surface():
    from consistent_depth_b200.monodepth.depth_model_registry import get_depth_model
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    assert get_depth_model("monodepth2") is Monodepth2Model
    assert (Monodepth2Model.align, Monodepth2Model.learning_rate, Monodepth2Model.lambda_view_baseline) == (1, 0.00004, 1)
    m = Monodepth2Model(feed_size=(64, 96))
    ps = list(m.parameters())
    assert len(ps) == 90 and sum(p.numel() for p in ps) == 14842236
    m.eval()
    with torch.no_grad():
        d = m(torch.rand(1, 2, 3, 24, 40, device=DEV))
    assert d.shape == (1, 2, 24, 40) and bool(torch.isfinite(d).all()) and bool((d > 0).all())
    assert m.save("unused") is None                                      # monodepth2_model.py:92-93
