"""GPU: MiDaS-v2 path — the passes of midas_ops.cu against the torch ops they replace, the grouped
(ResNeXt) convolution as block-diagonal chunks of the dense tcgen05 kernel, and the whole MidasEngine (forward,
backward, BN running statistics, fine-tune steps) against oracle/midas_oracle.py and the golden fixture produced
by the reference's own MidasNet.

Tolerances: element-wise passes 1e-6-class; convs 6e-5 of the output magnitude (bf16x3).
Network level, eval mode (BatchNorm = fixed affine map): well conditioned, depth rel 1e-4 -- the tight end-to-end check.
Network level, train mode on this fixture (96x160, batch statistics over as few as 30 samples in 2048 channels, random
weights) is ill-conditioned: in the fp64 oracle a relative perturbation of 1.8e-6 of every conv output (bf16x3-class;
fp32 is ~1e-7) already gives 1.5e-2 at layer4, 1.8e-3 on depth and gradient cosines of 0.965 (rel-L2 0.26); the
reference's own fp32 run differs from fp64 by 6e-4 / 9e-5 / 0.9987 (rel-L2 0.05).  The train-mode thresholds below
(layer 4e-2, depth 6e-3, gradient cosine >= 0.93 and norms within 10 %) are that measured envelope, not a kernel error.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import consistency_oracle as co
from oracle import midas_oracle as mo
from oracle import synth
from oracle.make_golden import MIDAS_CASE

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
def test_image_normalize():
    from consistent_depth_b200 import ops
    img = rnd(1, (2, 3, 10, 14), 0, 1)
    out = torch.full((2, 10, 14, 4), 9.0, device=DEV)
    ops.image_normalize(img, out, mo.NORM_MEAN, mo.NORM_STD)
    mean = torch.tensor(mo.NORM_MEAN, device=DEV).reshape(1, 3, 1, 1)
    std = torch.tensor(mo.NORM_STD, device=DEV).reshape(1, 3, 1, 1)
    close(out[..., :3], nhwc((img - mean) / std), rtol=1e-5, atol=1e-6)
    assert (out[..., 3] == 0).all()


def test_relu_add():
    from consistent_depth_b200 import ops
    x, o = rnd(2, (2, 5, 7, 16)), rnd(3, (2, 5, 7, 16))
    out = torch.zeros_like(x)
    ops.relu_add(x, None, out)
    close(out, F.relu(x))
    ops.relu_add(x, o, out)
    close(out, F.relu(x) + o)


@pytest.mark.parametrize("align", [True, False])
@pytest.mark.parametrize("h,w", [(3, 5), (1, 2), (6, 10), (2, 2)])
def test_bilinear_up2_forward_backward(align, h, w):
    from consistent_depth_b200 import ops
    N, C = 2, 8
    x = rnd(4, (N, h, w, C)).requires_grad_(True)
    r = rnd(5, (N, 2 * h, 2 * w, C))
    ref = F.interpolate(nchw(x), scale_factor=2, mode="bilinear", align_corners=align)
    out = torch.zeros(N, 2 * h, 2 * w, C, device=DEV)
    ops.up2_bilinear(x.detach(), None, False, out, align)
    close(out, nhwc(ref), rtol=1e-5, atol=1e-6)
    ops.up2_bilinear(x.detach(), r, True, out, align)
    close(out, nhwc(ref) + F.relu(r), rtol=1e-5, atol=1e-6)
    ops.up2_bilinear(x.detach(), r, False, out, align)
    close(out, nhwc(ref) + r, rtol=1e-5, atol=1e-6)
    g = rnd(6, (N, 2 * h, 2 * w, C))
    ref.backward(nchw(g))
    dx = rnd(7, (N, h, w, C)); d0 = dx.clone()
    ops.up2_bilinear_bwd(g, dx, align, True)
    close(dx - d0, x.grad, rtol=1e-5, atol=1e-5)
    ops.up2_bilinear_bwd(g, dx, align, False)
    close(dx, x.grad, rtol=1e-5, atol=1e-5)


def test_reciprocal_of_relu_forward_backward():
    from consistent_depth_b200 import ops
    raw = rnd(8, (2, 6, 8, 4), -0.5, 2.0)
    raw[0, 0, 0, 0] = 0.0                                     # clamped pixel: depth = inf, gradient exactly 0 (as torch)
    depth = torch.zeros(2, 6, 8, device=DEV)
    ops.recip_relu(raw, depth)
    r0 = raw[..., 0].clone().requires_grad_(True)
    ref = F.relu(r0).reciprocal()
    torch.cuda.synchronize()
    assert torch.equal(torch.isinf(depth), torch.isinf(ref))
    fin = torch.isfinite(ref)
    close(depth[fin], ref[fin])
    g = rnd(9, (2, 6, 8))
    ref.backward(g)
    draw = torch.full((2, 6, 8, 4), 3.0, device=DEV)
    ops.recip_relu_bwd(g, depth, raw, draw)
    close(draw[..., 0][fin], r0.grad[fin], rtol=1e-5, atol=1e-6)
    assert (draw[..., 0][~fin] == 0).all() and (draw[..., 1:] == 0).all()


# ------------------------------------------------------------------ grouped convolution as block-diagonal chunks
@pytest.mark.parametrize("width,H,W", [(256, 12, 20), (512, 6, 10), (1024, 6, 10), (2048, 3, 5)])
def test_grouped_conv_chunks_forward_dgrad_wgrad(width, H, W):
    """ResNeXt conv2: Conv2d(width, width, 3, padding=1, groups=32) through 64-channel chunks of the dense kernel."""
    from consistent_depth_b200 import ops
    N, gs, CH = 2, width // 32, 64
    x = rnd(10 + width, (N, width, H, W))
    w = rnd(11 + width, (width, gs, 3, 3), -0.2, 0.2)
    g = rnd(12 + width, (N, width, H, W))
    xb, gb = nhwc(x), nhwc(g)
    yb, dxb = torch.zeros(N, H, W, width, device=DEV), torch.zeros(N, H, W, width, device=DEV)
    dw = torch.zeros_like(w)
    for c0 in range(0, width, CH):
        pk = ops.pack_weights_grouped(w[c0:c0 + CH], CH, gs, False, 3)
        ops.conv(ops.make_src(ops.View(xb, c0)), pk, None, ops.make_dst(ops.View(yb, c0)), N, H, W, CH, CH, 3, 3, 0)
        pkt = ops.pack_weights_grouped(w[c0:c0 + CH], CH, gs, True, 3)
        ops.conv(ops.make_src(ops.View(gb, c0)), pkt, None, ops.make_dst(ops.View(dxb, c0)), N, H, W, CH, CH, 3, 3, 0)
        ops.conv_wgrad_grouped(ops.make_src(ops.View(gb, c0)), ops.make_src(ops.View(xb, c0)), dw[c0:c0 + CH], N, H, W, CH, gs, 3, 3)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    ref = F.conv2d(xr, wr, None, padding=1, groups=32)
    ref.backward(g.double())
    torch.cuda.synchronize()
    assert (nchw(yb).double() - ref).abs().max() <= 6e-5 * ref.abs().max()
    assert (nchw(dxb).double() - xr.grad).abs().max() <= 6e-5 * xr.grad.abs().max()
    assert (dw.double() - wr.grad).abs().max() <= 6e-5 * wr.grad.abs().max()


@pytest.mark.parametrize("width,H,W", [(256, 12, 20), (512, 6, 10), (2048, 3, 5)])
def test_grouped_conv_single_launch_equals_per_chunk_launches(width, H, W):
    """cvd_conv_fwd_chunks / cvd_conv_wgrad_grouped_chunks (blockIdx = chunk) against the per-chunk launches they
    replace: forward with BN+ReLU on load and fused output statistics, dgrad, wgrad -- forward / dgrad bit-identical."""
    from consistent_depth_b200 import ops
    N, gs, CH = 2, width // 32, 64
    nch = width // CH
    xb, gb = nhwc(rnd(20 + width, (N, width, H, W))), nhwc(rnd(21 + width, (N, width, H, W)))
    w = rnd(22 + width, (width, gs, 3, 3), -0.2, 0.2)
    a1, b1 = rnd(23, (width,), 0.5, 1.5), rnd(24, (width,), -0.3, 0.3)
    gamma, beta = rnd(25, (width,), 0.5, 1.5), rnd(26, (width,))
    nb = ops.packed_bytes(CH, CH, 3, 3)
    out = {}
    for mode in ("chunks", "loop"):
        yb, dxb = torch.zeros(N, H, W, width, device=DEV), torch.zeros(N, H, W, width, device=DEV)
        dw = torch.zeros_like(w)
        a, b, rstd, mean = (torch.zeros(width, device=DEV) for _ in range(4))
        rm, rv = torch.zeros(width, device=DEV), torch.ones(width, device=DEV)
        pk = torch.cat([ops.pack_weights_grouped(w[j * CH:(j + 1) * CH], CH, gs, False, 3) for j in range(nch)])
        pkt = torch.cat([ops.pack_weights_grouped(w[j * CH:(j + 1) * CH], CH, gs, True, 3) for j in range(nch)])
        if mode == "chunks":
            scratch = ops.bn_scratch(DEV, 256 * nch)
            bn = ops.make_bn(scratch, a, b, rstd, mean, gamma, beta, rm, rv)
            ops.conv_chunks(ops.make_src(ops.View(xb, 0), a1, b1, True), pk, None, ops.make_dst(ops.View(yb, 0)), N, H, W,
                            CH, CH, 3, nch, CH, CH, nb, 3, 0, bn=bn)
            ops.conv_chunks(ops.make_src(ops.View(gb, 0)), pkt, None, ops.make_dst(ops.View(dxb, 0)), N, H, W,
                            CH, CH, 3, nch, CH, CH, nb, 3, 0)
            ops.conv_wgrad_grouped_chunks(ops.make_src(ops.View(gb, 0)), ops.make_src(ops.View(xb, 0), a1, b1, True), dw,
                                          N, H, W, CH, nch, gs, 3, 3)
            torch.cuda.synchronize()
            assert (scratch == 0).all()
        else:
            scratch = ops.bn_scratch(DEV)
            for j in range(nch):
                c0 = j * CH
                bn = ops.make_bn(scratch, a, b, rstd, mean, gamma[c0:c0 + CH], beta[c0:c0 + CH], rm[c0:c0 + CH], rv[c0:c0 + CH])
                ops.conv(ops.make_src(ops.View(xb, c0), a1, b1, True), pk[j * nb:(j + 1) * nb], None, ops.make_dst(ops.View(yb, c0)),
                         N, H, W, CH, CH, 3, 3, 0, bn=bn)
                ops.conv(ops.make_src(ops.View(gb, c0)), pkt[j * nb:(j + 1) * nb], None, ops.make_dst(ops.View(dxb, c0)),
                         N, H, W, CH, CH, 3, 3, 0)
                ops.conv_wgrad_grouped(ops.make_src(ops.View(gb, c0)), ops.make_src(ops.View(xb, c0), a1, b1, True),
                                       dw[c0:c0 + CH], N, H, W, CH, gs, 3, 3)
            torch.cuda.synchronize()
        out[mode] = (yb, dxb, dw, a, b, rstd, mean, rm, rv)
    A, B = out["chunks"], out["loop"]
    assert torch.equal(A[0], B[0]) and torch.equal(A[1], B[1])
    assert (A[2] - B[2]).abs().max() <= 2e-6 * B[2].abs().max()          # RED order differs (different slab split)
    for i in range(3, 9):
        close(A[i], B[i], rtol=1e-5, atol=1e-6)
    # and against torch
    xin = torch.relu(nchw(xb).double() * a1.double()[None, :, None, None] + b1.double()[None, :, None, None])
    ref = F.conv2d(xin, w.double(), None, padding=1, groups=32)
    assert (nchw(A[0]).double() - ref).abs().max() <= 6e-5 * ref.abs().max()
    close(A[6].double(), ref.mean((0, 2, 3)), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------ the network
def _case():
    c = MIDAS_CASE
    return c, synth.make_pair_batch(c["seed"], c["pairs"], c["H"], c["W"]), mo.midas_init_state(c["seed"])


def _oracle_run(batch, sd, dtype=torch.float32):
    P, buffers = mo.to_torch(sd, dtype=dtype, requires_grad=True)
    cap = {}
    depth = mo.estimate_depth(torch.tensor(batch["images"], dtype=dtype), P, buffers, capture=cap)
    t = lambda a: torch.tensor(a, dtype=dtype)
    loss, _ = co.consistency_loss(depth, t(batch["extrinsics"]), t(batch["intrinsics"]),
                                  [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]], 1.0, 1e-4)
    depth.retain_grad()
    loss.backward()
    return P, buffers, cap, depth, loss


def test_midas_forward_matches_oracle_and_reference_golden():
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    c, batch, sd = _case()
    model = MidasV2Model(state_dict=sd)
    model.train()
    with torch.no_grad():
        depth = model(torch.tensor(batch["images"], device=DEV))
    P, buffers, cap, odepth, _ = _oracle_run(batch, sd)
    eng = model.engine(2, c["H"], c["W"])
    worst, seen = [], 0
    for key, raw in eng.raw_outputs.items():
        okey = key[:-len(".resConfUnit2")] if key.endswith(".resConfUnit2") else key
        if okey not in cap or key.endswith(".resConfUnit1"):
            continue
        want = cap[okey]
        got = nchw(raw)[:, :want.shape[1]].cpu()
        if got.shape != want.shape:
            continue
        seen += 1
        worst.append((float((got - want).abs().max() / want.abs().max()), key))
    worst.sort(reverse=True)
    assert seen > 130, seen
    assert worst[0][0] < 4e-2, worst[:6]
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "midas_small.npz"))
    np.testing.assert_allclose(depth.cpu().numpy(), odepth.detach().numpy(), rtol=6e-3)
    np.testing.assert_allclose(depth.cpu().numpy(), g["depth"], rtol=6e-3)
    st = model.state_dict()
    assert list(st.keys()) == list(mo.midas_param_shapes().keys())
    for k in g.files:
        if k.startswith("buf::"):
            np.testing.assert_allclose(st[k[5:]].cpu().numpy(), g[k], rtol=3e-2, atol=1e-4)
    assert int(st["pretrained.layer1.1.num_batches_tracked"]) == 1


def test_midas_eval_mode_forward_is_tight():
    """model.eval(): BatchNorm uses the running statistics (depth_fine_tuning.py:182 save_depth path) -- the composition
    of every forward kernel of the network, without the batch-statistics amplification."""
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    c, batch, sd = _case()
    sd = dict(sd)
    for i, k in enumerate(sd):                                   # non-trivial running statistics
        if k.endswith("running_mean"):
            sd[k] = synth.uniform(7, 3000 + i, sd[k].shape, -0.2, 0.2)
        elif k.endswith("running_var"):
            sd[k] = synth.uniform(7, 6000 + i, sd[k].shape, 0.5, 1.5)
    model = MidasV2Model(state_dict=sd)
    model.eval()
    with torch.no_grad():
        depth = model(torch.tensor(batch["images"], device=DEV))
    P, buffers = mo.to_torch(sd, dtype=torch.float64)
    with torch.no_grad():
        want = mo.estimate_depth(torch.tensor(batch["images"], dtype=torch.float64), P, buffers, train=False)
    np.testing.assert_allclose(depth.cpu().numpy(), want.numpy(), rtol=1e-4)


def test_midas_backward_matches_oracle():
    from consistent_depth_b200.monodepth import midas_arch
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    c, batch, sd = _case()
    model = MidasV2Model(state_dict=sd)
    model.train()
    P, buffers, cap, odepth, oloss = _oracle_run(batch, sd, torch.float64)
    depth = model(torch.tensor(batch["images"], device=DEV))
    model.P.grad_flat.zero_()
    depth.backward(odepth.grad.to(DEV, torch.float32))
    torch.cuda.synchronize()
    bad, checked = [], 0
    for k, _ in model.P.named_parameters():
        got = model.P._g(k).double().cpu()
        if midas_arch.dead_parameter(k):
            assert float(got.abs().max()) == 0.0, k
            continue
        want = P[k].grad
        nw = float(want.norm())
        if nw < 1e-9:
            continue
        checked += 1
        cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30))
        rel = float((got - want).norm() / nw)
        if not (cos > 0.93 and rel < 0.40 and abs(float(got.norm()) / nw - 1.0) < 0.10):
            bad.append((k, round(cos, 4), round(rel, 4), float(got.norm()), nw))
    assert checked > 300 and not bad, (len(bad), bad[:8])


def test_midas_fine_tune_steps_follow_oracle():
    """depth_fine_tuning.py:261-283 with model_type midas2 (lr 1e-4, lambda_view_baseline 1e-4): 2 steps."""
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    c, batch, sd = _case()
    model = MidasV2Model(state_dict=sd)
    model.train()
    step = FineTuneStep(model, 1, c["H"], c["W"], MidasV2Model.learning_rate)
    t = lambda a: torch.tensor(a)
    step.load_batch(t(batch["images"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]],
                    t(batch["extrinsics"]), t(batch["intrinsics"]))
    losses = [float(step.step()[0]) for _ in range(2)]
    P, buffers = mo.to_torch(sd, requires_grad=True)
    opt = torch.optim.Adam([P[k] for k in mo.trainable_keys()], MidasV2Model.learning_rate, betas=(0.9, 0.999))
    args = (t(batch["extrinsics"]), t(batch["intrinsics"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]])
    ol = []
    for _ in range(2):
        depth = mo.estimate_depth(t(batch["images"]), P, buffers)
        opt.zero_grad()
        loss, _ = co.consistency_loss(depth, *args, 1.0, 1e-4)
        loss.backward()
        opt.step()
        ol.append(float(loss.detach()[0]))
    assert losses[0] == pytest.approx(ol[0], rel=2e-3)
    np.testing.assert_allclose(losses, ol, rtol=5e-2)
    assert 500 < step.launches_per_step < 1200     # grouped / wide convs are single chunked launches (round 1: 2350)


def test_midas_registry_and_adapter_surface():
    from consistent_depth_b200.monodepth.depth_model_registry import get_depth_model
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    assert get_depth_model("midas2") is MidasV2Model
    m = MidasV2Model(pretrained=False)
    ps = list(m.parameters())
    assert len(ps) == 354 and sum(p.numel() for p in ps) == 105362945
    m.eval()
    with torch.no_grad():
        d = m(torch.rand(1, 2, 3, 64, 96, device=DEV))
    assert d.shape == (1, 2, 64, 96) and bool((d > 0).all())
