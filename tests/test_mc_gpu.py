"""GPU: the sm_100a hourglass engine (forward + hand-written backward + fused Adam) vs the oracle and the
reference-generated goldens (tests/golden/hourglass_small.npz, finetune_steps.npz).

Stated tolerances (north_star: per-pixel depth L1 <= 1e-3 vs reference; bf16x3-split tensor-core convs):
  forward: every conv's raw output max-abs err <= 2e-3 of its max magnitude; depth mean relative L1 <= 1e-3
  backward: gradient norms within 5 %, selected gradient tensors cosine >= 0.99 and relative L2 <= 15 % (the fp32 and
            fp64 oracles themselves differ by up to 1e-3 in these norms: error amplification ~2000x at random init)
  3 fine-tune steps: first loss rel 1e-4; trajectory within the reference's own fp32-vs-fp64 divergence envelope
"""
import os
import types

import numpy as np
import pytest
import torch

from oracle import synth, hourglass_oracle as ho, consistency_oracle as co

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_model(seed):
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    sd = {k: torch.tensor(np.asarray(v)) for k, v in ho.mc_init_state(seed).items()}
    return MannequinChallengeModel(state_dict=sd)


def metadata(batch):
    t = lambda a: torch.tensor(a, device=DEV)
    return {"extrinsics": t(batch["extrinsics"]), "intrinsics": t(batch["intrinsics"]),
            "geometry_consistency": {"indices": t(batch["indices"]), "flows": [t(f) for f in batch["flows"]],
                                     "masks": [t(m) for m in batch["masks"]]}}


def test_forward_layerwise_and_backward_match_reference(golden_dir):
    from consistent_depth_b200.loss.joint_loss import JointLoss
    g = np.load(os.path.join(golden_dir, "hourglass_small.npz"))
    seed, H, W = 21, 32, 48
    model = make_model(seed).train()
    batch = synth.make_pair_batch(seed, [(0, 1)], H, W)
    images = torch.tensor(batch["images"], device=DEV)
    depth = model(images, None)
    torch.cuda.synchronize()
    # layer-by-layer against the oracle (CPU fp32)
    P, buffers = ho.to_torch(ho.mc_init_state(seed))
    cap = {}
    with torch.no_grad():
        ho.estimate_depth(torch.tensor(batch["images"]), P, buffers, cap)
    eng = model.engine(2, H, W)
    worst = (0.0, None)
    for key, ref in cap.items():
        if key not in eng.raw_outputs:          # fused 1x1: members live at channel offsets of the first one
            continue
    for key, (buf, off, cout) in eng.raw_outputs.items():
        if key == "pred_layer":
            continue
        if key.endswith("convs.0.0"):           # fused 1x1 GEMM: o0 | a1 | a2 | a3
            pre = key[:-len("convs.0.0")]
            ref = torch.cat([cap[f"{pre}convs.{i}.0"] for i in range(4)], 1)
        else:
            ref = cap[key]
        got = buf[..., off:off + cout].permute(0, 3, 1, 2).cpu()
        err = (got - ref).abs().max().item() / ref.abs().max().item()
        if err > worst[0]:
            worst = (err, key)
        assert err <= 2e-3, f"{key}: raw conv output rel err {err:.3e}"
    d = depth.detach().cpu().numpy()
    rel = np.abs(d - g["depth"]) / g["depth"]
    print(f"worst layer err {worst}; depth mean rel L1 {rel.mean():.3e} max {rel.max():.3e}")
    assert rel.mean() <= 1e-3
    # BN running statistics (train-mode side effect)
    sd = model.state_dict()
    for k in g.files:
        if k.startswith("buf::"):
            np.testing.assert_allclose(sd[k[5:]].cpu().numpy(), g[k], rtol=2e-3, atol=2e-4)
    # loss + backward through the reference-style API
    opt = types.SimpleNamespace(lambda_view_baseline=0.1, lambda_reprojection=1.0, lambda_parameter=0)
    params = model.parameters()
    model.P.grad_flat.zero_()
    loss, _ = JointLoss(opt)(depth, metadata(batch))
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss"], rtol=2e-3)
    loss.backward()
    torch.cuda.synchronize()
    named = dict(model.P.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    norms = np.array([float(model.P._g(k).double().norm()) for k in names])
    big = g["grad_norms"] > 1e-4
    np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=5e-2)   # fp32-vs-fp64 oracle already differs by up to 1e-3 here: ~2000x error amplification at random init, BN over 2 frames
    # Gradient tensors: direction and size.  Element-wise max-norm bounds are meaningless here: ONE ReLU whose
    # pre-activation is within 1e-5 of zero flips between the bf16x3 engine and the fp32 reference and changes
    # that channel's weight gradient by a full single-pixel contribution (~5 % of its max; observed and traced
    # with tools/debug_mc_grads.py), so the bar is cosine similarity >= 0.99 and relative L2 error <= 15 %.
    for k in g.files:
        if k.startswith("grad::"):
            ref = g[k].astype(np.float64)
            got = model.P._g(k[6:]).cpu().numpy().astype(np.float64)
            cos = float((got * ref).sum() / np.sqrt((got * got).sum() * (ref * ref).sum()))
            rel = float(np.sqrt(((got - ref) ** 2).sum() / (ref * ref).sum()))
            print(f"{k[6:]}: rel-L2 {rel:.2e} cos {cos:.5f}")
            assert cos >= 0.99 and rel <= 0.15, (k, rel, cos)


def test_three_finetune_steps_match_reference(golden_dir):
    from consistent_depth_b200.loss.joint_loss import JointLoss
    from consistent_depth_b200 import optimizer
    g = np.load(os.path.join(golden_dir, "finetune_steps.npz"))
    seed, H, W = 41, 32, 48
    model = make_model(seed).train()
    batch = synth.make_pair_batch(seed, [(0, 2)], H, W)
    images = torch.tensor(batch["images"], device=DEV)
    meta = metadata(batch)
    crit = JointLoss(types.SimpleNamespace(lambda_view_baseline=0.1, lambda_reprojection=1.0, lambda_parameter=0))
    opt = optimizer.create("Adam", model.parameters(), 4e-4, betas=(0.9, 0.999))
    losses = []
    for _ in range(3):          # depth_fine_tuning.py:264-283
        depth = model(images, meta)
        opt.zero_grad()
        loss, _m = crit(depth, meta, parameters=model.parameters())
        loss.backward()
        opt.step()
        losses.append(float(loss[0]))
    with torch.no_grad():
        depth = model(images, None)
    torch.cuda.synchronize()
    print("losses", losses, "ref", g["losses"])
    # Step 0 sees identical weights: tight.  Later steps: Adam's first updates are ~lr*sign(g), so every
    # sign flip of a noise-level gradient moves a weight by 2*lr and the trajectories separate chaotically —
    # the reference's OWN fp32 and fp64 runs of this case give losses [2.22674, 1.93275, 1.47912] vs
    # [2.22674, 1.93348, 1.48286] and final depths 13 % apart (oracle, measured on CPU).  The engine must
    # stay inside that same envelope: loss within 3 % per step, final depth within 30 %.
    np.testing.assert_allclose(losses[0], g["losses"][0], rtol=1e-4)
    np.testing.assert_allclose(losses, g["losses"], rtol=3e-2)
    rel = np.abs(depth.cpu().numpy() - g["final_depth"]) / g["final_depth"]
    print(f"final depth mean rel L1 {rel.mean():.3e}")
    assert rel.mean() <= 0.3


def test_network_parity_at_bench_resolution():
    """2 frames at the BENCH resolution 224x384 (tile shapes MT=2/4/8, 148-CTA fused-BN epilogues, 5 hourglass levels)
    against the CPU fp32 oracle: every conv's raw output <= 2e-3 of its max, depth mean relative L1 <= 1e-3
    (north-star bar), loss rel 2e-3, gradient norms within 5 % and selected tensors cosine >= 0.99."""
    from consistent_depth_b200.utils.geometry import fused_consistency
    seed, H, W = 23, 224, 384
    model = make_model(seed).train()
    batch = synth.make_pair_batch(seed, [(0, 1)], H, W)
    images = torch.tensor(batch["images"], device=DEV)
    eng = model.engine(2, H, W)
    depth = eng.forward(images.view(2, 3, H, W)).view(1, 2, H, W)
    torch.cuda.synchronize()
    P, buffers = ho.to_torch(ho.mc_init_state(seed), requires_grad=True)
    cap = {}
    d_ref = ho.estimate_depth(torch.tensor(batch["images"]), P, buffers, cap)
    worst = (0.0, None)
    for key, (buf, off, cout) in eng.raw_outputs.items():
        if key == "pred_layer":
            continue
        if key.endswith("convs.0.0"):
            pre = key[:-len("convs.0.0")]
            ref = torch.cat([cap[f"{pre}convs.{i}.0"] for i in range(4)], 1)
        else:
            ref = cap[key]
        got = buf[..., off:off + cout].permute(0, 3, 1, 2).cpu()
        err = (got - ref.detach()).abs().max().item() / ref.abs().max().item()
        worst = max(worst, (err, key))
        assert err <= 2e-3, f"{key}: raw conv output rel err {err:.3e}"
    rel = (depth.cpu() - d_ref.detach()).abs() / d_ref.detach()
    print(f"224x384: worst layer err {worst}; depth mean rel L1 {rel.mean():.3e} max {rel.max():.3e}")
    assert rel.mean().item() <= 1e-3
    # loss + backward
    t = lambda a: torch.tensor(a, device=DEV)
    loss, pair, gd = fused_consistency(depth, [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]],
                                       t(batch["extrinsics"]), t(batch["intrinsics"]), 1.0, 0.1)
    model.P.grad_flat.zero_()
    eng.backward(gd.view(2, H, W))
    torch.cuda.synchronize()
    tc = lambda a: torch.tensor(a)
    lref, _ = co.consistency_loss(d_ref, tc(batch["extrinsics"]), tc(batch["intrinsics"]), [tc(f) for f in batch["flows"]],
                                  [tc(m) for m in batch["masks"]], 1.0, 0.1)
    np.testing.assert_allclose(float(loss), float(lref), rtol=2e-3)
    lref.backward()
    worst_n = 0.0
    for k in ho.trainable_keys():
        # a conv bias in front of a BatchNorm has an exactly-zero true gradient (the batch mean removes it): what either
        # implementation reports there is rounding noise, so only weights and the un-normalised head are compared
        if k.endswith(".bias") and k != "pred_layer.bias":
            continue
        ref = P[k].grad.double()
        got = model.P._g(k).cpu().double().reshape(ref.shape)
        rn, gn = float(ref.norm()), float(got.norm())
        if rn > 1e-4:
            worst_n = max(worst_n, abs(gn - rn) / rn)
            assert abs(gn - rn) <= 5e-2 * rn, (k, gn, rn)
    for k in ("seq.0.weight", "pred_layer.weight", "seq.3.list.1.0.convs.3.3.weight", "seq.3.list.0.1.convs.0.0.weight"):
        ref = P[k].grad.double().flatten()
        got = model.P._g(k).cpu().double().flatten()
        cos = float((got * ref).sum() / (got.norm() * ref.norm()))
        print(f"224x384 grad {k}: cos {cos:.5f}")
        assert cos >= 0.99, (k, cos)
    print(f"224x384: worst gradient-norm rel err {worst_n:.3e}")


def test_graph_replay_equals_serial_at_bench_size():
    """The captured CUDA graph (multi-stream branches) at the bench configuration 8 x 224 x 384 computes what the
    un-graphed, single-stream serial plan computes: depth rel <= 1e-5, flat gradient relative L2 <= 1e-4 (the only
    run-to-run differences are the order of fp32 REDs in wgrad / fp64 atomics in the fused BN statistics)."""
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    seed, H, W, B = 29, 224, 384, 4
    pairs = [(0, 1), (1, 2), (2, 4), (0, 3)]
    batch = synth.make_pair_batch(seed, pairs, H, W)
    t = lambda a: torch.tensor(a, device=DEV)
    res = {}
    for name, use_graph in (("serial", False), ("graph", True)):
        model = make_model(seed).train()
        step = FineTuneStep(model, B, H, W, lr=4e-4, use_graph=use_graph)
        if not use_graph:
            step.engine.multi_stream = False
        step.load_batch(t(batch["images"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]],
                        t(batch["extrinsics"]), t(batch["intrinsics"]))
        loss = step.step()
        torch.cuda.synchronize()
        res[name] = (float(loss), step.engine.depth.clone().cpu(), model.P.grad_flat.clone().cpu(), model.P.buf_flat.clone().cpu())
        del step, model
        torch.cuda.empty_cache()
    (l0, d0, g0, b0), (l1, d1, g1, b1) = res["serial"], res["graph"]
    assert abs(l0 - l1) <= 1e-5 * abs(l0), (l0, l1)
    assert ((d0 - d1).abs() / d0).max().item() <= 1e-5
    rel = float((g0.double() - g1.double()).norm() / g0.double().norm())
    print(f"graph vs serial at 8x224x384: loss {l0} {l1}; grad rel-L2 {rel:.2e}")
    assert rel <= 1e-4
    np.testing.assert_allclose(b1.numpy(), b0.numpy(), rtol=1e-5, atol=1e-6)      # BN running statistics
