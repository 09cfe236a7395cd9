#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    python -m oracle.make_golden

Inputs are regenerated deterministically by oracle/synth.py (not stored); the
fixtures hold only what the real reference modules produced for them:
  consistency_*.npz : loss/consistency_loss.py ConsistencyLoss via loss/joint_loss.py JointLoss
                      (+ autograd dL/d depth), fp32 and fp64
  hourglass_small.npz : monodepth/mannequin_challenge/models/hourglass.py HourglassModel(3)
                      forward (train mode) + grads of a consistency loss, BN running stats
  adam.npz          : optimizer.create("Adam", ...) == torch.optim.Adam, 6 steps
  monodepth2_small.npz: monodepth/monodepth2/networks ResnetEncoder(18) + DepthDecoder driven exactly as
                      monodepth/monodepth2_model.py:63-89 does (bicubic in, disp0, bicubic out, reciprocal),
                      train mode, + grads of the consistency loss (lambda_view_baseline = 1), BN running stats
  midas_small.npz   : monodepth/midas_v2/midas_net.py MidasNet (trunk = torchvision resnext101_32x8d in place of the
                      unreachable torch.hub WSL model, same layer graph) driven as monodepth/midas_v2_model.py:52-69
                      does, train mode, + grads of the consistency loss (lambda_view_baseline = 1e-4), BN running stats
  flowmask.npz      : utils/consistency.py consistent_flow_masks (+ its sample()) on two synthetic frame pairs
  parameter_loss.npz: loss/parameter_loss.py ParameterLoss through loss/joint_loss.py JointLoss (lambda_parameter > 0,
                      consistency terms off and on) + autograd gradients
  finetune_steps.npz: depth_fine_tuning.py:261-283 inner loop (model -> zero_grad -> JointLoss ->
                      backward -> step), 3 steps on one pair
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_import, synth, hourglass_oracle as ho, monodepth2_oracle as m2, midas_oracle as mo, flowmask_oracle as fo  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CONSISTENCY_CASES = {
    # name: (seed, pairs, H, W, stress, lambda_r, lambda_b)
    "geo_b2": (11, [(0, 1), (2, 5)], 24, 32, False, 1.0, 0.1),
    "stress_b3_oddw": (12, [(0, 2), (1, 3), (4, 6)], 20, 30, True, 1.0, 0.1),
    "reproj_only": (13, [(0, 1)], 16, 24, False, 1.0, 0.0),
    "disp_only": (14, [(0, 3), (1, 2)], 16, 24, False, 0.0, 1.0),
    "midas_lambda": (15, [(3, 4)], 24, 32, False, 1.0, 1e-4),
}


def ref_joint_loss(lambda_r, lambda_b, dtype):
    import loaders.video_dataset as vd
    import loss.joint_loss as jl
    import utils.torch_helpers as th
    vd._dtype = dtype
    jl._dtype = dtype
    assert str(th._device) == "cpu", "golden vectors are generated on the CPU reference path"
    opt = types.SimpleNamespace(lambda_view_baseline=lambda_b, lambda_reprojection=lambda_r, lambda_parameter=0)
    return jl.JointLoss(opt)


def to_metadata(batch, dtype):
    t = lambda a: torch.tensor(a, dtype=dtype)
    return {
        "extrinsics": t(batch["extrinsics"]), "intrinsics": t(batch["intrinsics"]),
        "geometry_consistency": {
            "indices": torch.tensor(batch["indices"]),
            "flows": [t(f) for f in batch["flows"]], "masks": [t(m) for m in batch["masks"]],
        },
    }


def gen_consistency():
    for name, (seed, pairs, H, W, stress, lr_, lb_) in CONSISTENCY_CASES.items():
        batch = synth.make_pair_batch(seed, pairs, H, W, stress=stress)
        depth = synth.synth_depth_pred(seed, len(pairs), H, W)
        out = {}
        for tag, dtype in (("f32", torch.float32), (("f64"), torch.float64)):
            crit = ref_joint_loss(lr_, lb_, dtype)
            d = torch.tensor(depth, dtype=dtype, requires_grad=True)
            loss, meta = crit(d, to_metadata(batch, dtype))
            loss.backward()
            out[f"loss_{tag}"] = loss.detach().numpy()
            out[f"reprojection_{tag}"] = meta["reprojection"].detach().numpy()
            out[f"disparity_{tag}"] = meta["disparity"].detach().numpy()
            out[f"grad_{tag}"] = d.grad.numpy()
        np.savez_compressed(os.path.join(OUT, f"consistency_{name}.npz"), **out)
        print(name, out["loss_f32"], out["loss_f64"])


def gen_hourglass():
    from monodepth.mannequin_challenge.models.hourglass import HourglassModel
    seed, H, W = 21, 32, 48
    sd = ho.mc_init_state(seed)
    net = HourglassModel(3)
    net.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    net.train()
    batch = synth.make_pair_batch(seed, [(0, 1)], H, W)
    images = torch.tensor(batch["images"])                                   # (1,2,3,H,W)
    pred, _ = net.forward(images.reshape(-1, 3, H, W))                       # adapter: mannequin_challenge_model.py:52-69
    depth = torch.exp(pred.reshape(1, 2, H, W))
    crit = ref_joint_loss(1.0, 0.1, torch.float32)
    loss, meta = crit(depth, to_metadata(batch, torch.float32))
    loss.backward()
    out = {"log_depth": pred.detach().numpy(), "depth": depth.detach().numpy(), "loss": loss.detach().numpy()}
    names, norms = [], []
    for k, p in net.named_parameters():
        if p.grad is None:
            continue
        names.append(k); norms.append(float(p.grad.double().norm()))
    out["grad_names"] = np.array(names); out["grad_norms"] = np.array(norms)
    for k in ("seq.0.weight", "seq.1.weight", "seq.1.bias", "pred_layer.weight", "pred_layer.bias",
              "seq.3.list.1.0.convs.3.3.weight", "seq.3.list.0.1.convs.0.0.weight",
              "seq.3.list.0.3.list.0.3.list.1.3.list.0.0.convs.2.3.weight"):
        out["grad::" + k] = dict(net.named_parameters())[k].grad.numpy()
    st = net.state_dict()
    for k in ("seq.1.running_mean", "seq.1.running_var", "seq.3.list.1.0.convs.3.4.running_var",
              "seq.3.list.0.5.convs.0.1.running_mean"):
        out["buf::" + k] = st[k].numpy()
    np.savez_compressed(os.path.join(OUT, "hourglass_small.npz"), **out)
    print("hourglass loss", out["loss"], "ngrads", len(names))


MONO2_CASE = dict(seed=51, H=24, W=40, feed=(64, 96), pairs=[(0, 1)])
MONO2_FULL_GRADS = ("encoder.conv1.weight", "encoder.bn1.weight", "encoder.bn1.bias", "encoder.layer1.0.conv1.weight",
                    "encoder.layer2.0.downsample.0.weight", "encoder.layer2.0.downsample.1.weight",
                    "encoder.layer4.1.bn2.weight", "encoder.layer4.1.bn2.bias", "decoder.0.conv.conv.bias",
                    "decoder.7.conv.conv.weight", "decoder.9.conv.conv.weight", "decoder.10.conv.weight",
                    "decoder.10.conv.bias")
MONO2_BUFS = ("encoder.bn1.running_mean", "encoder.bn1.running_var", "encoder.layer2.0.downsample.1.running_var",
              "encoder.layer4.1.bn2.running_mean", "encoder.layer3.0.bn1.running_var")


def gen_monodepth2():
    import warnings
    warnings.filterwarnings("ignore")
    from monodepth.monodepth2.networks.resnet_encoder import ResnetEncoder
    from monodepth.monodepth2.networks.depth_decoder import DepthDecoder
    c = MONO2_CASE
    seed, H, W = c["seed"], c["H"], c["W"]
    sd = m2.mono2_init_state(seed)
    enc = ResnetEncoder(18, False)
    dec = DepthDecoder(num_ch_enc=enc.num_ch_enc, scales=range(4))
    enc.load_state_dict({k: torch.tensor(v) for k, v in sd.items() if k.startswith("encoder.")})
    dec.load_state_dict({k: torch.tensor(v) for k, v in sd.items() if k.startswith("decoder.")})
    enc.train(); dec.train()
    batch = synth.make_pair_batch(seed, c["pairs"], H, W)
    images = torch.tensor(batch["images"])                                   # (1,2,3,H,W)
    # the adapter body, monodepth2_model.py:63-89 (the class itself hard-codes CUDA and a download)
    x = images.reshape(-1, 3, H, W)
    x = torch.nn.functional.interpolate(x, size=list(c["feed"]), mode="bicubic", align_corners=False)
    feats = enc(x)
    disp0 = dec(feats)[("disp", 0)]
    disp = torch.nn.functional.interpolate(disp0, size=[H, W], mode="bicubic", align_corners=False)
    depth = disp.reciprocal().reshape(1, 2, H, W)
    crit = ref_joint_loss(1.0, 1.0, torch.float32)
    loss, meta = crit(depth, to_metadata(batch, torch.float32))
    loss.backward()
    out = {"resized": x.detach().numpy(), "disp0": disp0.detach().numpy(), "depth": depth.detach().numpy(),
           "loss": loss.detach().numpy(), "feat4": feats[4].detach().numpy(),
           "feat0_mean": feats[0].detach().mean((0, 2, 3)).numpy()}
    params = dict(list(enc.named_parameters()) + list(dec.named_parameters()))
    names, norms = [], []
    for k, p in params.items():
        if p.grad is None:
            continue
        names.append(k); norms.append(float(p.grad.double().norm()))
    out["grad_names"] = np.array(names); out["grad_norms"] = np.array(norms)
    for k in MONO2_FULL_GRADS:
        out["grad::" + k] = params[k].grad.numpy()
    st = dict(enc.state_dict())
    for k in MONO2_BUFS:
        out["buf::" + k] = st[k].numpy()
    np.savez_compressed(os.path.join(OUT, "monodepth2_small.npz"), **out)
    print("monodepth2 loss", out["loss"], "ngrads", len(names), "depth range", float(depth.min()), float(depth.max()))


MIDAS_CASE = dict(seed=61, H=96, W=160, pairs=[(0, 1)])
MIDAS_FULL_GRADS = ("pretrained.layer1.0.weight", "pretrained.layer1.1.weight", "pretrained.layer1.4.0.conv2.weight",
                    "pretrained.layer2.0.downsample.1.weight", "pretrained.layer3.11.bn2.bias", "pretrained.layer4.2.bn3.weight",
                    "scratch.refinenet1.resConfUnit1.conv1.bias", "scratch.refinenet4.resConfUnit2.conv2.bias",
                    "scratch.output_conv.2.weight", "scratch.output_conv.4.weight", "scratch.output_conv.4.bias")
MIDAS_BUFS = ("pretrained.layer1.1.running_mean", "pretrained.layer1.1.running_var", "pretrained.layer2.0.downsample.1.running_var",
              "pretrained.layer3.22.bn3.running_mean", "pretrained.layer4.2.bn2.running_var")


def gen_midas():
    import warnings
    import torchvision
    warnings.filterwarnings("ignore")
    # blocks.py:27-29 pulls the trunk from torch.hub (no network here): torchvision defines the identical layer graph
    torch.hub.load = lambda *a, **k: torchvision.models.resnext101_32x8d(weights=None)
    from monodepth.midas_v2.midas_net import MidasNet
    c = MIDAS_CASE
    seed, H, W = c["seed"], c["H"], c["W"]
    sd = mo.midas_init_state(seed)
    net = MidasNet(non_negative=True)
    net.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    net.train()
    batch = synth.make_pair_batch(seed, c["pairs"], H, W)
    images = torch.tensor(batch["images"])
    # the adapter body, midas_v2_model.py:52-69 (the class itself downloads a checkpoint)
    x = images.reshape(-1, 3, H, W)
    mean = torch.Tensor([0.485, 0.456, 0.406]).reshape(1, -1, 1, 1)
    std = torch.Tensor([0.229, 0.224, 0.225]).reshape(1, -1, 1, 1)
    output = net((x - mean) / std)
    depth = output.reshape(1, 2, H, W).reciprocal()
    crit = ref_joint_loss(1.0, 1e-4, torch.float32)
    loss, meta = crit(depth, to_metadata(batch, torch.float32))
    loss.backward()
    out = {"disparity": output.detach().numpy(), "depth": depth.detach().numpy(), "loss": loss.detach().numpy()}
    params = dict(net.named_parameters())
    names, norms = [], []
    for k, p in params.items():
        if p.grad is None:
            continue
        names.append(k); norms.append(float(p.grad.double().norm()))
    out["grad_names"] = np.array(names); out["grad_norms"] = np.array(norms)
    for k in MIDAS_FULL_GRADS:
        out["grad::" + k] = params[k].grad.numpy()
    st = net.state_dict()
    for k in MIDAS_BUFS:
        out["buf::" + k] = st[k].numpy()
    np.savez_compressed(os.path.join(OUT, "midas_small.npz"), **out)
    print("midas loss", out["loss"], "ngrads", len(names), "disparity range", float(output.min()), float(output.max()))


FLOWMASK_CASES = {"a": (71, 24, 40, 1.0, 0.5), "b": (72, 17, 31, 1.5, 0.4)}     # seed, H, W, flow_thresh, color_thresh


def gen_flowmask():
    from utils import consistency
    out = {}
    for name, (seed, H, W, ft, ct) in FLOWMASK_CASES.items():
        flows, colors = fo.synthetic_pair(seed, H, W)
        masks = consistency.consistent_flow_masks(flows, colors, ft, ct)
        out[f"{name}_mask0"], out[f"{name}_mask1"] = masks[0], masks[1]
        X, Y = np.meshgrid(np.arange(W), np.arange(H))
        uv = np.stack((flows[0][..., 0] + X, flows[0][..., 1] + Y), -1)
        out[f"{name}_sample"] = consistency.sample(colors[1], uv)
        print("flowmask", name, "kept", float(masks[0].mean()), float(masks[1].mean()))
    np.savez_compressed(os.path.join(OUT, "flowmask.npz"), **out)


def gen_adam():
    import optimizer
    p0 = synth.normal(31, 1, (1003,), 0.1)
    p = torch.nn.Parameter(torch.tensor(p0.copy()))
    opt = optimizer.create("Adam", [p], 4e-4, betas=(0.9, 0.999))
    out = {}
    for t in range(6):
        g = synth.normal(31, 10 + t, (1003,), 10.0 ** (-t))      # widely varying gradient scales
        p.grad = torch.tensor(g)
        opt.step()
        out[f"p_{t}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "adam.npz"), **out)


def gen_finetune():
    import optimizer
    from monodepth.mannequin_challenge.models.hourglass import HourglassModel
    seed, H, W, steps = 41, 32, 48, 3
    sd = ho.mc_init_state(seed)
    net = HourglassModel(3)
    net.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    net.train()
    batch = synth.make_pair_batch(seed, [(0, 2)], H, W)
    images = torch.tensor(batch["images"])
    crit = ref_joint_loss(1.0, 0.1, torch.float32)
    opt = optimizer.create("Adam", net.parameters(), 4e-4, betas=(0.9, 0.999))
    meta = to_metadata(batch, torch.float32)
    losses = []
    for _ in range(steps):
        pred, _ = net.forward(images.reshape(-1, 3, H, W))
        depth = torch.exp(pred.reshape(1, 2, H, W))
        opt.zero_grad()
        loss, _m = crit(depth, meta)
        loss.backward()
        opt.step()
        losses.append(float(loss[0]))
    with torch.no_grad():
        pred, _ = net.forward(images.reshape(-1, 3, H, W))
    out = {"losses": np.array(losses), "final_depth": torch.exp(pred.reshape(1, 2, H, W)).numpy(),
           "pred_layer.bias": net.state_dict()["pred_layer.bias"].numpy(),
           "seq.0.weight": net.state_dict()["seq.0.weight"].numpy()}
    np.savez_compressed(os.path.join(OUT, "finetune_steps.npz"), **out)
    print("finetune losses", losses)


def gen_parameter():
    """JointLoss with lambda_parameter > 0 (joint_loss.py:34-39 -> parameter_loss.py:13-19) on the unmodified reference."""
    import loss.joint_loss as jl
    from oracle import parameter_oracle as po
    seed, lam = 31, 0.05
    inits, params = po.make_case(seed)
    p_init = [torch.tensor(a) for a in inits]
    ps = [torch.nn.Parameter(torch.tensor(a)) for a in params]
    jl._dtype = torch.float32
    opt = types.SimpleNamespace(lambda_view_baseline=0.0, lambda_reprojection=0.0, lambda_parameter=lam)
    loss, meta = jl.JointLoss(opt, p_init)(None, None, parameters=ps)
    loss.backward()
    out = {"lambda": np.float32(lam), "seed": np.int64(seed), "loss": loss.detach().numpy(),
           "parameter_loss": meta["parameter_loss"].detach().numpy()}
    for i, p in enumerate(ps):
        out[f"grad_{i}"] = p.grad.numpy()
    # combined with the consistency term (the total the fine-tuning loop sees)
    name = "geo_b2"
    cseed, pairs, H, W, stress, lr_, lb_ = CONSISTENCY_CASES[name]
    batch = synth.make_pair_batch(cseed, pairs, H, W, stress=stress)
    depth = torch.tensor(synth.synth_depth_pred(cseed, len(pairs), H, W))
    opt2 = types.SimpleNamespace(lambda_view_baseline=lb_, lambda_reprojection=lr_, lambda_parameter=lam)
    for p in ps:
        p.grad = None
    total, meta2 = jl.JointLoss(opt2, p_init)(depth, to_metadata(batch, torch.float32), parameters=ps)
    out["total_with_geo_b2"] = total.detach().numpy()
    out["meta_keys"] = np.array(sorted(meta2.keys()))
    np.savez_compressed(os.path.join(OUT, "parameter_loss.npz"), **out)
    print("parameter loss", float(loss), "total with consistency", float(total))


def main():
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--only", default=None, help="generate one fixture family: consistency|adam|hourglass|finetune|monodepth2|midas|flowmask|parameter")
    only = ap.parse_args().only
    ref_import.setup()
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    os.makedirs(OUT, exist_ok=True)
    gens = {"consistency": gen_consistency, "adam": gen_adam, "hourglass": gen_hourglass, "finetune": gen_finetune,
            "monodepth2": gen_monodepth2, "midas": gen_midas, "flowmask": gen_flowmask, "parameter": gen_parameter}
    for name, fn in gens.items():
        if only is None or only == name:
            fn()


if __name__ == "__main__":
    main()
