#!/usr/bin/env python3
"""Diagnostic: engine vs oracle (CPU fp32) parameter gradients on one small pair, per tensor and per tap."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, hourglass_oracle as ho, consistency_oracle as co
from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
from consistent_depth_b200.loss.joint_loss import JointLoss

seed, H, W = 21, 32, 48
DEV = "cuda:0"
sd = {k: torch.tensor(np.asarray(v)) for k, v in ho.mc_init_state(seed).items()}
model = MannequinChallengeModel(state_dict=sd).train()
batch = synth.make_pair_batch(seed, [(0, 1)], H, W)
t = lambda a: torch.tensor(a, device=DEV)
meta = {"extrinsics": t(batch["extrinsics"]), "intrinsics": t(batch["intrinsics"]),
        "geometry_consistency": {"indices": t(batch["indices"]), "flows": [t(f) for f in batch["flows"]], "masks": [t(m) for m in batch["masks"]]}}
depth = model(t(batch["images"]), None)
model.parameters(); model.P.grad_flat.zero_()
loss, _ = JointLoss(types.SimpleNamespace(lambda_view_baseline=0.1, lambda_reprojection=1.0, lambda_parameter=0))(depth, meta)
loss.backward(); torch.cuda.synchronize()
P, buf = ho.to_torch(ho.mc_init_state(seed), requires_grad=True)
tc = lambda a: torch.tensor(a)
d = ho.estimate_depth(tc(batch["images"]), P, buf)
l, _ = co.consistency_loss(d, tc(batch["extrinsics"]), tc(batch["intrinsics"]), [tc(f) for f in batch["flows"]], [tc(m) for m in batch["masks"]], 1.0, 0.1)
l.backward()
rows = []
for k in ho.trainable_keys():
    ref = P[k].grad.double().numpy(); got = model.P._g(k).cpu().double().numpy()
    if np.abs(ref).max() < 1e-6: continue
    err = np.abs(got - ref).max() / np.abs(ref).max()
    cos = (got * ref).sum() / np.sqrt((got * got).sum() * (ref * ref).sum() + 1e-300)
    rows.append((err, cos, k, ref.shape))
rows.sort(reverse=True)
for r in rows[:40]: print(f"{r[0]:.3e} cos {r[1]:.5f} {r[2]} {r[3]}")
k = "seq.3.list.1.0.convs.3.3.weight"
ref = P[k].grad.double().numpy(); got = model.P._g(k).cpu().double().numpy()
e = np.abs(got - ref)
print("per-tap max err (11x11):"); print(np.round(e.max(axis=(0, 1)) / np.abs(ref).max(), 3))
print("per-co max err:", np.round(e.max(axis=(1, 2, 3)) / np.abs(ref).max(), 3))
print("ratio got/ref at max:", got.flat[np.abs(ref).argmax()], ref.flat[np.abs(ref).argmax()])

# ---- localise: gradient wrt the raw output of A-inception's 11x11 conv (pre-BN), per channel
import torch.nn.functional as F
P, buf = ho.to_torch(ho.mc_init_state(seed), requires_grad=True)
cap = {}
orig = F.conv2d
store = {}
def hooked(x, w, b=None, **kw):
    y = orig(x, w, b, **kw)
    if w.shape == (16, 64, 11, 11) and "y" not in store:
        y.retain_grad(); store["y"] = y
    return y
F.conv2d = hooked; ho.F.conv2d = hooked
d = ho.estimate_depth(tc(batch["images"]), P, buf)
l, _ = co.consistency_loss(d, tc(batch["extrinsics"]), tc(batch["intrinsics"]), [tc(f) for f in batch["flows"]], [tc(m) for m in batch["masks"]], 1.0, 0.1)
l.backward()
F.conv2d = orig; ho.F.conv2d = orig
gref = store["y"].grad.permute(0, 2, 3, 1).double()          # (N,H,W,16)
eng = model.engine(2, H, W)
bufA, off, cout = eng.raw_outputs["seq.3.list.1.0.convs.3.3"]
rec = [r for r in eng.recs if r[0] == "inc" and r[2] == "seq.3.list.1.0"][0]
one = rec[6]
x = bufA[..., off:off + 16].cpu().double(); dy = one.dbuf[..., off:off + 16].cpu().double()
a = one.a[off:off + 16].cpu().double(); b = one.b[off:off + 16].cpu().double(); bw = one.bw[off:off + 16].cpu().double()
y = x * a + b
g = torch.where(y > 0, dy, torch.zeros_like(dy))
G = bw[:, 0] * g - bw[:, 1] - bw[:, 2] * y
print("dx_raw per-channel max err / max:", np.round(((G - gref).abs().amax((0, 1, 2)) / gref.abs().amax()).numpy(), 4))
print("bw ch12:", bw[12].numpy(), " ch11:", bw[11].numpy())
print("mean g per ch:", np.round(g.mean((0, 1, 2)).numpy(), 6))
print("frac y>0 per ch:", np.round((y > 0).double().mean((0, 1, 2)).numpy(), 3))
