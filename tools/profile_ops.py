#!/usr/bin/env python3
"""Per-op CUDA-event timing of one un-graphed fine-tune step (fwd + loss + bwd), grouped by op kind and layer shape."""
import collections
import json
import os
import sys

import torch

os.environ.setdefault("CVD_MULTI_STREAM", "0")     # per-op times: one kernel at a time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200 import ops  # noqa: E402
from consistent_depth_b200.fine_tune_step import FineTuneStep  # noqa: E402
from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel  # noqa: E402
from consistent_depth_b200.synthetic import SyntheticVideo  # noqa: E402

H, W, BS = 224, 384, 4
prec = int(os.environ.get("PREC", "3"))
dev = torch.device("cuda:0")
model = MannequinChallengeModel(precision=prec)
video = SyntheticVideo(10, H, W, dev, pairs=[(0, 1), (1, 3), (2, 6), (4, 5)])
step = FineTuneStep(model, BS, H, W, lr=4e-4, use_graph=False)
b = video.batch([0, 1, 2, 3])
step.load_batch(b["images"], b["flows"], b["masks"], b["extrinsics"], b["intrinsics"])
recs = []


def wrap(name, describe):
    orig = getattr(ops, name)

    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig(*a, **k)
        e1.record()
        recs.append((name, describe(*a, **k), e0, e1))
        return r
    setattr(ops, name, f)


wrap("conv", lambda s, pk, bias, d, N, h, w, cin, cout, k, p=3, fl=0, bn=None: (f"{'dgrad' if s.mode else 'fwd'} {cin}->{cout} k{k} {h}x{w}", 2.0 * k * k * cin * cout * N * h * w))
wrap("conv_wgrad", lambda g, x, dw, N, h, w, cin, cout, k, p=3: (f"wgrad {cin}->{cout} k{k} {h}x{w}", 2.0 * k * k * cin * cout * N * h * w))
wrap("bn_stats", lambda x, off, C, npix, *a, **k: (f"C{C} npix{npix}", 4.0 * C * npix))
wrap("bn_bwd_reduce", lambda x, off, C, dy, npix, *a, **k: (f"C{C} npix{npix}", 8.0 * C * npix))
wrap("pool_fwd", lambda xv, a, b, r, p, N, h, w, C: (f"C{C} {h}x{w}", 5.0 * C * N * h * w))
wrap("pool_bwd", lambda dp, xv, acc, N, h, w, C: (f"C{C} {h}x{w}", 9.0 * C * N * h * w))
wrap("merge_up_fwd", lambda *a: (f"C{a[-1]} {a[-3]}x{a[-2]}", 9.0 * a[-1] * a[-4] * a[-3] * a[-2]))
wrap("merge_up_bwd", lambda *a: (f"C{a[-1]} {a[-3]}x{a[-2]}", 9.0 * a[-1] * a[-4] * a[-3] * a[-2]))
wrap("pack_weights", lambda w, t=False, p=3, out=None: (f"{tuple(w.shape)}", 0.0))
wrap("image_to_nhwc4", lambda *a: ("", 0.0))
wrap("dlogdepth", lambda *a: ("", 0.0))

for rep in range(3):
    recs.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step._snapshot_and_restore(step._fwd_bwd)
    t1.record()
    torch.cuda.synchronize()
print("fwd+bwd wall (eager) ms:", t0.elapsed_time(t1))
kind = collections.OrderedDict()
shape = collections.OrderedDict()
for name, (desc, work), e0, e1 in recs:
    ms = e0.elapsed_time(e1)
    cat = name if name != "conv" else ("conv_" + desc.split()[0])
    kind.setdefault(cat, [0.0, 0, 0.0]); kind[cat][0] += ms; kind[cat][1] += 1; kind[cat][2] += work
    key = f"{cat} {desc.split(' ', 1)[1] if name.startswith('conv') else desc}"
    shape.setdefault(key, [0.0, 0, 0.0]); shape[key][0] += ms; shape[key][1] += 1; shape[key][2] += work
tot = sum(v[0] for v in kind.values())
print(f"sum of op times {tot:.2f} ms over {len(recs)} ops")
for k, v in sorted(kind.items(), key=lambda kv: -kv[1][0]):
    unit = f"{v[2] / (v[0] * 1e-3) / 1e12:.1f} TFLOP/s" if k.startswith("conv") else f"{v[2] / (v[0] * 1e-3) / 1e9:.0f} GB/s"
    print(f"{k:16s} {v[0]:8.2f} ms  n={v[1]:4d}  {unit}")
print("--- top shapes")
for k, v in sorted(shape.items(), key=lambda kv: -kv[1][0])[:45]:
    unit = f"{v[2] / (v[0] * 1e-3) / 1e12:.1f} TF/s" if k.startswith("conv") else f"{v[2] / (v[0] * 1e-3) / 1e9:.0f} GB/s"
    print(f"{k:44s} {v[0]:8.3f} ms n={v[1]:3d} avg {v[0] / v[1] * 1e3:8.1f} us  {unit}")
if len(sys.argv) > 1:
    json.dump({k: v for k, v in shape.items()}, open(sys.argv[1], "w"), indent=1)
