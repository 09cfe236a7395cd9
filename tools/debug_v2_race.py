#!/usr/bin/env python3
"""Diagnostic: one fine-tune step at 8 x 224 x 384 in several execution modes; prints loss and depth / gradient
differences against the serial first-generation run."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, hourglass_oracle as ho

DEV = "cuda:0"
seed, H, W, B = 29, 224, 384, 4
pairs = [(0, 1), (1, 2), (2, 4), (0, 3)]
batch = synth.make_pair_batch(seed, pairs, H, W)
t = lambda a: torch.tensor(a, device=DEV)


def run(env, use_graph, multi):
    for k, v in env.items():
        os.environ[k] = v
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    sd = {k: torch.tensor(np.asarray(v)) for k, v in ho.mc_init_state(seed).items()}
    model = MannequinChallengeModel(state_dict=sd).train()
    step = FineTuneStep(model, B, H, W, lr=4e-4, use_graph=use_graph)
    step.engine.multi_stream = multi
    step.load_batch(t(batch["images"]), [t(f) for f in batch["flows"]], [t(m) for m in batch["masks"]], t(batch["extrinsics"]), t(batch["intrinsics"]))
    loss = step.step()
    torch.cuda.synchronize()
    out = (float(loss), step.engine.depth.clone(), model.P.grad_flat.clone())
    for k in env:
        del os.environ[k]
    del step, model
    torch.cuda.empty_cache()
    return out


ref = run({"CVD_CONV2": "0"}, False, False)
print("v1 serial loss", ref[0])
cases = [("v2 serial", {}, False, False), ("v2 eager multi-stream", {}, False, True), ("v2 graph single-stream", {}, True, False),
         ("v2 graph multi-stream", {}, True, True), ("v2 graph multi-stream fork_first=0", {"CVD_FORK_FIRST": "0"}, True, True),
         ("v2 graph ms, wgrad v1", {"CVD_WGRAD2": "0"}, True, True),
         ("v1 graph multi-stream", {"CVD_CONV2": "0"}, True, True)]
for name, env, g, m in cases:
    for rep in range(2):
        l, d, gr = run(env, g, m)
        dd = ((d - ref[1]).abs() / ref[1]).max().item()
        gg = float((gr.double() - ref[2].double()).norm() / ref[2].double().norm())
        print(f"{name:40s} rep{rep}: loss {l:.6f} (rel {abs(l-ref[0])/abs(ref[0]):.2e})  depth max rel {dd:.2e}  grad rel-L2 {gg:.2e}  nan={bool(torch.isnan(d).any())}")
