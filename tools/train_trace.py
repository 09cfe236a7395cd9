#!/usr/bin/env python3
"""Loss / depth-range trajectory of the fused fine-tune step on the bench workload (diagnostic)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200.fine_tune_step import FineTuneStep
from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
from consistent_depth_b200.synthetic import SyntheticVideo
H, W, BS = 224, 384, 4
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
import math
from consistent_depth_b200.monodepth.mannequin_challenge_model import default_init_state
sd = default_init_state(0); sd['pred_layer.weight'] = sd['pred_layer.weight'] * 0.1; sd['pred_layer.bias'] = torch.full((1,), math.log(2.0))
model = MannequinChallengeModel(state_dict=sd)
video = SyntheticVideo(50, H, W, dev, seed=1236)
step = FineTuneStep(model, BS, H, W, lr=model.learning_rate, use_graph=(os.environ.get("GRAPH", "1") == "1"))
g = torch.Generator().manual_seed(0)
order = torch.randperm(len(video.pairs), generator=g).tolist()
for it in range(steps):
    ids = [order[(it * BS + j) % len(order)] for j in range(BS)]
    b = video.batch(ids)
    step.load_batch(b["images"], b["flows"], b["masks"], b["extrinsics"], b["intrinsics"])
    l = step.step()
    d = step.engine.depth
    print(it, f"loss {float(l):.4f} pair-losses r {step.pair_losses[0].mean().item():.3f} d {step.pair_losses[1].mean().item():.3f}"
          f" depth [{d.min().item():.3g}, {d.max().item():.3g}] mean {d.mean().item():.3g}")
