import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import flowmask_oracle as fo
from oracle.make_golden import FLOWMASK_CASES
from consistent_depth_b200.utils.consistency import consistent_flow_masks, consistent_flow_masks_batched
for name, (seed, H, W, ft, ct) in FLOWMASK_CASES.items():
    flows, colors = fo.synthetic_pair(seed, H, W)
    want = fo.consistent_flow_masks(flows, colors, ft, ct)
    print(name, H, W, ft, ct, "oracle ones", [float(w.mean()) for w in want])
    for rep in range(4):
        m = consistent_flow_masks(flows, colors, ft, ct)
        print("  rep", rep, "kernel ones", [float(x.mean()) for x in m], "mismatch", [float((m[d] != want[d]).mean()) for d in range(2)])
    dev = torch.device("cuda:0")
    f = torch.tensor(np.stack([np.asarray(x, np.float32).transpose(2, 0, 1) for x in flows])[None], device=dev)
    c = torch.tensor(np.stack([np.asarray(x, np.float32).transpose(2, 0, 1) for x in colors])[None], device=dev)
    torch.cuda.synchronize()
    for rep in range(3):
        mm = consistent_flow_masks_batched(f, c, ft, ct)[0]
        torch.cuda.synchronize()
        print("  batched rep", rep, [float((mm[d].cpu().numpy() > 0.5).mean()) for d in range(2)],
              "mismatch", [float(((mm[d].cpu().numpy() > 0.5) != want[d]).mean()) for d in range(2)])
