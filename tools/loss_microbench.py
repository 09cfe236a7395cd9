#!/usr/bin/env python3
"""BASELINE config 5: fused reproject+consistency-loss microbench (fwd+bwd and fwd-only).

Times `cvd_consistency_fwd_bwd` (memsets + kernel + finalize, i.e. the whole C-ABI call) with
CUDA events on the launching stream; L2 is flushed between repetitions by writing a 256 MiB
buffer.  Achieved GB/s = algorithmic bytes (40 B/px/pair fwd+bwd, 32 fwd-only) / time.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200.utils.geometry import fused_consistency  # noqa: E402
from consistent_depth_b200 import _lib  # noqa: E402
import ctypes as C  # noqa: E402


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def make_inputs(B, H, W, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    y, x = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                          torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    depth = 2.0 + 0.5 * torch.sin(x / W * 6.0) * torch.cos(y / H * 4.0)
    depth = depth[None, None].expand(B, 2, H, W).contiguous() * (1 + 0.1 * torch.rand(B, 2, H, W, device=dev, generator=g))
    flows = [(3.0 * torch.sin(x / 37.0 + k) [None, None] + torch.randn(B, 2, H, W, device=dev, generator=g) * 0.5).contiguous()
             for k in range(2)]
    masks = [(torch.rand(B, 1, H, W, device=dev, generator=g) < 0.7).float() for _ in range(2)]
    extr = torch.zeros(B, 2, 3, 4, device=dev)
    extr[..., :3] = torch.eye(3, device=dev)
    extr[:, 1, 0, 3] = 0.05
    intr = torch.tensor([0.8 * W, 0.8 * W, (W - 1) / 2, (H - 1) / 2], device=dev).expand(B, 2, 4).contiguous()
    return depth, flows, masks, extr, intr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="224x384,384x672,540x960,720x1280,1080x1920")
    ap.add_argument("--batches", default="1,4,16,64")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak, src = peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for size in args.sizes.split(","):
        H, W = [int(v) for v in size.split("x")]
        for B in [int(b) for b in args.batches.split(",")]:
            if 40.0 * H * W * B > 12e9:
                continue
            depth, flows, masks, extr, intr = make_inputs(B, H, W, dev)
            for mode, want_grad, bpp in (("fwd+bwd", True, 40), ("fwd", False, 32)):
                ts = []
                for it in range(args.warmup + args.reps):
                    if not args.no_flush:
                        flush.fill_(it & 0xFF)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fused_consistency(depth, flows, masks, extr, intr, 1.0, 0.1, want_grad=want_grad)
                    e1.record()
                    e1.synchronize()
                    if it >= args.warmup:
                        ts.append(e0.elapsed_time(e1))
                t = float(np.median(ts))
                gbs = bpp * H * W * B / (t * 1e-3) / 1e9
                row = {"H": H, "W": W, "B": B, "mode": mode, "ms": t, "GBps": gbs, "frac_of_peak": gbs / peak,
                       "peak": peak, "peak_src": src, "l2_flush": not args.no_flush}
                rows.append(row)
                print(json.dumps(row), flush=True)
    if args.out:
        json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
