#!/usr/bin/env python3
"""Timing of the dominant MiDaS-v2 (ResNeXt-101 32x8d encoder) convolution launches at the bench size (2 frames, 384 x 672):
weight gradients of the wide 1x1 convs and of the decoder 3x3 convs, forward of the same, the chunked grouped conv.
CUDA events, L2 flushed between repetitions.   python tools/midas_microbench.py [--out gpurun_out/midas_mb.json]"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200 import ops

WG = [(1024, 1024, 1, 24, 42), (512, 512, 1, 48, 84), (2048, 2048, 1, 12, 21), (256, 256, 1, 96, 168),
      (256, 256, 3, 96, 168), (256, 128, 3, 192, 336), (256, 256, 3, 48, 84), (1024, 2048, 1, 12, 21)]


def time_it(run, reps, flush):
    run(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev, N = "cuda:0", 2
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=dev)
    res = []
    for (cin, cout, k, H, W) in WG:
        if a.only and a.only != f"{cin},{cout},{k},{H},{W}":
            continue
        x = torch.rand(N, H, W, cin, device=dev) - 0.5
        g = torch.rand(N, H, W, cout, device=dev) - 0.5
        w = (torch.rand(cout, cin, k, k, device=dev) - 0.5) * 0.05
        dw = torch.zeros_like(w)
        y = torch.empty(N, H, W, cout, device=dev)
        xs, gs = ops.make_src(ops.View(x, 0)), ops.make_src(ops.View(g, 0))
        pk = ops.pack_weights(w, False, 3)
        d = ops.make_dst(ops.View(y, 0))
        r = {"cin": cin, "cout": cout, "k": k, "H": H, "W": W,
             "wgrad_us": time_it(lambda: ops.conv_wgrad(gs, xs, dw, N, H, W, cin, cout, k, 3), a.reps, flush),
             "fwd_us": time_it(lambda: ops.conv(xs, pk, None, d, N, H, W, cin, cout, k, 3, 0), a.reps, flush)}
        r["gflop"] = 2.0 * N * H * W * cin * cout * k * k / 1e9
        print(json.dumps(r), flush=True)
        res.append(r)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
