#!/usr/bin/env python3
"""Time one conv layer (fwd / dgrad-style / wgrad) of the hourglass at full size; usable under ncu."""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=16)
ap.add_argument("--k", type=int, default=11); ap.add_argument("--N", type=int, default=8)
ap.add_argument("--H", type=int, default=224); ap.add_argument("--W", type=int, default=384)
ap.add_argument("--mode", default="fwd", choices=["fwd", "wgrad"])
ap.add_argument("--prec", type=int, default=3); ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--bn", type=int, default=1, help="apply BN affine+relu on load (fwd) / BN backward on load (wgrad G)")
a = ap.parse_args()
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
cx = (a.cin + 3) // 4 * 4
x = torch.rand(a.N, a.H, a.W, cx, device=dev, generator=g) - 0.5
w = (torch.rand(a.cout, a.cin, a.k, a.k, device=dev, generator=g) - 0.5) * 0.05
sa = torch.rand(cx, device=dev) + 0.5; sb = torch.rand(cx, device=dev) - 0.5
flops = 2.0 * a.k * a.k * a.cin * a.cout * a.N * a.H * a.W
if a.mode == "fwd":
    y = torch.empty(a.N, a.H, a.W, max(4, a.cout), device=dev)
    pk = ops.pack_weights(w, False, a.prec)
    src = ops.make_src(ops.View(x), sa if a.bn else None, sb if a.bn else None, bool(a.bn)); dst = ops.make_dst(ops.View(y))
    run = lambda: ops.conv(src, pk, None, dst, a.N, a.H, a.W, a.cin, a.cout, a.k, a.prec)
else:
    cg = (a.cout + 3) // 4 * 4
    gy = torch.rand(a.N, a.H, a.W, cg, device=dev, generator=g) - 0.5
    xr = torch.rand(a.N, a.H, a.W, cg, device=dev, generator=g) - 0.5
    ga = torch.rand(cg, device=dev) + 0.5; gb = torch.rand(cg, device=dev) - 0.5; bw = torch.rand(cg, 4, device=dev)
    dw = torch.zeros(a.cout, a.cin, a.k, a.k, device=dev)
    gsrc = ops.make_src(ops.View(xr), ga, gb, True, dy=ops.View(gy), bw=bw) if a.bn else ops.make_src(ops.View(gy))
    xsrc = ops.make_src(ops.View(x), sa if a.bn else None, sb if a.bn else None, bool(a.bn))
    run = lambda: ops.conv_wgrad(gsrc, xsrc, dw, a.N, a.H, a.W, a.cin, a.cout, a.k, a.prec)
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"{a.mode} {a.cin}->{a.cout} k{a.k} {a.N}x{a.H}x{a.W} prec{a.prec}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
