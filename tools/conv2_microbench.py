#!/usr/bin/env python3
"""Per-layer timing of the hourglass convolutions at the bench size: first-generation kernel (conv_tc, transform + split
in the producer warps) vs second-generation (cvd_prep_operand once + conv2: TMA-fed, kx-fused), forward and dgrad shapes.
Environment knobs of conv2.cu (CVD2_MT, CVD2_GMAX, CVD2_WS, CVD2_NO_RESIDENT) can be swept with --sweep.

    python tools/conv2_microbench.py --out gpurun_out/conv2_mb.json [--sweep]
"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200 import ops

# (GEMM cin, GEMM cout, k, H, W, count per step, tag) -- forward shapes of HourglassModel(3) at 8 x 224 x 384 and their dgrads
FWD = [
    (64, 16, 11, 224, 384, 1), (64, 16, 7, 224, 384, 1), (64, 16, 3, 224, 384, 1), (128, 208, 1, 224, 384, 1),
    (64, 32, 11, 112, 192, 1), (64, 32, 7, 112, 192, 2), (64, 32, 3, 112, 192, 2), (64, 32, 5, 112, 192, 1),
    (32, 32, 7, 112, 192, 3), (32, 32, 5, 112, 192, 3), (32, 32, 3, 112, 192, 3),
    (32, 16, 11, 112, 192, 1), (32, 16, 7, 112, 192, 1), (32, 16, 3, 112, 192, 1),
    (128, 128, 1, 112, 192, 3), (128, 224, 1, 112, 192, 2), (128, 112, 1, 112, 192, 1),
    (32, 64, 7, 56, 96, 3), (32, 64, 5, 56, 96, 3), (32, 64, 3, 56, 96, 3), (64, 64, 11, 56, 96, 1), (64, 64, 7, 56, 96, 1),
    (64, 64, 3, 56, 96, 1), (32, 32, 7, 56, 96, 2), (32, 32, 5, 56, 96, 2), (32, 32, 3, 56, 96, 2),
    (128, 160, 1, 56, 96, 1), (256, 160, 1, 56, 96, 2), (256, 256, 1, 56, 96, 1), (128, 128, 1, 56, 96, 1), (256, 128, 1, 56, 96, 1),
    (32, 64, 7, 28, 48, 5), (32, 64, 5, 28, 48, 5), (32, 64, 3, 28, 48, 5), (256, 160, 1, 28, 48, 5), (256, 256, 1, 28, 48, 1),
    (32, 64, 7, 14, 24, 3), (32, 64, 5, 14, 24, 3), (32, 64, 3, 14, 24, 3), (256, 160, 1, 14, 24, 3),
]


def time_it(run, reps):
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench_shape(cin, cout, k, H, W, N, reps, do_v1=True):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(N, H, W, cin, device=dev, generator=g) - 0.5
    w = (torch.rand(cout, cin, k, k, device=dev, generator=g) - 0.5) * 0.05
    sa = torch.rand(cin, device=dev) + 0.5; sb = torch.rand(cin, device=dev) - 0.5
    y = torch.empty(N, H, W, cout, device=dev)
    src, dst = ops.make_src(ops.View(x), sa, sb, True), ops.make_dst(ops.View(y))
    out = {}
    if do_v1:
        pk = ops.pack_weights(w, False, 3)
        out["v1_ms"] = time_it(lambda: ops.conv(src, pk, None, dst, N, H, W, cin, cout, k, 3), reps)
    z = ops.z_alloc(N, cin, H, W, dev)
    out["prep_ms"] = time_it(lambda: ops.prep_operand(src, cin, z), reps)
    pk2 = ops.conv2_pack(w, False)
    out["v2_ms"] = time_it(lambda: ops.conv2(z, 0, pk2, None, dst, N, H, W, cin, cout, k, 0, None), reps)
    return out


def bench_wgrad(cin, cout, k, H, W, N, reps):
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.rand(N, H, W, cin, device=dev, generator=g) - 0.5
    gy = torch.rand(N, H, W, cout, device=dev, generator=g) - 0.5
    xr = torch.rand(N, H, W, cout, device=dev, generator=g) - 0.5
    sa = torch.rand(cin, device=dev) + 0.5; sb = torch.rand(cin, device=dev) - 0.5
    ga = torch.rand(cout, device=dev) + 0.5; gb = torch.rand(cout, device=dev) - 0.5; bw = torch.rand(cout, 4, device=dev)
    dw = torch.zeros(cout, cin, k, k, device=dev)
    gsrc = ops.make_src(ops.View(xr), ga, gb, True, dy=ops.View(gy), bw=bw)
    xsrc = ops.make_src(ops.View(x), sa, sb, True)
    out = {"v1_ms": time_it(lambda: ops.conv_wgrad(gsrc, xsrc, dw, N, H, W, cin, cout, k, 3), reps)}
    xz, gz = ops.z_alloc(N, cin, H, W, dev), ops.z_alloc(N, cout, H, W, dev)
    ops.prep_operand(xsrc, cin, xz); ops.prep_operand(gsrc, cout, gz)
    ok = ops.conv2_wgrad(xz, 0, gz, 0, dw, N, H, W, cin, cout, k)
    out["v2_ms"] = time_it(lambda: ops.conv2_wgrad(xz, 0, gz, 0, dw, N, H, W, cin, cout, k), reps) if ok else None
    return out


def main_wgrad(a):
    rows, tot = [], {"v1": 0.0, "v2": 0.0}
    for (ci, co, k, H, W, cnt) in FWD:
        r = bench_wgrad(ci, co, k, H, W, a.N, a.reps)
        gf = 2.0 * k * k * ci * co * a.N * H * W
        v2 = r["v2_ms"] if r["v2_ms"] is not None else r["v1_ms"]
        rows.append({"kind": "wgrad", "cin": ci, "cout": co, "k": k, "H": H, "W": W, "count": cnt, **r})
        tot["v1"] += cnt * r["v1_ms"]; tot["v2"] += cnt * v2
        print(f"wgrad {ci:3d}->{co:3d} k{k:2d} {H:3d}x{W:3d} x{cnt}: v1 {r['v1_ms']:.3f} ms ({gf / r['v1_ms'] / 1e9:.0f} TF)  "
              f"v2 {v2:.3f} ms ({gf / v2 / 1e9:.0f} TF){'' if r['v2_ms'] is not None else ' [unsupported: v1]'}", flush=True)
    print("wgrad per-step totals (count-weighted):", tot)
    json.dump({"N": a.N, "rows": rows, "totals_ms": tot}, open(a.out, "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/conv2_mb.json")
    ap.add_argument("--N", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--only-big", action="store_true")
    ap.add_argument("--wgrad", action="store_true", help="weight-gradient kernels (conv_wgrad vs conv2_wgrad) instead")
    ap.add_argument("--one", default=None, help="kind,cin,cout,k,H,W (kind = fwd | wgrad): run only this shape (for ncu)")
    a = ap.parse_args()
    if a.one:
        kind, ci, co, k, H, W = a.one.split(",")
        ci, co, k, H, W = int(ci), int(co), int(k), int(H), int(W)
        r = bench_wgrad(ci, co, k, H, W, a.N, a.reps) if kind == "wgrad" else bench_shape(ci, co, k, H, W, a.N, a.reps)
        print(a.one, r)
        return
    if a.wgrad:
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        return main_wgrad(a)
    shapes = []
    for (ci, co, k, H, W, cnt) in FWD:
        shapes.append(("fwd", ci, co, k, H, W, cnt))
        shapes.append(("dgrad", co, ci, k, H, W, cnt))
    if a.only_big:
        shapes = [s for s in shapes if s[4] >= 112]
    rows = []
    tot = {"v1": 0.0, "v2": 0.0, "prep": 0.0}
    for (kind, ci, co, k, H, W, cnt) in shapes:
        r = bench_shape(ci, co, k, H, W, a.N, a.reps)
        gf = 2.0 * k * k * ci * co * a.N * H * W
        row = {"kind": kind, "cin": ci, "cout": co, "k": k, "H": H, "W": W, "count": cnt, **r,
               "v1_tflops": gf / r["v1_ms"] / 1e9, "v2_tflops": gf / r["v2_ms"] / 1e9}
        if a.sweep and H >= 56:
            for name, env in (("mt2", {"CVD2_MT": "2"}), ("mt1", {"CVD2_MT": "1"}), ("g4", {"CVD2_GMAX": "4"}), ("g2", {"CVD2_GMAX": "2"}),
                              ("nores", {"CVD2_NO_RESIDENT": "1"})):
                if name == "nores" and k != 1:
                    continue
                if name in ("g4", "g2") and k == 1:
                    continue
                os.environ.update(env)
                try:
                    row[name + "_ms"] = bench_shape(ci, co, k, H, W, a.N, a.reps, do_v1=False)["v2_ms"]
                except Exception as e:      # a knob combination the kernel rejects
                    row[name + "_ms"] = None
                for kk in env:
                    del os.environ[kk]
        rows.append(row)
        tot["v1"] += cnt * r["v1_ms"]; tot["v2"] += cnt * r["v2_ms"]; tot["prep"] += cnt * r["prep_ms"]
        extra = " ".join(f"{n}={row[n]:.3f}" for n in row if n.endswith("_ms") and n not in ("v1_ms", "v2_ms", "prep_ms") and row[n] is not None)
        print(f"{kind:5s} {ci:3d}->{co:3d} k{k:2d} {H:3d}x{W:3d} x{cnt}: v1 {r['v1_ms']:.3f} ms ({row['v1_tflops']:.0f} TF)  v2 {r['v2_ms']:.3f} ms "
              f"({row['v2_tflops']:.0f} TF)  prep {r['prep_ms']:.3f}  {extra}", flush=True)
    print("per-step totals (count-weighted):", tot)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump({"N": a.N, "rows": rows, "totals_ms": tot}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
