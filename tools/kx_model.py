#!/usr/bin/env python3
"""Static tensor-issue model of the hourglass convs: tcgen05.mma count and a cycle estimate per layer for the
per-tap kernel (N = Cout) and for the kx-fused column conv (N = k*Cout), forward pass, 224x384, 8 frames, bf16x3.
Cycle model per MMA (M = 128, K = 16): max(issue floor 32 clk, 128*N*16 MACs / 4000 MAC/clk/SM).  CPU only; prints markdown."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent_depth_b200.monodepth import mc_arch  # noqa: E402

SMS, CLK = 148, 1.9e9


def mma_clk(n):
    return max(32.0, 128.0 * n * 16 / 4000.0)


def walk(node, h, w, out):
    if node[0] == "inc":
        cfg = node[2]
        cin = node[1]
        A = sum(c[1] for c in cfg[1:])
        out.append((cin, cfg[0][0] + A, 1, h, w))
        for k, a, b in cfg[1:]:
            out.append((a, b, k, h, w))
    elif node[0] == "chan":
        for branch in node[1:]:
            hh, ww = h, w
            for op in branch:
                if op[0] == "pool":
                    hh, ww = hh // 2, ww // 2
                elif op[0] in ("inc", "chan"):
                    walk(op, hh, ww, out)


def main():
    N, H, W = 8, 224, 384
    layers = [(3, 128, 7, H, W)]
    walk(mc_arch.structure(), H, W, layers)
    layers.append((64, 1, 3, H, W))
    agg = collections.OrderedDict()
    for cin, cout, k, h, w in layers:
        key = (cin, cout, k, h, w)
        agg[key] = agg.get(key, 0) + 1
    rows, tot_tap, tot_best = [], 0.0, 0.0
    for (cin, cout, k, h, w), cnt in agg.items():
        px = N * h * w
        kb = -(-cin // 16)
        coutp = -(-cout // 16) * 16
        mt = px / 128.0
        tap_mmas = mt * k * k * kb * 3
        t_tap = tap_mmas * mma_clk(coutp) / SMS / CLK * 1e3
        if k >= 3 and k * coutp <= 256:
            mtx = N * h * (w + k - 1) / 128.0
            kx_mmas = mtx * k * kb * 3
            t_kx = kx_mmas * mma_clk(k * coutp) / SMS / CLK * 1e3
            extra = 2 * N * h * (w + k - 1) * k * coutp * 4 / 6.5e12 * 1e3      # D written + read through HBM
        else:
            kx_mmas, t_kx, extra = None, None, 0.0
        best = min(t_tap, (t_kx + extra) if t_kx is not None else 1e9)
        tot_tap += cnt * t_tap
        tot_best += cnt * best
        rows.append((cnt * t_tap, cin, cout, k, h, w, cnt, tap_mmas, t_tap, kx_mmas, t_kx, extra))
    rows.sort(reverse=True)
    print("| Cin→Cout k | map | × | per-tap MMAs (N) | est. ms | kx-fused MMAs (N) | est. ms (+D traffic) |")
    print("|---|---|---|---|---|---|---|")
    for _, cin, cout, k, h, w, cnt, tm, tt, km, tk, ex in rows[:18]:
        kxs = f"{km / 1e6:.2f} M ({k * (-(-cout // 16) * 16)})" if km else "—"
        tks = f"{tk:.3f} (+{ex:.3f})" if tk is not None else "—"
        print(f"| {cin}→{cout} k{k} | {h}×{w} | {cnt} | {tm / 1e6:.2f} M ({-(-cout // 16) * 16}) | {tt:.3f} | {kxs} | {tks} |")
    print(f"\nforward pass, tensor-issue estimate: per-tap {tot_tap:.2f} ms; best of per-tap / kx-fused per layer {tot_best:.2f} ms "
          f"(measured forward conv time: ~8.7 ms of the 28.8 ms step)")


if __name__ == "__main__":
    main()
