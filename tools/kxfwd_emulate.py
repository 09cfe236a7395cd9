#!/usr/bin/env python3
"""numpy emulation of the planned kx-fused forward conv: validates the tiling / packing / shifted-sum
epilogue index arithmetic against torch's conv2d before any CUDA is written.  CPU only.

    per CTA tile (R rows x WT cols of one image):
      window rows y0-pad .. y0+R-1+pad, window slots x' = 0 .. WP-1 (image col x0 - pad + x'), WP = WT + k - 1 <= 128
      for ky, for each 16-channel k-block:  D[r][x'][kx*Cout + co] += A[r + ky][x'][c] * B[ky][c][kx*Cout + co]
      epilogue: out[y0 + r][x0 + x][co] = sum_kx D[r][x + kx][kx*Cout + co]        (x < WT)
"""
import numpy as np
import torch
import torch.nn.functional as F


def conv_kx_fused(x, w, R=2, WT=96):
    """x (N,Cin,H,W), w (Cout,Cin,k,k) -> (N,Cout,H,W), "same" zero padding, emulating the tile loop."""
    N, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    pad = k // 2
    WP = WT + k - 1
    assert WP <= 128
    cin_p = (Cin + 15) // 16 * 16
    # B operand per ky: [cin_p][k*Cout], column n = kx*Cout + co
    Bm = np.zeros((k, cin_p, k * Cout))
    for ky in range(k):
        for kx in range(k):
            Bm[ky, :Cin, kx * Cout:(kx + 1) * Cout] = w[:, :, ky, kx].T
    out = np.zeros((N, Cout, H, W))
    for n in range(N):
        for y0 in range(0, H, R):
            for x0 in range(0, W, WT):
                # stage the window: rows y0-pad .. y0+R-1+pad, 128 slots (slots >= WP and out-of-image pixels are zero)
                win = np.zeros((R + k - 1, 128, cin_p))
                for r in range(R + k - 1):
                    iy = y0 - pad + r
                    if not 0 <= iy < H:
                        continue
                    for s in range(WP):
                        ix = x0 - pad + s
                        if 0 <= ix < W:
                            win[r, s, :Cin] = x[n, :, iy, ix]
                D = np.zeros((R, 128, k * Cout))                      # one M = 128 accumulator per output row
                for r in range(R):
                    for ky in range(k):
                        for kb in range(cin_p // 16):                 # one MMA (x3 for the bf16 split) per (ky, k-block)
                            D[r] += win[r + ky, :, kb * 16:(kb + 1) * 16] @ Bm[ky, kb * 16:(kb + 1) * 16]
                for r in range(R):
                    if y0 + r >= H:
                        continue
                    for xx in range(min(WT, W - x0)):
                        acc = np.zeros(Cout)
                        for kx in range(k):
                            acc += D[r, xx + kx, kx * Cout:(kx + 1) * Cout]
                        out[n, :, y0 + r, x0 + xx] = acc
    return out


def main():
    rng = np.random.default_rng(0)
    for (N, Cin, Cout, k, H, W, R, WT) in [(1, 64, 16, 11, 5, 130, 2, 96), (2, 3, 8, 7, 6, 50, 2, 32), (1, 16, 32, 3, 4, 100, 4, 96),
                                            (1, 32, 16, 5, 3, 96, 2, 96)]:
        x = rng.standard_normal((N, Cin, H, W))
        w = rng.standard_normal((Cout, Cin, k, k))
        ref = F.conv2d(torch.tensor(x), torch.tensor(w), padding=k // 2).numpy()
        got = conv_kx_fused(x, w, R, WT)
        err = np.abs(got - ref).max() / np.abs(ref).max()
        mmas = -(-H // R) * -(-W // WT) * N * R * k * ((Cin + 15) // 16)
        taps = -(-H // 16) * -(-W // 8) * N * k * k * ((Cin + 15) // 16)
        print(f"Cin {Cin} Cout {Cout} k{k} {H}x{W}: rel err {err:.2e}; MMAs kx-fused {mmas} (N={k * Cout}) vs per-tap {taps} (N={Cout})")
        assert err < 1e-12


if __name__ == "__main__":
    main()
