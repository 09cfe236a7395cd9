"""CPU restatement of the three FlowNet2 custom ops, forward only.

Follows third_party/flownet2/networks/
  correlation_package/correlation_cuda_kernel.cu:51-128 + correlation_cuda.cc:25-36   (explicit zero-padded copies, as there)
  resample2d_package/resample2d_kernel.cu:17-73                                        (kernel_size 1, bilinear)
  channelnorm_package/channelnorm_kernel.cu:16-60                                      (norm_deg 2)
PARITY UNPINNED: the reference implements these ops in CUDA only (built for sm_50..sm_70, setup.py), ships no tests or
fixtures for them, and there is no GPU in the build container, so this restatement follows the published kernel text and
cannot be checked against the reference's own output.  TEST INFRASTRUCTURE ONLY.
"""
import numpy as np


def correlation(in1, in2, pad, K, md, s1, s2):
    B, C, H, W = in1.shape
    kr, R = (K - 1) // 2, md // s2
    D, border = 2 * R + 1, (K - 1) // 2 + md
    Ho = -(-(H + 2 * pad - 2 * border) // s1)
    Wo = -(-(W + 2 * pad - 2 * border) // s1)
    P1 = np.pad(in1.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    P2 = np.pad(in2.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    out = np.zeros((B, D * D, Ho, Wo))
    ys, xs = np.arange(Ho) * s1 + md, np.arange(Wo) * s1 + md
    for tj in range(-R, R + 1):
        for ti in range(-R, R + 1):
            acc = np.zeros((B, Ho, Wo))
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    a = P1[:, :, (ys + j)[:, None], (xs + i)[None, :]]
                    b = P2[:, :, (ys + tj * s2 + j)[:, None], (xs + ti * s2 + i)[None, :]]
                    acc += (a * b).sum(1)
            out[:, (tj + R) * D + (ti + R)] = acc / (K * K * C)
    return out


def resample2d(in1, flow):
    B, C, H, W = in1.shape
    y, x = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    out = np.zeros_like(in1, dtype=np.float64)
    for n in range(B):
        xf, yf = x + flow[n, 0].astype(np.float32), y + flow[n, 1].astype(np.float32)
        fx, fy = np.floor(xf), np.floor(yf)
        al, be = (xf - fx).astype(np.float64), (yf - fy).astype(np.float64)
        xL, xR = np.clip(fx, 0, W - 1).astype(int), np.clip(fx + 1, 0, W - 1).astype(int)
        yT, yB = np.clip(fy, 0, H - 1).astype(int), np.clip(fy + 1, 0, H - 1).astype(int)
        p = in1[n].astype(np.float64)
        out[n] = ((1 - al) * (1 - be) * p[:, yT, xL] + al * (1 - be) * p[:, yT, xR]
                  + (1 - al) * be * p[:, yB, xL] + al * be * p[:, yB, xR])
    return out


def channelnorm(x):
    return np.sqrt((x.astype(np.float64) ** 2).sum(1, keepdims=True))
