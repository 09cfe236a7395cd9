"""CPU restatement of the flow / photometric consistency masks.

Follows utils/consistency.py:8-67:
  sample (:8-24)            grid = 2 uv / (W, H) - 1 and grid_sample(padding_mode="border", align_corners=False)
                            => source coordinates (u - 0.5, v - 0.5) clamped to the image, bilinear
  consistency_mask (:32-50) inside-image test on (x + u, y + v) AND sum of squared differences < threshold
  consistent_flow_masks (:53-67)   flow mask (flow_ref vs sampled -flow_tgt, thresh^2) AND photo mask (colours, C thresh^2)
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/flowmask.npz.
"""
import numpy as np


def sample(data, uv):
    """data (H, W, C), uv (H, W, 2) in pixels -> (H, W, C): bilinear at (u - 0.5, v - 0.5), border-clamped."""
    H, W = data.shape[:2]
    sx = np.clip(uv[..., 0].astype(np.float32) - np.float32(0.5), 0, W - 1)
    sy = np.clip(uv[..., 1].astype(np.float32) - np.float32(0.5), 0, H - 1)
    x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
    x1, y1 = np.minimum(x0 + 1, W - 1), np.minimum(y0 + 1, H - 1)
    tx, ty = (sx - x0)[..., None], (sy - y0)[..., None]
    d = data.reshape(H, W, -1).astype(np.float32)
    out = (d[y0, x0] * (1 - tx) * (1 - ty) + d[y0, x1] * tx * (1 - ty) + d[y1, x0] * (1 - tx) * ty + d[y1, x1] * tx * ty)
    return out.reshape(data.shape)


def consistency_mask(im_ref, im_tgt, flow, threshold):
    H, W = im_ref.shape[:2]
    X, Y = np.meshgrid(np.arange(W), np.arange(H))
    idx_x, idx_y = flow[..., 0] + X, flow[..., 1] + Y
    inside = (idx_x >= 0) & (idx_x <= W - 1) & (idx_y >= 0) & (idx_y <= H - 1)
    warped = sample(im_tgt.reshape(H, W, -1), np.stack((idx_x, idx_y), -1))
    d = im_ref.reshape(H, W, -1) - warped
    return inside & (np.sum(d * d, -1) < threshold)


def consistent_flow_masks(flows, colors, flow_thresh, color_thresh):
    mf = [consistency_mask(fr, -ft, fr, flow_thresh ** 2) for fr, ft in zip(flows, flows[::-1])]
    C = colors[0].shape[-1]
    mp = [consistency_mask(cr, ct, fr, C * color_thresh ** 2) for cr, ct, fr in zip(colors, colors[::-1], flows)]
    return [a & b for a, b in zip(mf, mp)]


def synthetic_pair(seed, H, W):
    """Two colour frames and the two flows between them: a smooth warp + noise so that both mask tests cut."""
    from . import synth
    y, x = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    u = 2.5 * np.sin(y / H * 3.0) + 1.0
    v = 1.5 * np.cos(x / W * 4.0)
    f01 = np.stack([u, v], -1).astype(np.float32) + synth.normal(seed, 1, (H, W, 2), 0.4)
    f10 = -np.stack([u, v], -1).astype(np.float32) + synth.normal(seed, 2, (H, W, 2), 0.4)
    base = 0.5 + 0.4 * np.stack([np.sin(x / 5.0 + y / 7.0), np.cos(x / 6.0), np.sin(y / 4.0)], -1).astype(np.float32)
    c0 = (base + synth.normal(seed, 3, (H, W, 3), 0.02)).astype(np.float32)
    c1 = (sample(base, np.stack([x - u + 0.5, y - v + 0.5], -1)) + synth.normal(seed, 4, (H, W, 3), 0.3)).astype(np.float32)
    return [f01, f10], [c0, c1]
