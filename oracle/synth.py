"""Deterministic synthetic inputs shared by the oracle, the tests and bench.py.

All randomness comes from a counter-based splitmix64 hash so that numpy on the
CPU box and numpy on the GPU box (or any other language) produce bit-identical
inputs from (seed, stream, index) -- no dependence on torch/numpy RNG versions.
Follows the input contract of loaders/video_dataset.py:131-207 (images BGR in
[0,1], flows in pixels, masks {0,1}, extrinsics [R|t] camera->world, intrinsics
fx,fy,cx,cy) and SURVEY.md §8(d)'s synthetic-data recipe.
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform(seed, stream, shape, lo=0.0, hi=1.0, dtype=np.float32):
    """U[lo,hi) array of `shape`, a pure function of (seed, stream)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        base = _splitmix64(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        idx = np.arange(n, dtype=np.uint64)
        h = _splitmix64(base + idx * np.uint64(0x9E3779B97F4A7C15))
    u = (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return (lo + (hi - lo) * u).reshape(shape).astype(dtype)


def normal(seed, stream, shape, std=1.0, dtype=np.float32):
    u1 = uniform(seed, stream * 2 + 1, shape, 1e-12, 1.0, np.float64)
    u2 = uniform(seed, stream * 2 + 2, shape, 0.0, 1.0, np.float64)
    return (std * np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)).astype(dtype)


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]], dtype=np.float64)


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], dtype=np.float64)


def camera_track(n_frames, H, W):
    """Smooth rigid track: R = Ry(.01 i) Rx(.003 i), t = (.05 i, .01 sin(.3 i), 0)."""
    extr = np.zeros((n_frames, 3, 4), dtype=np.float64)
    for i in range(n_frames):
        extr[i, :, :3] = _rot_y(0.01 * i) @ _rot_x(0.003 * i)
        extr[i, :, 3] = (0.05 * i, 0.01 * np.sin(0.3 * i), 0.0)
    intr = np.tile(np.array([0.8 * W, 0.8 * W, (W - 1) / 2.0, (H - 1) / 2.0], dtype=np.float64), (n_frames, 1))
    return extr.astype(np.float32), intr.astype(np.float32)


def scene_depth(i, H, W):
    """Smooth synthetic depth of frame i: 2 + sin-bumps (stays well away from 0)."""
    y, x = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    return 2.0 + 0.5 * np.sin(x / W * 6.0 + 0.1 * i) * np.cos(y / H * 4.0) + 0.3 * np.sin((x + y) / (W + H) * 9.0)


def geometric_flow(depth_i, extr_i, intr_i, extr_j, intr_j):
    """Flow i->j implied by depth_i and the poses (same maths as utils/geometry.py:38-128)."""
    H, W = depth_i.shape
    y, x = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    fx, fy, cx, cy = [float(v) for v in intr_i]
    rays = np.stack([(x - cx) / fx, -(y - cy) / fy, -np.ones_like(x)], 0)
    P = rays * depth_i[None]
    Ri, ti = extr_i[:, :3].astype(np.float64), extr_i[:, 3].astype(np.float64)
    Rj, tj = extr_j[:, :3].astype(np.float64), extr_j[:, 3].astype(np.float64)
    Pw = np.einsum("ab,bhw->ahw", Ri, P) + ti[:, None, None]
    Q = np.einsum("ba,bhw->ahw", Rj, Pw - tj[:, None, None])
    fxj, fyj, cxj, cyj = [float(v) for v in intr_j]
    u = -fxj * Q[0] / Q[2] + cxj
    v = fyj * Q[1] / Q[2] + cyj
    return np.stack([u - x, v - y], 0)


def make_pair_batch(seed, pairs, H, W, n_frames=None, flow_noise=0.5, mask_p=0.7, stress=False):
    """Batch dict in the reference's collated layout for `pairs` = [(i,j),...].

    Returns numpy arrays: images (B,2,3,H,W), extrinsics (B,2,3,4), intrinsics (B,2,4),
    flows [2 x (B,2,H,W)], masks [2 x (B,1,H,W)], indices (B,2) int64.
    """
    B = len(pairs)
    n_frames = n_frames or (max(max(p) for p in pairs) + 1)
    extr, intr = camera_track(n_frames, H, W)
    images = np.zeros((B, 2, 3, H, W), np.float32)
    flows = [np.zeros((B, 2, H, W), np.float32) for _ in range(2)]
    masks = [np.zeros((B, 1, H, W), np.float32) for _ in range(2)]
    for b, (i, j) in enumerate(pairs):
        for k, (r, t) in enumerate(((i, j), (j, i))):
            images[b, k] = uniform(seed, 1000 + r, (3, H, W))
            if stress:
                fl = normal(seed, 5000 + r * 4096 + t, (2, H, W), 2.0, np.float64)
            else:
                fl = geometric_flow(scene_depth(r, H, W), extr[r], intr[r], extr[t], intr[t])
                fl = fl + normal(seed, 5000 + r * 4096 + t, (2, H, W), flow_noise, np.float64)
            flows[k][b] = fl.astype(np.float32)
            masks[k][b, 0] = (uniform(seed, 9000000 + r * 4096 + t, (H, W)) < mask_p).astype(np.float32)
    ext_b = np.stack([np.stack([extr[i], extr[j]]) for i, j in pairs])
    int_b = np.stack([np.stack([intr[i], intr[j]]) for i, j in pairs])
    return {
        "images": images, "extrinsics": ext_b, "intrinsics": int_b,
        "flows": flows, "masks": masks, "indices": np.array(pairs, dtype=np.int64),
    }


def synth_depth_pred(seed, B, H, W):
    """A plausible network output: scene depth with a smooth multiplicative error."""
    d = np.zeros((B, 2, H, W), np.float32)
    for b in range(B):
        for k in range(2):
            base = scene_depth(b * 2 + k, H, W)
            d[b, k] = (base * (1.0 + 0.2 * (uniform(seed, 777 + b * 2 + k, (H, W), dtype=np.float64) - 0.5))).astype(np.float32)
    return d
