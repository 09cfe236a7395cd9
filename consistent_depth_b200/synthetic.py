"""Synthetic fine-tuning data in the reference's input contract (loaders/video_dataset.py:131-207):
images BGR in [0,1], flows in pixels, masks {0,1}, extrinsics [R|t] camera->world, intrinsics fx,fy,cx,cy.
Recipe of SURVEY.md §8(d): smooth rigid camera track, flow = reprojection of a smooth synthetic depth
+ N(0, 0.5 px) noise, masks Bernoulli(0.7).  Used by bench.py / smoke (no dataset or network here)."""
import math

import torch

from .loaders.frame_sampling import hierarchical2_one_way


def camera_track(n_frames, H, W, device):
    extr = torch.zeros(n_frames, 3, 4, dtype=torch.float64)
    for i in range(n_frames):
        ay, ax = 0.01 * i, 0.003 * i
        Ry = torch.tensor([[math.cos(ay), 0, math.sin(ay)], [0, 1, 0], [-math.sin(ay), 0, math.cos(ay)]], dtype=torch.float64)
        Rx = torch.tensor([[1, 0, 0], [0, math.cos(ax), -math.sin(ax)], [0, math.sin(ax), math.cos(ax)]], dtype=torch.float64)
        extr[i, :, :3] = Ry @ Rx
        extr[i, :, 3] = torch.tensor([0.05 * i, 0.01 * math.sin(0.3 * i), 0.0], dtype=torch.float64)
    intr = torch.tensor([0.8 * W, 0.8 * W, (W - 1) / 2.0, (H - 1) / 2.0], dtype=torch.float64).repeat(n_frames, 1)
    return extr.float().to(device), intr.float().to(device)


def scene_depth(i, H, W, device):
    y, x = torch.meshgrid(torch.arange(H, device=device, dtype=torch.float32),
                          torch.arange(W, device=device, dtype=torch.float32), indexing="ij")
    return 2.0 + 0.5 * torch.sin(x / W * 6.0 + 0.1 * i) * torch.cos(y / H * 4.0) + 0.3 * torch.sin((x + y) / (W + H) * 9.0)


def geometric_flow(depth, Ei, Ii, Ej, Ij):
    H, W = depth.shape
    dev = depth.device
    y, x = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32),
                          torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    rays = torch.stack([(x - Ii[2]) / Ii[0], -(y - Ii[3]) / Ii[1], -torch.ones_like(x)], 0)
    P = rays * depth[None]
    Pw = torch.einsum("ab,bhw->ahw", Ei[:, :3], P) + Ei[:, 3][:, None, None]
    Q = torch.einsum("ba,bhw->ahw", Ej[:, :3], Pw - Ej[:, 3][:, None, None])
    u = -Ij[0] * Q[0] / Q[2] + Ij[2]
    v = Ij[1] * Q[1] / Q[2] + Ij[3]
    return torch.stack([u - x, v - y], 0)


class SyntheticVideo:
    """All frames / flows / masks of a synthetic clip, resident on `device` (fits trivially in HBM)."""

    def __init__(self, n_frames, H, W, device, seed=1234, flow_noise=0.5, mask_p=0.7, pairs=None):
        g = torch.Generator(device=device).manual_seed(seed)
        self.n_frames, self.H, self.W, self.device = n_frames, H, W, device
        self.pairs = pairs if pairs is not None else hierarchical2_one_way(n_frames)
        self.extr, self.intr = camera_track(n_frames, H, W, device)
        self.frames = torch.rand(n_frames, 3, H, W, device=device, generator=g)
        depths = [scene_depth(i, H, W, device) for i in range(n_frames)]
        self.flow, self.mask = {}, {}
        for (i, j) in self.pairs:
            for (r, t) in ((i, j), (j, i)):
                fl = geometric_flow(depths[r], self.extr[r], self.intr[r], self.extr[t], self.intr[t])
                self.flow[(r, t)] = fl + flow_noise * torch.randn(2, H, W, device=device, generator=g)
                self.mask[(r, t)] = (torch.rand(1, H, W, device=device, generator=g) < mask_p).float()

    def batch(self, pair_ids):
        """Collated mini-batch dict (device tensors) for the given indices into self.pairs."""
        ps = [self.pairs[k] for k in pair_ids]
        return {
            "images": torch.stack([torch.stack([self.frames[i], self.frames[j]]) for i, j in ps]),
            "extrinsics": torch.stack([torch.stack([self.extr[i], self.extr[j]]) for i, j in ps]),
            "intrinsics": torch.stack([torch.stack([self.intr[i], self.intr[j]]) for i, j in ps]),
            "flows": [torch.stack([self.flow[(i, j)] for i, j in ps]), torch.stack([self.flow[(j, i)] for i, j in ps])],
            "masks": [torch.stack([self.mask[(i, j)] for i, j in ps]), torch.stack([self.mask[(j, i)] for i, j in ps])],
            "indices": torch.tensor(ps),
        }
