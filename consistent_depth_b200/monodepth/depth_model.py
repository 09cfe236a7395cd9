"""DepthModel — the plugin base class every depth backbone derives from (contract of the reference's
monodepth/depth_model.py:8-38).

    forward(images[, metadata]) -> depth
        images   (..., 3, H, W) BGR in [0, 1], any leading dimensions
        depth    (..., H, W) fp32 on the GPU, autograd-connected to parameters()
        metadata optional dict; if it carries "scales" (one factor per frame) the depth is multiplied by them
    estimate_depth(images) and save(label) are what a backbone implements.
"""
from abc import abstractmethod

import torch


class DepthModel(torch.nn.Module):
    def forward(self, images, metadata=None):
        depth = self.estimate_depth(images)
        scales = None if metadata is None else metadata.get("scales")
        if scales is None:
            return depth
        return depth * scales.unsqueeze(3).to(depth.device)       # (B, 2, 1) -> broadcast over (B, 2, H, W)

    @abstractmethod
    def estimate_depth(self, images, metadata=None) -> torch.Tensor:
        ...

    @abstractmethod
    def save(self, label):
        ...
