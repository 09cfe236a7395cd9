"""DepthModel plugin base class — same contract as the reference's monodepth/depth_model.py:8-38.

forward(images[, metadata]) -> depth: images (...,3,H,W) BGR in [0,1], any leading dims; depth (...,H,W),
fp32, on the GPU, autograd-connected to parameters().  metadata["scales"] (optional) post-scales depth.
"""
from abc import abstractmethod

import torch


class DepthModel(torch.nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, images, metadata=None):
        depth = self.estimate_depth(images)
        if metadata is not None:
            if "scales" in metadata:
                factor = metadata["scales"].unsqueeze(3).cuda()
                depth = depth * factor
        return depth

    @abstractmethod
    def estimate_depth(self, images, metadata=None) -> torch.Tensor:
        pass

    @abstractmethod
    def save(self, label):
        pass
