"""VideoDataset / VideoFrameDataset — same files and sample layout as the reference's
loaders/video_dataset.py:80-242:

    color_down/frame_{id:06d}.raw (or .png)    flow/flow_{ref:06d}_{tgt:06d}.raw
    mask/mask_{ref:06d}_{tgt:06d}.png          <meta>.npz: extrinsics (N,3,4), intrinsics (N,4)
    flow_list.json: [[i, j], ...]

__getitem__ -> (images (2,3,H,W) BGR in [0,1], {"extrinsics" (2,3,4), "intrinsics" (2,4),
"geometry_consistency": {"indices" (2,), "flows" [2 x (2,H,W)], "masks" [2 x (1,H,W)]}}).
"""
import json
import os
from os.path import join as pjoin

import numpy as np
import torch
import torch.utils.data as data

from ..utils import image_io
from . import frame_sampling

_dtype = torch.float32


def _imread(path):
    try:
        import cv2
        return cv2.imread(path, cv2.IMREAD_UNCHANGED)
    except ImportError:
        from PIL import Image
        im = np.array(Image.open(path))
        return im[..., ::-1] if im.ndim == 3 else im


def load_image(path, channels_first, check_channels=None, post_proc_raw=lambda x: x, post_proc_other=lambda x: x):
    if os.path.splitext(path)[-1] == ".raw":
        im = post_proc_raw(image_io.load_raw_float32_image(path))
    else:
        im = post_proc_other(_imread(path))
    im = im.reshape(im.shape[:2] + (-1,))
    if check_channels is not None:
        assert im.shape[-1] == check_channels, f"receive image of shape {im.shape} whose #channels != {check_channels}"
    if channels_first:
        im = im.transpose((2, 0, 1))
    return torch.tensor(np.ascontiguousarray(im), dtype=_dtype)


def load_color(path, channels_first):
    return load_image(path, channels_first,
                      post_proc_raw=lambda im: im[..., [2, 1, 0]] if im.ndim == 3 else im,
                      post_proc_other=lambda im: im / 255)


def load_flow(path, channels_first):
    return load_image(path, channels_first, check_channels=2)


def load_mask(path, channels_first):
    return (load_image(path, channels_first, check_channels=1) > 0).to(_dtype)


class VideoDataset(data.Dataset):
    def __init__(self, path, meta_file=None):
        self.color_fmt = pjoin(path, "color_down", "frame_{:06d}.raw")
        if not os.path.isfile(self.color_fmt.format(0)):
            self.color_fmt = pjoin(path, "color_down", "frame_{:06d}.png")
        self.mask_fmt = pjoin(path, "mask", "mask_{:06d}_{:06d}.png")
        self.flow_fmt = pjoin(path, "flow", "flow_{:06d}_{:06d}.raw")
        if meta_file is not None:
            with open(meta_file, "rb") as f:
                meta = np.load(f)
                self.extrinsics = torch.tensor(meta["extrinsics"], dtype=_dtype)
                self.intrinsics = torch.tensor(meta["intrinsics"], dtype=_dtype)
            assert self.extrinsics.shape[0] == self.intrinsics.shape[0]
        flow_list_fn = pjoin(path, "flow_list.json")
        if os.path.isfile(flow_list_fn):
            with open(flow_list_fn, "r") as f:
                self.flow_indices = json.load(f)
        else:
            names = os.listdir(os.path.dirname(self.flow_fmt))
            self.flow_indices = [[int(s) for s in os.path.splitext(n)[0].split("_")[-2:]]
                                 for n in names if n.endswith(".raw")]
            self.flow_indices = frame_sampling.to_in_range(self.flow_indices)
        self.flow_indices = sorted(frame_sampling.to_one_way(tuple(p) for p in self.flow_indices))

    def __getitem__(self, index):
        pair = self.flow_indices[index]
        images = torch.stack([load_color(self.color_fmt.format(k), channels_first=True) for k in pair], dim=0)
        flows = [load_flow(self.flow_fmt.format(a, b), channels_first=True) for a, b in (pair, pair[::-1])]
        masks = [load_mask(self.mask_fmt.format(a, b), channels_first=True) for a, b in (pair, pair[::-1])]
        metadata = {
            "extrinsics": torch.stack([self.extrinsics[k] for k in pair], dim=0),
            "intrinsics": torch.stack([self.intrinsics[k] for k in pair], dim=0),
            "geometry_consistency": {"indices": torch.tensor(pair), "flows": flows, "masks": masks},
        }
        # optional per-frame / global depth scales (video_dataset.py:196-204): an attribute a caller may set on the dataset;
        # DepthModel.forward multiplies the prediction by metadata["scales"] when present (depth_model.py:24-28)
        scales = getattr(self, "scales", None)
        if scales:
            if isinstance(scales, dict):
                metadata["scales"] = torch.stack([torch.Tensor([scales[k]]) for k in pair], dim=0)
            else:
                metadata["scales"] = torch.Tensor([scales, scales]).reshape(2, 1)
        return images, metadata

    def __len__(self):
        return len(self.flow_indices)


class VideoFrameDataset(data.Dataset):
    def __init__(self, color_fmt, frames=None):
        self.color_fmt = color_fmt
        if frames is None:
            frames = range(len(os.listdir(os.path.dirname(self.color_fmt))))
        self.frames = frames

    def __getitem__(self, index):
        frame_id = self.frames[index]
        return load_color(self.color_fmt.format(frame_id), channels_first=True), {"frame_id": frame_id}

    def __len__(self):
        return len(self.frames)
