"""Write a synthetic clip in the reference's on-disk layout (loaders/video_dataset.py:83-88) so that
`DepthFineTuner.fine_tune()` / `save_depth()` can run end to end without COLMAP / FlowNet2 / ffmpeg:

    <root>/color_down/frame_%06d.raw      BGR-swizzled fp32 HWC (video.py:174 convention)
    <root>/flow/flow_%06d_%06d.raw        fp32 HWC (u, v) in pixels
    <root>/mask/mask_%06d_%06d.png        0 / 255
    <root>/flow_list.json                 [[i, j], ...]
    <range_dir>/metadata_scaled.npz       extrinsics (N,3,4), intrinsics (N,4)
"""
import json
import os

import numpy as np
import torch

from .synthetic import SyntheticVideo
from .utils import image_io


def _imwrite(path, arr):
    try:
        import cv2
        cv2.imwrite(path, arr)
    except ImportError:
        from PIL import Image
        Image.fromarray(arr).save(path)


def write_synthetic_dataset(root, range_dir, n_frames, H, W, device="cpu", seed=1234, pairs=None):
    video = SyntheticVideo(n_frames, H, W, torch.device(device), seed=seed, pairs=pairs)
    for d in ("color_down", "flow", "mask"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    os.makedirs(range_dir, exist_ok=True)
    for i in range(n_frames):
        im = video.frames[i].permute(1, 2, 0).cpu().numpy()              # HWC, BGR in memory
        image_io.save_raw_float32_image(os.path.join(root, "color_down", f"frame_{i:06d}.raw"), im[..., ::-1])
    for (r, t), fl in video.flow.items():
        image_io.save_raw_float32_image(os.path.join(root, "flow", f"flow_{r:06d}_{t:06d}.raw"), fl.permute(1, 2, 0).cpu().numpy())
        _imwrite(os.path.join(root, "mask", f"mask_{r:06d}_{t:06d}.png"), (video.mask[(r, t)][0].cpu().numpy() * 255).astype(np.uint8))
    json.dump([list(p) for p in video.pairs], open(os.path.join(root, "flow_list.json"), "w"))
    np.savez(os.path.join(range_dir, "metadata_scaled.npz"), extrinsics=video.extr.cpu().numpy(), intrinsics=video.intr.cpu().numpy())
    return video
