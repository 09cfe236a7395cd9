"""Model registry — same surface as the reference's monodepth/depth_model_registry.py:12-29.

"mc" and "monodepth2" are backed by the sm_100a engine.  "midas2" is named by the reference registry but
its ResNeXt-101 backbone is not built yet: asking for it raises a clear error instead
of silently falling back to another implementation.
"""
from typing import List

from .depth_model import DepthModel
from .mannequin_challenge_model import MannequinChallengeModel
from .monodepth2_model import Monodepth2Model


def get_depth_model_list() -> List[str]:
    return ["mc", "midas2", "monodepth2"]


def get_depth_model(type: str) -> DepthModel:
    if type == "mc":
        return MannequinChallengeModel
    elif type == "monodepth2":
        return Monodepth2Model
    elif type == "midas2":
        raise NotImplementedError(
            f"model type '{type}' is registered by the reference but its sm_100a backbone is not built yet "
            "(consistent_depth_b200 has no PyTorch fallback)")
    else:
        raise ValueError(f"Unsupported model type '{type}'.")


def create_depth_model(type: str) -> DepthModel:
    model_class = get_depth_model(type)
    return model_class()
