"""Model registry — same surface as the reference's monodepth/depth_model_registry.py:12-29.

"mc" is backed by the sm_100a engine.  "midas2" / "monodepth2" are named by the reference registry
but their backbones are not built in this round (SURVEY.md §8 rows a7/a8 are "next"): asking for
them raises a clear error instead of silently falling back to another implementation.
"""
from typing import List

from .depth_model import DepthModel
from .mannequin_challenge_model import MannequinChallengeModel


def get_depth_model_list() -> List[str]:
    return ["mc", "midas2", "monodepth2"]


def get_depth_model(type: str) -> DepthModel:
    if type == "mc":
        return MannequinChallengeModel
    elif type in ("midas2", "monodepth2"):
        raise NotImplementedError(
            f"model type '{type}' is registered by the reference but its sm_100a backbone is not built yet "
            "(consistent_depth_b200 has no PyTorch fallback)")
    else:
        raise ValueError(f"Unsupported model type '{type}'.")


def create_depth_model(type: str) -> DepthModel:
    model_class = get_depth_model(type)
    return model_class()
