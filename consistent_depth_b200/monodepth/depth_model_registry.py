"""Model registry — same surface as the reference's monodepth/depth_model_registry.py:12-29.

All three model types the reference registers ("mc", "midas2", "monodepth2") are backed by the sm_100a engine;
there is no PyTorch fallback behind any of them.
"""
from typing import List

from .depth_model import DepthModel
from .mannequin_challenge_model import MannequinChallengeModel
from .midas_v2_model import MidasV2Model
from .monodepth2_model import Monodepth2Model


def get_depth_model_list() -> List[str]:
    return ["mc", "midas2", "monodepth2"]


def get_depth_model(type: str) -> DepthModel:
    if type == "mc":
        return MannequinChallengeModel
    elif type == "monodepth2":
        return Monodepth2Model
    elif type == "midas2":
        return MidasV2Model
    else:
        raise ValueError(f"Unsupported model type '{type}'.")


def create_depth_model(type: str) -> DepthModel:
    model_class = get_depth_model(type)
    return model_class()
