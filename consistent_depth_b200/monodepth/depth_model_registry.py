"""Model registry — the lookup surface of the reference's monodepth/depth_model_registry.py:12-29
(`get_depth_model_list`, `get_depth_model`, `create_depth_model`; unknown type -> ValueError).

All three model types the reference registers are backed by the sm_100a engines; none has a PyTorch fallback.
"""
from typing import Dict, List, Type

from .depth_model import DepthModel
from .mannequin_challenge_model import MannequinChallengeModel
from .midas_v2_model import MidasV2Model
from .monodepth2_model import Monodepth2Model

# listing order = the reference's get_depth_model_list()
_REGISTRY: Dict[str, Type[DepthModel]] = {
    "mc": MannequinChallengeModel,
    "midas2": MidasV2Model,
    "monodepth2": Monodepth2Model,
}


def get_depth_model_list() -> List[str]:
    return list(_REGISTRY)


def get_depth_model(type: str) -> Type[DepthModel]:
    """The model CLASS (params.py:110-119 reads its class attributes before anything is instantiated)."""
    try:
        return _REGISTRY[type]
    except KeyError:
        raise ValueError(f"Unsupported model type '{type}'.") from None


def create_depth_model(type: str) -> DepthModel:
    return get_depth_model(type)()
