"""CPU: the pure-Python plugin surface that mirrors the reference (registry lookup, DepthModel.forward's optional
"scales", JointLoss term plumbing, to_device's in-place container contract) -- no kernels involved."""
import types

import pytest
import torch


def test_registry_lookup_and_errors():
    from consistent_depth_b200.monodepth.depth_model_registry import get_depth_model, get_depth_model_list
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    from consistent_depth_b200.monodepth.monodepth2_model import Monodepth2Model
    assert get_depth_model_list() == ["mc", "midas2", "monodepth2"]                       # depth_model_registry.py:12-13
    assert [get_depth_model(t) for t in get_depth_model_list()] == [MannequinChallengeModel, MidasV2Model, Monodepth2Model]
    with pytest.raises(ValueError, match="Unsupported model type 'nope'"):                # :23-24
        get_depth_model("nope")
    # params.py:110-119 reads these from the CLASS
    assert [(m.align, m.learning_rate, m.lambda_view_baseline) for m in (MannequinChallengeModel, MidasV2Model, Monodepth2Model)] \
        == [(16, 0.0004, 0.1), (32, 0.0001, 0.0001), (1, 0.00004, 1)]


def test_depth_model_forward_applies_optional_scales():
    from consistent_depth_b200.monodepth.depth_model import DepthModel

    class Two(DepthModel):
        def estimate_depth(self, images, metadata=None):
            return images[..., 0, :, :] * 0 + 2

        def save(self, label):
            pass
    m, x = Two(), torch.rand(2, 2, 3, 4, 5)
    assert m(x).shape == (2, 2, 4, 5) and torch.equal(m(x, {"unrelated": 1}), m(x))
    scaled = m(x, {"scales": torch.tensor([[1.0, 2.0], [3.0, 4.0]])[..., None]})             # depth_model.py:22-28
    assert torch.equal(scaled[1, 1], torch.full((4, 5), 8.0)) and torch.equal(scaled[0, 0], torch.full((4, 5), 2.0))


def test_joint_loss_term_plumbing(monkeypatch):
    import consistent_depth_b200.loss.joint_loss as jl
    off = types.SimpleNamespace(lambda_parameter=0, lambda_view_baseline=0, lambda_reprojection=0)
    loss, parts = jl.JointLoss(off)(torch.zeros(1, 2, 4, 5), {})
    assert loss.shape == (1,) and float(loss) == 0 and parts == {}

    class Stub:
        def __init__(self, *a):
            pass

        def __call__(self, *a):
            return torch.full((1,), 3.0), {"stub": torch.full((1, 1), 3.0)}
    monkeypatch.setattr(jl, "ParameterLoss", Stub)
    monkeypatch.setattr(jl, "ConsistencyLoss", Stub)
    both = types.SimpleNamespace(lambda_parameter=1, lambda_view_baseline=0.1, lambda_reprojection=1.0)
    loss, parts = jl.JointLoss(both, [torch.zeros(1)])(torch.zeros(1, 2, 4, 5), {}, [torch.zeros(1)])
    assert float(loss) == 6.0 and list(parts) == ["stub"]
    with pytest.raises(AssertionError):
        jl.JointLoss(both)                                   # parameter term without the initial parameters (joint_loss.py:20-21)


def test_to_device_updates_containers_in_place():
    from consistent_depth_b200.utils.torch_helpers import _device, to_device
    d = {"a": torch.zeros(2), "b": [torch.ones(1), 3, "s", None], "c": {"d": torch.zeros(1)}}
    inner_list, inner_dict = d["b"], d["c"]
    r = to_device(d)
    assert r is d and d["b"] is inner_list and d["c"] is inner_dict and d["b"][1:] == [3, "s", None]
    assert d["a"].device.type == _device.type and d["c"]["d"].device.type == _device.type
