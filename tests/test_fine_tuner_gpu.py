"""GPU: the DepthFineTuner drop-in end to end on a tiny synthetic clip written in the reference's on-disk layout:
dataset files -> VideoDataset -> fine_tune() (validation passes, fused CUDA-graph steps, checkpoints) -> save_depth().
Checks the boundary contract of depth_fine_tuning.py:139-406 (attributes, files written, checkpoint keys)."""
import json
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def make_params(root, **kw):
    p = types.SimpleNamespace(path=root, model_type="mc", batch_size=2, learning_rate=0, optimizer="Adam", num_epochs=2,
                              lambda_view_baseline=-1, lambda_reprojection=1.0, lambda_parameter=0, val_epoch_freq=1,
                              print_freq=1, display_freq=100, save_epoch_freq=1, log_dir=None)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


@pytest.mark.parametrize("resident", [True, False])
def test_fine_tune_and_save_depth_end_to_end(tmp_path, resident):
    from consistent_depth_b200.depth_fine_tuning import DepthFineTuner, make_tag
    from consistent_depth_b200.monodepth import mc_arch
    from consistent_depth_b200.synthetic_dataset import write_synthetic_dataset
    from consistent_depth_b200.utils import image_io
    root, range_dir = str(tmp_path / "clip"), str(tmp_path / "clip" / "R0-4_hierarchical2_mc")
    H, W, n = 32, 48, 4
    write_synthetic_dataset(root, range_dir, n, H, W, pairs=[(0, 1), (1, 2), (2, 3), (0, 2)])
    params = make_params(root, resident_dataset=resident)     # HBM-resident clip (default) or the reference's file DataLoader
    ft = DepthFineTuner(range_dir, list(range(n)), params)
    # params.py:110-119 semantics: sentinels resolved from the model class
    assert params.learning_rate == 0.0004 and params.lambda_view_baseline == 0.1
    assert ft.out_dir == os.path.join(range_dir, make_tag(params)) and os.path.isdir(os.path.join(ft.out_dir, "checkpoints"))
    w0 = ft.model.P.flat.clone()
    ft.save_depth(os.path.join(range_dir, "depth_mc"), list(range(n)))            # initial depth (eval mode)
    d0 = image_io.load_raw_float32_image(os.path.join(range_dir, "depth_mc", "depth", "frame_000002.raw"))
    assert d0.shape == (H, W) and np.isfinite(d0).all() and (d0 > 0).all()
    ft.fine_tune(writer=None)
    assert (ft.model.P.flat - w0).abs().max().item() > 0                            # weights moved
    for e in (1, 2):
        ck = torch.load(os.path.join(ft.out_dir, "checkpoints", f"{e:04d}.pth"), map_location="cpu")
        # the reference saves netG.state_dict() of the DataParallel-wrapped hourglass: same keys / order, `module.` prefix
        assert list(ck.keys()) == ["module." + k for k in mc_arch.state_dict_shapes().keys()]
    # round trip in both formats: a reference-format checkpoint and a bare HourglassModel state_dict load identically
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    m2 = MannequinChallengeModel(state_dict=ck)
    assert torch.equal(m2.P.flat, ft.model.P.flat) and torch.equal(m2.P.buf_flat, ft.model.P.buf_flat)
    m2.load_state_dict({k[7:]: v for k, v in ck.items()})
    assert torch.equal(m2.P.flat, ft.model.P.flat)
    ev = os.path.join(ft.out_dir, "eval")
    losses = json.load(open(os.path.join(ev, "loss_e0002_iter000008.json")))
    assert set(losses) == {"reprojection", "disparity", "mean"} and len(losses["reprojection"]) == 4
    assert os.path.isfile(os.path.join(ev, "depth_000003_e0000_iter000000.raw"))
    ft.save_depth()                                                                   # final depth export
    d1 = image_io.load_raw_float32_image(os.path.join(ft.out_dir, "depth", "frame_000002.raw"))
    assert d1.shape == (H, W) and np.isfinite(d1).all()
    # BN running statistics were updated by training AND by the train-mode validation passes
    assert ft.model.P.num_batches_tracked > 8


def test_registry_surface():
    from consistent_depth_b200.monodepth.depth_model_registry import get_depth_model, get_depth_model_list
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    assert get_depth_model_list() == ["mc", "midas2", "monodepth2"]
    assert get_depth_model("mc") is MannequinChallengeModel
    assert (MannequinChallengeModel.align, MannequinChallengeModel.learning_rate, MannequinChallengeModel.lambda_view_baseline) == (16, 0.0004, 0.1)
    with pytest.raises(ValueError):
        get_depth_model("nope")
    from consistent_depth_b200.monodepth.midas_v2_model import MidasV2Model
    assert get_depth_model("midas2") is MidasV2Model
    assert (MidasV2Model.align, MidasV2Model.learning_rate, MidasV2Model.lambda_view_baseline) == (32, 0.0001, 0.0001)


@pytest.mark.parametrize("model_type,H,W,lr,lam_b", [("monodepth2", 32, 48, 0.00004, 1), ("midas2", 64, 96, 0.0001, 0.0001)])
def test_fine_tune_end_to_end_other_backbones(tmp_path, model_type, H, W, lr, lam_b):
    """The same DepthFineTuner run with the other two registered model types (params.py:110-119 defaults, validation
    passes, CUDA-graph steps, checkpoints where the reference writes them, depth export)."""
    from consistent_depth_b200.depth_fine_tuning import DepthFineTuner
    from consistent_depth_b200.synthetic_dataset import write_synthetic_dataset
    from consistent_depth_b200.utils import image_io
    root, range_dir = str(tmp_path / "clip"), str(tmp_path / "clip" / f"R0-4_hierarchical2_{model_type}")
    n = 4
    write_synthetic_dataset(root, range_dir, n, H, W, pairs=[(0, 1), (1, 2), (2, 3), (0, 2)])
    params = make_params(root, model_type=model_type, num_epochs=1)
    ft = DepthFineTuner(range_dir, list(range(n)), params)
    assert params.learning_rate == lr and params.lambda_view_baseline == lam_b
    w0 = ft.model.P.flat.clone()
    ft.fine_tune(writer=None)
    assert (ft.model.P.flat - w0).abs().max().item() > 0
    ck = os.path.join(ft.out_dir, "checkpoints", "0001.pth")
    if model_type == "monodepth2":
        assert not os.path.exists(ck)                              # Monodepth2Model.save is a no-op (monodepth2_model.py:92-93)
    else:
        assert len(torch.load(ck, map_location="cpu")) == 666      # MidasNet.state_dict()
    losses = json.load(open(os.path.join(ft.out_dir, "eval", "loss_e0001_iter000004.json")))
    assert len(losses["reprojection"]) == 4 and all(np.isfinite(v) for v in losses["mean"].values())
    ft.save_depth()
    d1 = image_io.load_raw_float32_image(os.path.join(ft.out_dir, "depth", "frame_000001.raw"))
    assert d1.shape == (H, W) and np.isfinite(d1).all() and (d1 > 0).all()
