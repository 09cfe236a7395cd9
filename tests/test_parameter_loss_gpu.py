"""GPU: the lambda_parameter regulariser (cvd_param_l1) against the oracle and the reference-generated golden
(tests/golden/parameter_loss.npz: loss/parameter_loss.py:13-19 through loss/joint_loss.py:34-39), through the
ParameterLoss / JointLoss mirrors and inside the fused FineTuneStep.  Tolerances: loss rel 1e-5 (fp32 sums of ~1e3
terms), gradients exact (+-lambda or 0)."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import synth, parameter_oracle as po, hourglass_oracle as ho

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_parameter_loss_module_matches_reference_golden(golden_dir):
    from consistent_depth_b200.loss.joint_loss import JointLoss
    g = np.load(os.path.join(golden_dir, "parameter_loss.npz"))
    lam = float(g["lambda"])
    inits, params = po.make_case(int(g["seed"]))
    p_init = [torch.tensor(a, device=DEV) for a in inits]
    ps = [torch.nn.Parameter(torch.tensor(a, device=DEV)) for a in params]
    opt = types.SimpleNamespace(lambda_view_baseline=0.0, lambda_reprojection=0.0, lambda_parameter=lam)
    depths = torch.zeros(1, 2, 4, 4, device=DEV)
    loss, meta = JointLoss(opt, p_init)(depths, None, parameters=ps)
    assert loss.shape == (1,) and meta["parameter_loss"].shape == (1, 1)
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(meta["parameter_loss"].detach().cpu().numpy(), g["parameter_loss"], rtol=1e-5)
    loss.backward()
    ref_loss, ref_grads = po.parameter_loss_and_grad(params, inits, lam)
    for i, p in enumerate(ps):
        np.testing.assert_array_equal(p.grad.cpu().numpy(), g[f"grad_{i}"])
        np.testing.assert_array_equal(p.grad.cpu().numpy(), ref_grads[i])
    # together with the consistency term: the total the fine-tuning loop sees (joint_loss.py:32-46)
    from oracle.make_golden import CONSISTENCY_CASES  # noqa: F401  (constants only; no reference import)
    cseed, pairs, H, W, stress, lr_, lb_ = CONSISTENCY_CASES["geo_b2"]
    batch = synth.make_pair_batch(cseed, pairs, H, W, stress=stress)
    t = lambda a: torch.tensor(a, device=DEV)
    md = {"extrinsics": t(batch["extrinsics"]), "intrinsics": t(batch["intrinsics"]),
          "geometry_consistency": {"indices": t(batch["indices"]), "flows": [t(f) for f in batch["flows"]],
                                   "masks": [t(m) for m in batch["masks"]]}}
    opt2 = types.SimpleNamespace(lambda_view_baseline=lb_, lambda_reprojection=lr_, lambda_parameter=lam)
    total, meta2 = JointLoss(opt2, p_init)(t(synth.synth_depth_pred(cseed, len(pairs), H, W)), md, parameters=ps)
    np.testing.assert_allclose(total.detach().cpu().numpy(), g["total_with_geo_b2"], rtol=1e-5)
    assert sorted(meta2.keys()) == list(g["meta_keys"])


def test_fused_step_applies_parameter_loss():
    """FineTuneStep(lambda_parameter > 0): loss and flat gradient gain lambda*sum|p - p0| and lambda*sign(p - p0) exactly
    (compared with the same step at lambda_parameter = 0 on identical weights; depth_fine_tuning.py:223-224,272)."""
    from consistent_depth_b200.fine_tune_step import FineTuneStep
    from consistent_depth_b200.monodepth.mannequin_challenge_model import MannequinChallengeModel
    seed, H, W, lam = 41, 32, 48, 0.05
    sd = {k: torch.tensor(np.asarray(v)) for k, v in ho.mc_init_state(seed).items()}
    b = synth.make_pair_batch(seed, [(0, 2)], H, W)
    t = lambda a: torch.tensor(a, device=DEV)
    out = {}
    for name, lp in (("off", 0.0), ("on", lam)):
        model = MannequinChallengeModel(state_dict=sd)
        p0 = model.P.flat.clone()
        g = torch.Generator(device="cpu").manual_seed(5)
        model.P.flat.add_((torch.rand(p0.shape, generator=g) - 0.5).to(DEV) * 2e-3)   # weights already moved away from p0
        step = FineTuneStep(model, 1, H, W, lr=4e-4, use_graph=False, lambda_parameter=lp, parameters_init=p0)
        step.load_batch(t(b["images"]), [t(f) for f in b["flows"]], [t(m) for m in b["masks"]], t(b["extrinsics"]), t(b["intrinsics"]))
        w = model.P.flat.clone()
        step._fwd_bwd()
        torch.cuda.synchronize()
        out[name] = (float(step.loss), model.P.grad_flat.clone(), w, p0, step.loss_meta())
    l0, g0, w, p0, _ = out["off"]
    l1, g1, _, _, meta = out["on"]
    d = (w - p0).double()
    exp_loss = lam * float(d.abs().sum())
    assert abs((l1 - l0) - exp_loss) <= 1e-4 * exp_loss + 1e-5 * abs(l0), (l1 - l0, exp_loss)
    np.testing.assert_allclose(float(meta["parameter_loss"]), exp_loss, rtol=2e-3)
    # gradients: conv wgrad REDs are order-dependent in the last ulps, so compare the difference with a small tolerance
    diff = (g1 - g0).cpu().numpy()
    np.testing.assert_allclose(diff, (lam * torch.sign(d)).float().cpu().numpy(), atol=2e-4 * float(g0.abs().max()) + 1e-7)
