"""CPU restatement of the lambda_parameter regulariser.  TEST INFRASTRUCTURE ONLY.

Follows loss/parameter_loss.py:13-19: `sq_diff` is |p - p_init| (L1 despite its name), summed over ALL parameter
tensors and scaled by lambda_parameter; joint_loss.py:34-39 adds it to the (1,)-shaped total and reports it as
`parameter_loss` of shape (1, 1).  Autograd gives d loss / d p = lambda * sign(p - p_init) (torch.abs has
subgradient 0 at 0).  Pinned by tests/golden/parameter_loss.npz (oracle/make_golden.py --only parameter).
"""
import numpy as np

from . import synth


def make_case(seed, shapes=((16, 3, 3, 3), (16,), (8, 16, 1, 1), (5,))):
    """Deterministic (p_init, p) lists; a few entries of p equal p_init exactly (the |.| kink)."""
    inits = [synth.uniform(seed, 10 + i, s, -0.5, 0.5) for i, s in enumerate(shapes)]
    params = [a + synth.uniform(seed, 50 + i, s, -0.01, 0.01) for i, (a, s) in enumerate(zip(inits, shapes))]
    for a, p in zip(inits, params):
        p.reshape(-1)[::7] = a.reshape(-1)[::7]
    return inits, params


def parameter_loss_and_grad(params, inits, lam):
    """(loss scalar, [d loss / d p_i]) in float64 accumulation, fp32 outputs."""
    total = 0.0
    grads = []
    for p, p0 in zip(params, inits):
        d = p.astype(np.float64) - p0.astype(np.float64)
        total += np.abs(d).sum()
        grads.append((lam * np.sign(d)).astype(np.float32))
    return np.float32(lam * total), grads
