"""GPU: the FlowNet2 custom ops on sm_100a against the oracle (their arithmetic is also checked on the host by
tests/test_flownet_ops_core_cpu.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import flownet_ops_oracle as fo
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_flownet2_custom_ops_match_oracle():
    from consistent_depth_b200.third_party.flownet2.networks import ChannelNorm, Correlation, Resample2d
    a, b = synth.uniform(1, 1, (2, 32, 12, 17), -1, 1), synth.uniform(1, 2, (2, 32, 12, 17), -1, 1)
    with torch.no_grad():
        out = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)(
            torch.tensor(a, device=DEV), torch.tensor(b, device=DEV))
    np.testing.assert_allclose(out.cpu().numpy(), fo.correlation(a, b, 20, 1, 20, 1, 2), rtol=1e-4, atol=1e-6)
    assert out.shape == (2, 441, 12, 17)
    flow = synth.normal(2, 2, (2, 2, 12, 17), 3.0)
    with torch.no_grad():
        r = Resample2d()(torch.tensor(a, device=DEV), torch.tensor(flow, device=DEV))
        n = ChannelNorm()(torch.tensor(a, device=DEV))
    np.testing.assert_allclose(r.cpu().numpy(), fo.resample2d(a, flow), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(n.cpu().numpy(), fo.channelnorm(a), rtol=1e-5)
