"""GPU: fused flat Adam vs torch.optim.Adam golden (reference optimizer/__init__.py) and NaN guard."""
import os

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def test_fused_adam_matches_torch_adam_golden(golden_dir):
    from consistent_depth_b200 import optimizer
    g = np.load(os.path.join(golden_dir, "adam.npz"))
    dev = torch.device("cuda:0")
    p = torch.nn.Parameter(torch.tensor(synth.normal(31, 1, (1003,), 0.1), device=dev))
    opt = optimizer.create("Adam", [p], 4e-4, betas=(0.9, 0.999))
    for t in range(6):
        opt.zero_grad()
        p.grad.copy_(torch.tensor(synth.normal(31, 10 + t, (1003,), 10.0 ** (-t)), device=dev))
        opt.step()
        torch.cuda.synchronize()
        np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"p_{t}"], rtol=2e-6, atol=1e-9)


def test_multi_tensor_and_nan_skip():
    from consistent_depth_b200 import optimizer
    dev = torch.device("cuda:0")
    shapes = [(7,), (3, 5, 2), (1,), (130, 9)]
    ps = [torch.nn.Parameter(torch.tensor(synth.normal(5, i, s, 1.0), device=dev)) for i, s in enumerate(shapes)]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    opt = optimizer.create("Adam", ps, 1e-3, betas=(0.9, 0.999))
    ropt = torch.optim.Adam(ref, 1e-3, betas=(0.9, 0.999))
    flag = torch.zeros(1, device=dev)
    opt.loss_flag = flag
    for t in range(4):
        opt.zero_grad()
        for i, (p, r) in enumerate(zip(ps, ref)):
            gr = torch.tensor(synth.normal(6, 10 * t + i, p.shape, 0.3), device=dev)
            p.grad.copy_(gr); r.grad = gr.clone()
        if t == 2:                      # NaN loss: reference `continue`s before backward/step
            flag.fill_(float("nan"))
            before = [p.detach().clone() for p in ps]
            opt.step()
            for p, b in zip(ps, before):
                assert torch.equal(p.detach(), b)
            flag.zero_()
            continue
        opt.step(); ropt.step()
        for p, r in zip(ps, ref):
            torch.testing.assert_close(p.detach(), r.detach(), rtol=2e-6, atol=1e-9)
