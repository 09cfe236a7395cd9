"""JointLoss — sum of the enabled loss terms, with the call contract of the reference's loss/joint_loss.py:15-47:

    JointLoss(opt, parameters_init=None)(depths (B,2,H,W), metadata, parameters=None)
        -> (total loss of shape (1,), {name: per-sample losses})

The parameter term (lambda_parameter > 0) needs the initial parameters at construction and the current ones at call
time; the geometric-consistency term is on when either of its two weights is positive.
"""
from typing import List, Optional

import torch
from torch.nn import Parameter

from .consistency_loss import ConsistencyLoss
from .parameter_loss import ParameterLoss


class JointLoss(torch.nn.Module):
    def __init__(self, opt, parameters_init=None):
        super().__init__()
        self.opt = opt
        self.parameter_loss = self.consistency_loss = None
        if opt.lambda_parameter > 0:
            assert parameters_init is not None, "lambda_parameter > 0 needs the initial parameters"
            self.parameter_loss = ParameterLoss(parameters_init, opt)
        if max(opt.lambda_view_baseline, opt.lambda_reprojection) > 0:
            self.consistency_loss = ConsistencyLoss(opt)

    def __call__(self, depths, metadata, parameters: Optional[List[Parameter]] = None):
        terms = []
        if self.parameter_loss is not None:
            assert parameters is not None, "lambda_parameter > 0 needs the current parameters"
            terms.append(self.parameter_loss(parameters))
        if self.consistency_loss is not None:
            terms.append(self.consistency_loss(depths, metadata))
        total = torch.zeros(1, dtype=torch.float32, device=depths.device)
        per_sample = {}
        for value, parts in terms:
            total = total + value
            per_sample.update(parts)
        return total, per_sample
