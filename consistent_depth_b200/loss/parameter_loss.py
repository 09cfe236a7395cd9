"""ParameterLoss mirror (loss/parameter_loss.py:6-19): lambda * sum_i |p_i - p_i^0| (L1)."""
import ctypes as C

import torch

from .. import _lib


class _ParamL1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lam, n_params, *tensors):
        params, inits = tensors[:n_params], tensors[n_params:]
        L = _lib.lib()
        dev = params[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        grads = []
        for p, p0 in zip(params, inits):
            g = torch.zeros_like(p)
            pc, p0c = p.detach().contiguous(), p0.contiguous()      # named: temporaries must outlive the launch's argument list
            _lib.check(L.cvd_param_l1(_lib.ptr(pc), _lib.ptr(p0c),
                                      C.c_longlong(p.numel()), C.c_float(lam), _lib.ptr(g), _lib.ptr(out),
                                      _lib.stream()), "cvd_param_l1")
            grads.append(g)
        ctx.grads = grads
        ctx.n = n_params
        return out.reshape(())

    @staticmethod
    def backward(ctx, g_out):
        return (None, None) + tuple(g * g_out for g in ctx.grads) + (None,) * ctx.n


class ParameterLoss(torch.nn.Module):
    def __init__(self, parameters_init, opt):
        super().__init__()
        self.parameters_init = [p.detach() for p in parameters_init]
        self.opt = opt
        assert opt.lambda_parameter > 0

    def __call__(self, parameters):
        params = list(parameters)
        loss = _ParamL1Fn.apply(float(self.opt.lambda_parameter), len(params), *params, *self.parameters_init)
        return loss, {"parameter_loss": loss.reshape(1, -1)}
