"""Optimizer registry mirror (optimizer/__init__.py:5-17): name -> class, `create(name, params, lr, betas=)`.

"Adam" maps to FusedAdam: torch.optim.Adam semantics (eps=1e-8, no weight decay /
amsgrad) as ONE flat-buffer CUDA kernel (`cvd_adam_flat`).  Parameters are
re-homed into a single contiguous fp32 buffer (and their .grad into a matching
one) so that step() is a single launch and a multi-GPU gradient all-reduce is a
single NCCL call on `grad_flat`.
"""
import ctypes as C

import torch

from .. import _lib


class FlatParamList(list):
    """A parameter list whose tensors are already views of one flat buffer (`.flat`, with `.grad_flat`)."""
    flat = None
    grad_flat = None


class FusedAdam:
    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8):
        adopt = params if isinstance(params, FlatParamList) and params.flat is not None else None
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise _lib.CvdError("FusedAdam needs CUDA parameters (no CPU path)")
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.grad_scale = 1.0
        self.loss_flag = None       # optional device scalar; NaN => step skipped on device
        if adopt is not None:       # engine-owned flat storage: no re-homing
            self.flat, self.grad_flat, self.numel = adopt.flat, adopt.grad_flat, adopt.flat.numel()
            self.exp_avg = torch.zeros_like(self.flat)
            self.exp_avg_sq = torch.zeros_like(self.flat)
            self.state = torch.zeros(4, dtype=torch.int32, device=dev)
            return
        sizes = [p.numel() for p in self.params]
        offs, tot = [], 0
        for n in sizes:
            offs.append(tot)
            tot += (n + 3) // 4 * 4            # keep every tensor 16-byte aligned inside the flat buffer
        self.numel = tot
        self.flat = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.grad_flat = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        for p, o, n in zip(self.params, offs, sizes):
            view = self.flat[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad_flat[o:o + n].view(p.shape)

    def zero_grad(self, set_to_none=False):
        self.grad_flat.zero_()

    def step(self):
        _lib.check(_lib.lib().cvd_adam_flat(
            _lib.ptr(self.flat), _lib.ptr(self.grad_flat), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
            C.c_longlong(self.numel), C.c_float(self.lr), C.c_float(self.betas[0]), C.c_float(self.betas[1]),
            C.c_float(self.eps), C.c_float(self.grad_scale), _lib.ptr(self.state), _lib.ptr(self.loss_flag),
            _lib.stream()), "cvd_adam_flat")


OPTIMIZER_MAP = {
    "Adam": FusedAdam,
}

OPTIMIZER_NAMES = OPTIMIZER_MAP.keys()

OPTIMIZER_CLASSES = OPTIMIZER_MAP.values()


def create(optimizer_name: str, *args, **kwargs):
    return OPTIMIZER_MAP[optimizer_name](*args, **kwargs)
