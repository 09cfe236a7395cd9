"""MannequinChallengeModel adapter — drop-in for monodepth/mannequin_challenge_model.py:15-73.

Same class attributes (read by params.py:110-119 before instantiation), zero-argument constructor,
train()/eval()/parameters()/estimate_depth()/save() — but netG is the sm_100a hourglass engine
(mc_engine.McEngine: tcgen05 convs, fused BN/pool/upsample kernels) instead of a torch nn.Module tree.
Weights: `checkpoints/mc.pth` (the file the reference caches, utils/url_helpers.py:14-16) if present,
else a deterministic default-scale initialisation (this sandbox has no network for the download).
"""
import math
import os

import torch

from .. import optimizer as _optimizer
from .depth_model import DepthModel
from .mc_engine import McEngine, McParams


class _EngineFn(torch.autograd.Function):
    """depth = engine.forward(images); backward runs the engine's hand-written backward, which writes the
    parameter gradients straight into the flat gradient buffer (the .grad views of parameters())."""

    @staticmethod
    def forward(ctx, images, anchor, engine):
        ctx.engine = engine
        return engine.forward(images).clone()

    @staticmethod
    def backward(ctx, grad_depth):
        ctx.engine.backward(grad_depth.contiguous())
        return None, None, None


def default_init_state(seed=0):
    """Conv weight/bias ~ U(+-1/sqrt(fan_in)) (torch's default scale), BN gamma=1 beta=0, running (0,1)."""
    from . import mc_arch
    g = torch.Generator().manual_seed(seed)
    shapes = mc_arch.state_dict_shapes()
    sd = {}
    for k, s in shapes.items():
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(s)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(s)
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif k == "seq.1.weight":
            sd[k] = torch.ones(s)
        elif k == "seq.1.bias":
            sd[k] = torch.zeros(s)
        else:
            ws = shapes[k[:-5] + ".weight"] if k.endswith(".bias") else s
            bound = 1.0 / math.sqrt(ws[1] * ws[2] * ws[3])
            sd[k] = (torch.rand(s, generator=g) * 2 - 1) * bound
    return sd


def _strip_module(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


class MannequinChallengeModel(DepthModel):
    # Requirements and default settings (mannequin_challenge_model.py:17-19)
    align = 16
    learning_rate = 0.0004
    lambda_view_baseline = 0.1

    def __init__(self, state_dict=None, precision=3):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("MannequinChallengeModel (consistent_depth_b200) needs a CUDA device: there is no CPU path")
        self.device_ = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.P = McParams(self.device_)
        if state_dict is None:
            path = os.path.join("checkpoints", "mc.pth")
            if os.path.isfile(path):
                state_dict = torch.load(path, map_location="cpu")
            else:
                state_dict = default_init_state(0)
        self.P.load_state_dict(_strip_module(state_dict))
        self.engines = {}
        self.training_ = True
        self._anchor = torch.zeros((), device=self.device_, requires_grad=True)
        self._params = None

    def engine(self, n, H, W):
        key = (n, H, W)
        e = self.engines.get(key)
        if e is None:
            e = McEngine(self.P, n, H, W, self.precision)
            self.engines[key] = e
        e.train_mode = self.training_
        return e

    def train(self, mode=True):
        self.training_ = bool(mode)
        return self

    def eval(self):
        self.training_ = False
        return self

    def parameters(self):
        """The 314 trainable tensors as leaf views of ONE flat buffer (FusedAdam adopts it as-is).
        uncertainty_layer is excluded: the reference never gives it a gradient, so Adam skips it."""
        if self._params is None:
            plist = _optimizer.FlatParamList()
            plist.flat, plist.grad_flat = self.P.flat, self.P.grad_flat
            for k, t in self.P.named_parameters():
                p = torch.nn.Parameter(t, requires_grad=True)
                p.data = t
                p.grad = self.P._g(k)
                plist.append(p)
            self._params = plist
        return self._params

    def estimate_depth(self, images, metadata=None):
        images = images.to(self.device_, dtype=torch.float32)
        shape = images.shape
        C, H, W = shape[-3:]
        flat = images.reshape(-1, C, H, W).contiguous()
        eng = self.engine(flat.shape[0], H, W)
        if torch.is_grad_enabled():
            depth = _EngineFn.apply(flat, self._anchor, eng)
        else:
            depth = eng.forward(flat).clone()
        return depth.reshape(shape[:-3] + (H, W))

    def state_dict(self, *args, **kwargs):
        return self.P.state_dict()

    def load_state_dict(self, sd, strict=True):
        """Accepts the reference's checkpoint format (keys prefixed `module.`: netG is DataParallel-wrapped,
        pix2pix_model.py:108) and bare HourglassModel keys alike."""
        self.P.load_state_dict(_strip_module(sd))

    def save(self, file_name):
        """checkpoints/%04d.pth in the reference's format: `netG.state_dict()` of the DataParallel-wrapped hourglass
        (mannequin_challenge_model.py:71-73), i.e. every key carries the `module.` prefix, so the files are
        interchangeable with the reference in both directions."""
        torch.save({"module." + k: v for k, v in self.P.state_dict().items()}, file_name)
