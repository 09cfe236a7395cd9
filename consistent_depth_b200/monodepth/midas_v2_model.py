"""MidasV2Model adapter — drop-in for monodepth/midas_v2_model.py:12-73.

Same class attributes (read by params.py:110-119 before instantiation), constructor arguments
(`support_cpu`, `pretrained`), train()/eval()/parameters()/estimate_depth()/save() — MidasNet runs on the sm_100a
engine (midas_engine.MidasEngine).  Weights: `checkpoints/model-f46da743.pt` (the file torch.hub caches for the
reference, midas_v2_model.py:33-39) if present, else a deterministic default-scale initialisation whose last layer is
made positive so that the ReLU-clamped disparity is > 0 the way a trained checkpoint's is (this sandbox has no network).
There is no CPU path: `support_cpu=True` without a GPU raises instead of silently computing elsewhere.
"""
import math
import os

import torch

from .. import optimizer as _optimizer
from . import midas_arch as arch
from .depth_model import DepthModel
from .mannequin_challenge_model import _EngineFn
from .midas_engine import MidasEngine, MidasParams


def default_init_state(seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = arch.state_dict_shapes()
    sd = {}
    for k, s in shapes.items():
        if k.endswith("running_mean"):
            sd[k] = torch.zeros(s)
        elif k.endswith("running_var"):
            sd[k] = torch.ones(s)
        elif k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif ".bn" in k or ".downsample.1." in k or k.startswith("pretrained.layer1.1."):
            sd[k] = torch.ones(s) if k.endswith("weight") else torch.zeros(s)
        else:
            ws = shapes[k[:-5] + ".weight"] if k.endswith(".bias") else s
            fan_in = 1
            for v in ws[1:]:
                fan_in *= v
            sd[k] = (torch.rand(s, generator=g) * 2 - 1) / math.sqrt(fan_in)
    sd["scratch.output_conv.4.weight"] = sd["scratch.output_conv.4.weight"].abs()
    sd["scratch.output_conv.4.bias"] = torch.full((1,), 0.5)
    return sd


class MidasV2Model(DepthModel):
    # Requirements and default settings (midas_v2_model.py:13-16)
    align = 32
    learning_rate = 0.0001
    lambda_view_baseline = 0.0001

    def __init__(self, support_cpu: bool = False, pretrained: bool = True, state_dict=None, precision=3):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("MidasV2Model (consistent_depth_b200) needs a CUDA device: there is no CPU path")
        self.device_ = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.P = MidasParams(self.device_)
        if state_dict is None:
            path = os.path.join("checkpoints", "model-f46da743.pt")
            if pretrained and os.path.isfile(path):
                state_dict = torch.load(path, map_location="cpu")
            else:
                state_dict = default_init_state(0)
        self.P.load_state_dict(state_dict)
        self.engines = {}
        self.training_ = True
        self._anchor = torch.zeros((), device=self.device_, requires_grad=True)
        self._params = None

    def engine(self, n, H, W):
        key = (n, H, W)
        e = self.engines.get(key)
        if e is None:
            e = MidasEngine(self.P, n, H, W, self.precision)
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
        """The 354 tensors of MidasNet.parameters() as leaf views of ONE flat buffer; refinenet4.resConfUnit1
        (constructed, never called) keeps a zero gradient."""
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
        # the reference wraps MidasNet in DataParallel only when it sees several GPUs (midas_v2_model.py:41-43): accept both
        self.P.load_state_dict({(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()})

    def save(self, file_name):
        torch.save(self.P.state_dict(), file_name)
