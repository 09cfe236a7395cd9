"""Monodepth2Model adapter — drop-in for monodepth/monodepth2_model.py:15-93.

Same class attributes (read by params.py:110-119 before instantiation), zero-argument constructor,
train()/eval()/parameters()/estimate_depth()/save() — the ResNet-18 encoder and the depth decoder run on the
sm_100a engine (mono2_engine.Mono2Engine) instead of torch nn.Modules.
Weights: `checkpoints/monodepth2_mono+stereo_1024x320/{encoder,depth}.pth` (the directory the reference unpacks its
download into, monodepth2_model.py:26-29) if present, including the `height`/`width` entries that fix the feed size
(:35-36); else a deterministic default-scale initialisation with the stock checkpoint's 320 x 1024 feed
(this sandbox has no network for the download).  As in the reference, save() is a no-op (:92-93).
"""
import math
import os

import torch

from .. import optimizer as _optimizer
from . import mono2_arch as arch
from .depth_model import DepthModel
from .mannequin_challenge_model import _EngineFn
from .mono2_engine import Mono2Engine, Mono2Params

STOCK_FEED = (320, 1024)          # mono+stereo_1024x320 (monodepth2_model.py:26)


def default_init_state(seed=0):
    """Conv / linear weight and bias ~ U(+-1/sqrt(fan_in)) (torch's default scale), BN gamma=1 beta=0, running (0,1)."""
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
        elif ".bn" in k or ".downsample.1." in k:
            sd[k] = torch.ones(s) if k.endswith("weight") else torch.zeros(s)
        else:
            ws = shapes[k[:-5] + ".weight"] if k.endswith(".bias") else s
            fan_in = 1
            for v in ws[1:]:
                fan_in *= v
            sd[k] = (torch.rand(s, generator=g) * 2 - 1) / math.sqrt(fan_in)
    return sd


class Monodepth2Model(DepthModel):
    # Requirements and default settings (monodepth2_model.py:16-19)
    align = 1
    learning_rate = 0.00004
    lambda_view_baseline = 1

    def __init__(self, state_dict=None, feed_size=None, precision=3):
        super().__init__()
        if not torch.cuda.is_available():
            raise RuntimeError("Monodepth2Model (consistent_depth_b200) needs a CUDA device: there is no CPU path")
        self.device_ = torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.P = Mono2Params(self.device_)
        if state_dict is None:
            d = os.path.join("checkpoints", "monodepth2_mono+stereo_1024x320")
            enc, dec = os.path.join(d, "encoder.pth"), os.path.join(d, "depth.pth")
            if os.path.isfile(enc) and os.path.isfile(dec):
                state_dict = dict(torch.load(enc, map_location="cpu"))
                state_dict.update(torch.load(dec, map_location="cpu"))
            else:
                state_dict = default_init_state(0)
        self.P.load_state_dict(state_dict)
        self.feed_height, self.feed_width = feed_size or self.P.feed_size or STOCK_FEED
        self.engines = {}
        self.training_ = True
        self._anchor = torch.zeros((), device=self.device_, requires_grad=True)
        self._params = None

    def engine(self, n, H, W):
        key = (n, H, W)
        e = self.engines.get(key)
        if e is None:
            e = Mono2Engine(self.P, n, H, W, (self.feed_height, self.feed_width), self.precision)
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
        """All 90 tensors of encoder.parameters() + depth_decoder.parameters() (monodepth2_model.py:58-59) as leaf views
        of ONE flat buffer; the classifier head and the unused disparity heads keep a zero gradient."""
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
        sd = self.P.state_dict()
        sd["height"], sd["width"] = self.feed_height, self.feed_width
        return sd

    def load_state_dict(self, sd, strict=True):
        self.P.load_state_dict(sd)

    def save(self, file_name):
        pass
