"""CPU restatement of the mannequin-challenge hourglass depth network + adapter.

Follows monodepth/mannequin_challenge/models/hourglass.py:
  inception (:19-55)    branch0 = 1x1 conv + BN(affine=False) + ReLU ; branch i = 1x1 + BN + ReLU + kxk("same") + BN + ReLU ; cat
  Channels1-4 (:58-156) recursive hourglass: skip branch + (AvgPool2d(2) .. UpsamplingBilinear2d(2)) branch, summed
  HourglassModel (:159-181) Conv7x7(3->128)+BN(affine)+ReLU -> Channels4 -> pred_layer 3x3(64->1)
                           (uncertainty_layer output is discarded by the adapter, mannequin_challenge_model.py:60)
and monodepth/mannequin_challenge_model.py:52-69 (estimate_depth: flatten to (2B,3,H,W), netG, reshape, exp).
BatchNorm runs in TRAIN mode (batch statistics, eps 1e-5, momentum 0.1, running
stats updated) because depth_fine_tuning.py:241 calls model.train().

State-dict key names equal the reference's `HourglassModel(3).state_dict()` so a
weight dict moves between the two unchanged.  Written functionally (no nn.Module
tree) so it doubles as the layer-by-layer spec of the CUDA engine's plan.
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/hourglass_*.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import synth

# (module path prefix, config) in execution order is produced by `mc_layer_table()`.
_A = [[16], [3, 64, 16], [7, 64, 16], [11, 64, 16]]          # Channels4 skip  (in 128)
_B = [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]]           # in 128
_C = [[32], [3, 64, 32], [7, 64, 32], [11, 64, 32]]          # in 128
_D = [[64], [3, 32, 64], [5, 32, 64], [7, 32, 64]]           # in 128 -> 256
_E = [[64], [3, 32, 64], [5, 32, 64], [7, 32, 64]]           # in 256
_F = [[64], [3, 64, 64], [7, 64, 64], [11, 64, 64]]          # in 256
_G = [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]]           # in 256 -> 128
_B2 = [[32], [3, 64, 32], [5, 64, 32], [7, 64, 32]]          # Channels4 "B" after Channels3 (in 128)
_A2 = [[16], [3, 32, 16], [7, 32, 16], [11, 32, 16]]         # Channels4 last (in 128 -> 64)


def mc_structure():
    """Nested description of HourglassModel.seq[3] (Channels4), hourglass.py:58-156.

    ("inc", cin, config) | ("pool",) | ("up",) | ("chan", [branch0 ops], [branch1 ops]).
    Branch order matches self.list[0], self.list[1] of each ChannelsN.
    """
    inc = lambda cin, cfg: ("inc", cin, cfg)
    ch1 = ("chan",
           [inc(256, _E), inc(256, _E)],
           [("pool",), inc(256, _E), inc(256, _E), inc(256, _E), ("up",)])
    ch2 = ("chan",
           [inc(256, _E), inc(256, _F)],
           [("pool",), inc(256, _E), inc(256, _E), ch1, inc(256, _E), inc(256, _F), ("up",)])
    ch3 = ("chan",
           [("pool",), inc(128, _B), inc(128, _D), ch2, inc(256, _E), inc(256, _G), ("up",)],
           [inc(128, _B), inc(128, _C)])
    ch4 = ("chan",
           [("pool",), inc(128, _B), inc(128, _B), ch3, inc(128, _B2), inc(128, _A2), ("up",)],
           [inc(128, _A)])
    return ch4


def mc_param_shapes():
    """Ordered {state_dict key: shape} of HourglassModel(3) (params and BN buffers)."""
    out = {}

    def conv(prefix, cin, cout, k):
        out[prefix + ".weight"] = (cout, cin, k, k)
        out[prefix + ".bias"] = (cout,)

    def bn(prefix, c, affine):
        if affine:
            out[prefix + ".weight"] = (c,); out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,); out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    def walk(node, prefix):
        kind = node[0]
        if kind == "inc":
            _, cin, cfg = node
            conv(f"{prefix}.convs.0.0", cin, cfg[0][0], 1); bn(f"{prefix}.convs.0.1", cfg[0][0], False)
            for i in range(1, len(cfg)):
                k, a, b = cfg[i]
                conv(f"{prefix}.convs.{i}.0", cin, a, 1); bn(f"{prefix}.convs.{i}.1", a, False)
                conv(f"{prefix}.convs.{i}.3", a, b, k); bn(f"{prefix}.convs.{i}.4", b, False)
        elif kind == "chan":
            for bi, ops in enumerate(node[1:]):
                for oi, op in enumerate(ops):
                    walk(op, f"{prefix}.list.{bi}.{oi}")

    conv("seq.0", 3, 128, 7); bn("seq.1", 128, True)
    walk(mc_structure(), "seq.3")
    conv("uncertainty_layer.0", 64, 1, 3)
    conv("pred_layer", 64, 1, 3)
    return out


def mc_init_state(seed):
    """Deterministic state dict (numpy) with torch's default init SCALES:
    conv weight/bias ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); BN gamma=1, beta=0, running (0,1)."""
    sd = {}
    for i, (k, shp) in enumerate(mc_param_shapes().items()):
        if k.endswith("running_mean"):
            sd[k] = np.zeros(shp, np.float32)
        elif k.endswith("running_var"):
            sd[k] = np.ones(shp, np.float32)
        elif k.endswith("num_batches_tracked"):
            sd[k] = np.zeros((), np.int64)
        elif k == "seq.1.weight":
            sd[k] = np.ones(shp, np.float32)
        elif k == "seq.1.bias":
            sd[k] = np.zeros(shp, np.float32)
        else:
            wk = k[:-5] + ".weight" if k.endswith(".bias") else k
            wshape = mc_param_shapes()[wk] if k.endswith(".bias") else shp
            fan_in = wshape[1] * wshape[2] * wshape[3]
            bound = 1.0 / np.sqrt(fan_in)
            sd[k] = synth.uniform(seed, 100 + i, shp, -bound, bound)
    return sd


def _bn_train(x, prefix, P, buffers, affine, momentum=0.1, eps=1e-5):
    rm, rv = buffers.get(prefix + ".running_mean"), buffers.get(prefix + ".running_var")
    w = P[prefix + ".weight"] if affine else None
    b = P[prefix + ".bias"] if affine else None
    return F.batch_norm(x, rm, rv, w, b, True, momentum, eps)


def hourglass_forward(x, P, buffers=None, capture=None):
    """x (N,3,H,W) -> log-depth (N,1,H,W).  P: dict key->tensor (requires_grad for training),
    buffers: dict of BN running stats, updated in place (may be {} to ignore);
    capture: optional dict filled with every conv's raw (pre-BatchNorm) output, keyed by conv prefix."""
    buffers = {} if buffers is None else buffers

    def conv_bn_relu(x, prefix_conv, prefix_bn, pad, affine=False):
        y = F.conv2d(x, P[prefix_conv + ".weight"], P[prefix_conv + ".bias"], padding=pad)
        if capture is not None:
            capture[prefix_conv] = y.detach()
        return F.relu(_bn_train(y, prefix_bn, P, buffers, affine))

    def inception(x, prefix, cfg):
        outs = [conv_bn_relu(x, f"{prefix}.convs.0.0", f"{prefix}.convs.0.1", 0)]
        for i in range(1, len(cfg)):
            k = cfg[i][0]
            t = conv_bn_relu(x, f"{prefix}.convs.{i}.0", f"{prefix}.convs.{i}.1", 0)
            outs.append(conv_bn_relu(t, f"{prefix}.convs.{i}.3", f"{prefix}.convs.{i}.4", (k - 1) // 2))
        return torch.cat(outs, dim=1)

    def run(node, x, prefix):
        kind = node[0]
        if kind == "inc":
            return inception(x, prefix, node[2])
        if kind == "pool":
            return F.avg_pool2d(x, 2)
        if kind == "up":
            return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        res = []
        for bi, ops in enumerate(node[1:]):
            t = x
            for oi, op in enumerate(ops):
                t = run(op, t, f"{prefix}.list.{bi}.{oi}")
            res.append(t)
        return res[0] + res[1]

    y = conv_bn_relu(x, "seq.0", "seq.1", 3, affine=True)
    y = run(mc_structure(), y, "seq.3")
    return F.conv2d(y, P["pred_layer.weight"], P["pred_layer.bias"], padding=1)


def estimate_depth(images, P, buffers=None, capture=None):
    """mannequin_challenge_model.py:52-69: (...,3,H,W) -> (...,H,W) depth = exp(log-depth)."""
    shape = images.shape
    C, H, W = shape[-3:]
    pred = hourglass_forward(images.reshape(-1, C, H, W), P, buffers, capture)
    pred = pred.reshape(shape[:-3] + pred.shape[-2:])     # X1HW -> (...,H,W)
    return torch.exp(pred)


def trainable_keys():
    """Keys Adam actually updates: everything with a gradient.  uncertainty_layer never gets one
    (hourglass.py:179 output dropped at mannequin_challenge_model.py:60) so Adam skips it."""
    return [k for k in mc_param_shapes()
            if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"))
            and not k.startswith("uncertainty_layer")]


def to_torch(sd, dtype=torch.float32, requires_grad=False):
    P, buffers = {}, {}
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if k.endswith("running_mean") or k.endswith("running_var"):
            buffers[k] = t
        else:
            P[k] = t.requires_grad_(requires_grad and not k.startswith("uncertainty_layer"))
    return P, buffers
