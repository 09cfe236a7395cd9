"""CPU restatement of the MiDaS-v2 depth network + adapter.

Follows
  monodepth/midas_v2_model.py:52-69           estimate_depth: flatten, (x - mean) / std with ImageNet RGB statistics applied
                                               to the BGR input as is, MidasNet, depth = 1 / output
  monodepth/midas_v2/midas_net.py:49-76       encoder taps layer1..4 -> layerK_rn 3x3 (no bias) -> refinenet4..1 -> output_conv
  monodepth/midas_v2/blocks.py:7-155          _make_encoder (ResNeXt-101 32x8d trunk: conv1, bn1, relu, maxpool, layer1..4),
                                               ResidualConvUnit (nn.ReLU(inplace=True) on its input: the skip adds relu(x)),
                                               FeatureFusionBlock (bilinear x2, align_corners=True), Interpolate (align_corners=False)
  third party, NOT under /root/reference      the trunk is `torch.hub.load("facebookresearch/WSL-Images", "resnext101_32x8d_wsl")`
                                               (blocks.py:27-29, unpinned hub branch): torchvision's
                                               ResNet(Bottleneck, [3, 4, 23, 3], groups=32, width_per_group=8) -- restated here
                                               from torchvision's published definition (Bottleneck v1.5: the stride sits on the 3x3)
BatchNorm runs in TRAIN mode during fine-tuning.  State-dict keys equal `MidasNet().state_dict()` (666 entries).
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/midas_small.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import synth

# (name, inplanes, planes, blocks, stride); width = out channels = 4 * planes for 32x8d
STAGES = [("layer1", 64, 64, 3, 1), ("layer2", 256, 128, 4, 2), ("layer3", 512, 256, 23, 2), ("layer4", 1024, 512, 3, 2)]
GROUPS = 32
FEATURES = 256
NORM_MEAN = (0.485, 0.456, 0.406)
NORM_STD = (0.229, 0.224, 0.225)


def block_prefix(stage, b):
    return f"pretrained.layer1.4.{b}" if stage == "layer1" else f"pretrained.{stage}.{b}"


def midas_param_shapes():
    out = {}

    def bn(p, c):
        out[p + ".weight"] = (c,); out[p + ".bias"] = (c,)
        out[p + ".running_mean"] = (c,); out[p + ".running_var"] = (c,)
        out[p + ".num_batches_tracked"] = ()

    out["pretrained.layer1.0.weight"] = (64, 3, 7, 7)
    bn("pretrained.layer1.1", 64)
    for name, inplanes, planes, blocks, stride in STAGES:
        width = planes * 4
        for b in range(blocks):
            p = block_prefix(name, b)
            cin = inplanes if b == 0 else width
            out[p + ".conv1.weight"] = (width, cin, 1, 1); bn(p + ".bn1", width)
            out[p + ".conv2.weight"] = (width, width // GROUPS, 3, 3); bn(p + ".bn2", width)
            out[p + ".conv3.weight"] = (width, width, 1, 1); bn(p + ".bn3", width)
            if b == 0:
                out[p + ".downsample.0.weight"] = (width, cin, 1, 1); bn(p + ".downsample.1", width)
    for i, c in enumerate((256, 512, 1024, 2048)):
        out[f"scratch.layer{i + 1}_rn.weight"] = (FEATURES, c, 3, 3)
    for r in (4, 3, 2, 1):
        for u in (1, 2):
            for c in (1, 2):
                out[f"scratch.refinenet{r}.resConfUnit{u}.conv{c}.weight"] = (FEATURES, FEATURES, 3, 3)
                out[f"scratch.refinenet{r}.resConfUnit{u}.conv{c}.bias"] = (FEATURES,)
    out["scratch.output_conv.0.weight"] = (128, FEATURES, 3, 3); out["scratch.output_conv.0.bias"] = (128,)
    out["scratch.output_conv.2.weight"] = (32, 128, 3, 3); out["scratch.output_conv.2.bias"] = (32,)
    out["scratch.output_conv.4.weight"] = (1, 32, 1, 1); out["scratch.output_conv.4.bias"] = (1,)
    return out


def is_buffer(k):
    return k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")


def trainable_keys():
    """refinenet4.resConfUnit1 is constructed but never called (FeatureFusionBlock.forward with one input)."""
    return [k for k in midas_param_shapes() if not is_buffer(k) and not k.startswith("scratch.refinenet4.resConfUnit1.")]


def midas_init_state(seed):
    """Deterministic numpy state dict: conv weight/bias ~ U(+-1/sqrt(fan_in)), BN gamma=1 beta=0, running (0,1).
    The last 1x1 conv gets |w| and bias 0.5 so that the (ReLU-clamped) disparity is strictly positive, as a trained
    checkpoint's is: with a raw random head half of the outputs are 0 and depth = 1/0 (midas_v2_model.py:67)."""
    shapes = midas_param_shapes()
    sd = {}
    for i, (k, shp) in enumerate(shapes.items()):
        if k.endswith("running_mean"):
            sd[k] = np.zeros(shp, np.float32)
        elif k.endswith("running_var"):
            sd[k] = np.ones(shp, np.float32)
        elif k.endswith("num_batches_tracked"):
            sd[k] = np.zeros((), np.int64)
        elif ".bn" in k or ".downsample.1." in k or k.startswith("pretrained.layer1.1."):
            sd[k] = (np.ones if k.endswith("weight") else np.zeros)(shp, np.float32)
        else:
            ws = shapes[k[:-5] + ".weight"] if k.endswith(".bias") else shp
            bound = 1.0 / np.sqrt(int(np.prod(ws[1:])))
            sd[k] = synth.uniform(seed, 500 + i, shp, -bound, bound)
    sd["scratch.output_conv.4.weight"] = np.abs(sd["scratch.output_conv.4.weight"])
    sd["scratch.output_conv.4.bias"] = np.full((1,), 0.5, np.float32)
    return sd


def to_torch(sd, dtype=torch.float32, requires_grad=False):
    P, buffers = {}, {}
    live = set(trainable_keys())
    for k, v in sd.items():
        if k.endswith("num_batches_tracked"):
            continue
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if is_buffer(k):
            buffers[k] = t
        else:
            P[k] = t.requires_grad_(requires_grad and k in live)
    return P, buffers


def _bn(x, prefix, P, buffers, train):
    return F.batch_norm(x, buffers.get(prefix + ".running_mean"), buffers.get(prefix + ".running_var"),
                        P[prefix + ".weight"], P[prefix + ".bias"], train, 0.1, 1e-5)


def encoder_forward(x, P, buffers, train=True, capture=None):
    def cap(k, v):
        if capture is not None:
            capture[k] = v.detach()
        return v

    y = cap("pretrained.layer1.0", F.conv2d(x, P["pretrained.layer1.0.weight"], None, stride=2, padding=3))
    t = F.max_pool2d(F.relu(_bn(y, "pretrained.layer1.1", P, buffers, train)), 3, 2, 1)
    taps = []
    for name, inplanes, planes, blocks, stride in STAGES:
        for b in range(blocks):
            p = block_prefix(name, b)
            s = stride if b == 0 else 1
            o = cap(p + ".conv1", F.conv2d(t, P[p + ".conv1.weight"]))
            o = F.relu(_bn(o, p + ".bn1", P, buffers, train))
            o = cap(p + ".conv2", F.conv2d(o, P[p + ".conv2.weight"], None, stride=s, padding=1, groups=GROUPS))
            o = F.relu(_bn(o, p + ".bn2", P, buffers, train))
            o = cap(p + ".conv3", F.conv2d(o, P[p + ".conv3.weight"]))
            o = _bn(o, p + ".bn3", P, buffers, train)
            if b == 0:
                idn = cap(p + ".downsample.0", F.conv2d(t, P[p + ".downsample.0.weight"], None, stride=s))
                idn = _bn(idn, p + ".downsample.1", P, buffers, train)
            else:
                idn = t
            t = cap(p, F.relu(o + idn))
        taps.append(t)
    return taps


def _rcu(x, P, p, capture=None):
    r = F.relu(x)                      # nn.ReLU(inplace=True) on the input: the skip below adds relu(x), not x
    o = F.conv2d(r, P[p + ".conv1.weight"], P[p + ".conv1.bias"], padding=1)
    if capture is not None:
        capture[p + ".conv1"] = o.detach()
    o = F.conv2d(F.relu(o), P[p + ".conv2.weight"], P[p + ".conv2.bias"], padding=1)
    return o + r


def decoder_forward(taps, P, capture=None):
    rn = [F.conv2d(t, P[f"scratch.layer{i + 1}_rn.weight"], None, padding=1) for i, t in enumerate(taps)]
    if capture is not None:
        for i, t in enumerate(rn):
            capture[f"scratch.layer{i + 1}_rn"] = t.detach()
    path = None
    for r in (4, 3, 2, 1):
        p = f"scratch.refinenet{r}"
        out = rn[r - 1] if path is None else path + _rcu(rn[r - 1], P, p + ".resConfUnit1", capture)
        out = _rcu(out, P, p + ".resConfUnit2", capture)
        if capture is not None:
            capture[p] = out.detach()
        path = F.interpolate(out, scale_factor=2, mode="bilinear", align_corners=True)
    h = F.conv2d(path, P["scratch.output_conv.0.weight"], P["scratch.output_conv.0.bias"], padding=1)
    if capture is not None:
        capture["scratch.output_conv.0"] = h.detach()
    h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=False)
    h = F.conv2d(h, P["scratch.output_conv.2.weight"], P["scratch.output_conv.2.bias"], padding=1)
    if capture is not None:
        capture["scratch.output_conv.2"] = h.detach()
    h = F.conv2d(F.relu(h), P["scratch.output_conv.4.weight"], P["scratch.output_conv.4.bias"])
    if capture is not None:
        capture["scratch.output_conv.4"] = h.detach()
    return F.relu(h).squeeze(1)


def estimate_depth(images, P, buffers, train=True, capture=None):
    """midas_v2_model.py:52-69: (...,3,H,W) -> (...,H,W) depth = 1 / MidasNet((x - mean) / std)."""
    shape = images.shape
    C, H, W = shape[-3:]
    x = images.reshape(-1, C, H, W)
    mean = torch.tensor(NORM_MEAN, dtype=x.dtype).reshape(1, -1, 1, 1)
    std = torch.tensor(NORM_STD, dtype=x.dtype).reshape(1, -1, 1, 1)
    out = decoder_forward(encoder_forward((x - mean) / std, P, buffers, train, capture), P, capture)
    if capture is not None:
        capture["disparity"] = out.detach()
    return out.reshape(shape[:-3] + out.shape[-2:]).reciprocal()
