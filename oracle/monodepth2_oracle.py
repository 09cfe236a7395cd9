"""CPU restatement of the monodepth2 depth network + adapter.

Follows
  monodepth/monodepth2_model.py:63-89          estimate_depth: flatten to (N,3,H,W), bicubic resize to the
                                                 checkpoint's feed size (align_corners=False), encoder, decoder,
                                                 disp0 -> bicubic resize back -> depth = 1/disp
  monodepth/monodepth2/networks/resnet_encoder.py:87-98   (x-0.45)/0.225 -> torchvision resnet18 trunk, 5 feature taps
                                                 (relu(bn1(conv1)), layer1(maxpool), layer2, layer3, layer4)
  torchvision resnet18 BasicBlock              conv3x3(stride)-BN-ReLU-conv3x3-BN (+ 1x1(stride)-BN downsample) + add, ReLU
  monodepth/monodepth2/networks/depth_decoder.py:50-65    for i=4..0: ConvBlock -> nearest x2 -> cat skip -> ConvBlock;
                                                 disp_i = sigmoid(Conv3x3) for i in scales (only disp_0 is consumed)
  monodepth/monodepth2/layers.py:106-136,196-199  ConvBlock = Conv3x3 (ReflectionPad2d(1) + 3x3 conv) + ELU; upsample
BatchNorm runs in TRAIN mode during fine-tuning (monodepth2_model.py:50-52 via depth_fine_tuning.py:241).

State-dict keys are the union of `ResnetEncoder(18).state_dict()` ("encoder.*", the fc layer included: it is a
parameter Adam owns but never gets a gradient) and `DepthDecoder.state_dict()` ("decoder.N.*"), i.e. the two files
encoder.pth / depth.pth of the stock checkpoint merged.  The bicubic resize is restated explicitly
(`bicubic_resize`) because the CUDA kernels implement exactly this index arithmetic.
TEST INFRASTRUCTURE ONLY.  Pinned by tests/golden/monodepth2_small.npz.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import synth

NUM_CH_ENC = [64, 64, 128, 256, 512]
NUM_CH_DEC = [16, 32, 64, 128, 256]
LAYERS = [("layer1", 64, 64, 1), ("layer2", 64, 128, 2), ("layer3", 128, 256, 2), ("layer4", 256, 512, 2)]


def decoder_index(i, j):
    """Index of ("upconv", i, j) in DepthDecoder.decoder (ModuleList in insertion order, depth_decoder.py:31-48)."""
    return (4 - i) * 2 + j


def mono2_param_shapes():
    """Ordered {state_dict key: shape}: ResnetEncoder(18) then DepthDecoder(scales=range(4))."""
    out = {}

    def bn(p, c):
        out[p + ".weight"] = (c,); out[p + ".bias"] = (c,)
        out[p + ".running_mean"] = (c,); out[p + ".running_var"] = (c,)
        out[p + ".num_batches_tracked"] = ()

    out["encoder.conv1.weight"] = (64, 3, 7, 7)
    bn("encoder.bn1", 64)
    for name, cin, cout, stride in LAYERS:
        for b in range(2):
            p = f"encoder.{name}.{b}"
            out[p + ".conv1.weight"] = (cout, cin if b == 0 else cout, 3, 3)
            bn(p + ".bn1", cout)
            out[p + ".conv2.weight"] = (cout, cout, 3, 3)
            bn(p + ".bn2", cout)
            if b == 0 and (stride != 1 or cin != cout):
                out[p + ".downsample.0.weight"] = (cout, cin, 1, 1)
                bn(p + ".downsample.1", cout)
    out["encoder.fc.weight"] = (1000, 512)
    out["encoder.fc.bias"] = (1000,)
    for i in range(4, -1, -1):
        cin0 = NUM_CH_ENC[4] if i == 4 else NUM_CH_DEC[i + 1]
        cin1 = NUM_CH_DEC[i] + (NUM_CH_ENC[i - 1] if i > 0 else 0)
        for j, cin in ((0, cin0), (1, cin1)):
            p = f"decoder.{decoder_index(i, j)}.conv.conv"
            out[p + ".weight"] = (NUM_CH_DEC[i], cin, 3, 3)
            out[p + ".bias"] = (NUM_CH_DEC[i],)
    for s in range(4):
        out[f"decoder.{10 + s}.conv.weight"] = (1, NUM_CH_DEC[s], 3, 3)
        out[f"decoder.{10 + s}.conv.bias"] = (1,)
    return out


def is_buffer(k):
    return k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked")


def trainable_keys():
    """Keys that receive a gradient: not the fc layer (never called, resnet_encoder.py:87-98) and not the
    scale-1..3 disparity heads (computed, never consumed: monodepth2_model.py:78)."""
    dead = ("encoder.fc.",) + tuple(f"decoder.{10 + s}." for s in (1, 2, 3))
    return [k for k in mono2_param_shapes() if not is_buffer(k) and not k.startswith(dead)]


def mono2_init_state(seed):
    """Deterministic numpy state dict with torch's default Conv2d/Linear init SCALE U(+-1/sqrt(fan_in)),
    BN gamma=1, beta=0, running stats (0,1)."""
    shapes = mono2_param_shapes()
    sd = {}
    for i, (k, shp) in enumerate(shapes.items()):
        if k.endswith("running_mean"):
            sd[k] = np.zeros(shp, np.float32)
        elif k.endswith("running_var"):
            sd[k] = np.ones(shp, np.float32)
        elif k.endswith("num_batches_tracked"):
            sd[k] = np.zeros((), np.int64)
        elif ".bn" in k or ".downsample.1." in k:
            sd[k] = (np.ones if k.endswith("weight") else np.zeros)(shp, np.float32)
        else:
            ws = shapes[k[:-5] + ".weight"] if k.endswith(".bias") else shp
            fan_in = int(np.prod(ws[1:]))
            bound = 1.0 / np.sqrt(fan_in)
            sd[k] = synth.uniform(seed, 300 + i, shp, -bound, bound)
    return sd


def to_torch(sd, dtype=torch.float32, requires_grad=False):
    P, buffers = {}, {}
    live = set(trainable_keys())
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k in ("height", "width"):
            continue
        t = torch.tensor(np.asarray(v), dtype=dtype)
        if is_buffer(k):
            buffers[k] = t
        else:
            P[k] = t.requires_grad_(requires_grad and k in live)
    return P, buffers


# ------------------------------------------------------------------ bicubic resize (align_corners=False)
def _cubic_weights(t, A=-0.75):
    """Keys cubic convolution coefficients for taps at floor-1 .. floor+2 (torch upsample_bicubic2d, A=-0.75)."""
    def near(x):      # |x| <= 1
        return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0

    def far(x):       # 1 < |x| < 2
        return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A
    return [far(t + 1.0), near(t), near(1.0 - t), far(2.0 - t)]


def bicubic_axis(n_in, n_out, dtype=torch.float64):
    """(indices (n_out,4) int64 clamped to [0,n_in-1], weights (n_out,4)) of one axis:
    src = (dst + 0.5) * n_in / n_out - 0.5 (NOT clamped at 0 for cubic), taps floor(src)-1 .. floor(src)+2."""
    dst = torch.arange(n_out, dtype=dtype)
    src = (dst + 0.5) * (float(n_in) / float(n_out)) - 0.5
    fl = torch.floor(src)
    t = src - fl
    w = torch.stack(_cubic_weights(t), 1)
    idx = (fl.long()[:, None] + torch.arange(-1, 3)[None, :]).clamp_(0, n_in - 1)
    return idx, w


def bicubic_resize(x, oh, ow):
    """F.interpolate(x, size=(oh, ow), mode='bicubic', align_corners=False), restated as two separable
    4-tap gathers with border-clamped indices."""
    N, C, H, W = x.shape
    iy, wy = bicubic_axis(H, oh, x.dtype)
    ix, wx = bicubic_axis(W, ow, x.dtype)
    rows = (x[:, :, iy.reshape(-1), :].reshape(N, C, oh, 4, W) * wy[None, None, :, :, None]).sum(3)
    return (rows[:, :, :, ix.reshape(-1)].reshape(N, C, oh, ow, 4) * wx[None, None, None, :, :]).sum(4)


# ------------------------------------------------------------------ network
def _bn(x, prefix, P, buffers, train, momentum=0.1, eps=1e-5):
    rm, rv = buffers.get(prefix + ".running_mean"), buffers.get(prefix + ".running_var")
    if not train:
        return F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], False, momentum, eps)
    return F.batch_norm(x, rm, rv, P[prefix + ".weight"], P[prefix + ".bias"], True, momentum, eps)


def encoder_forward(x, P, buffers, train=True, capture=None):
    """resnet_encoder.py:87-98 on an already resized image batch (N,3,h,w) -> 5 feature maps."""
    def cap(k, v):
        if capture is not None:
            capture[k] = v.detach()
        return v

    x = (x - 0.45) / 0.225
    y = cap("encoder.conv1", F.conv2d(x, P["encoder.conv1.weight"], None, stride=2, padding=3))
    feats = [F.relu(_bn(y, "encoder.bn1", P, buffers, train))]
    t = F.max_pool2d(feats[0], kernel_size=3, stride=2, padding=1)
    for name, cin, cout, stride in LAYERS:
        for b in range(2):
            p = f"encoder.{name}.{b}"
            s = stride if b == 0 else 1
            o = cap(p + ".conv1", F.conv2d(t, P[p + ".conv1.weight"], None, stride=s, padding=1))
            o = F.relu(_bn(o, p + ".bn1", P, buffers, train))
            o = cap(p + ".conv2", F.conv2d(o, P[p + ".conv2.weight"], None, stride=1, padding=1))
            o = _bn(o, p + ".bn2", P, buffers, train)
            if p + ".downsample.0.weight" in P:
                idn = cap(p + ".downsample.0", F.conv2d(t, P[p + ".downsample.0.weight"], None, stride=s))
                idn = _bn(idn, p + ".downsample.1", P, buffers, train)
            else:
                idn = t
            t = cap(p, F.relu(o + idn))
        feats.append(t)
    return feats


def _conv3x3_reflect(x, P, prefix):
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), P[prefix + ".weight"], P[prefix + ".bias"])


def decoder_forward(feats, P, capture=None):
    """depth_decoder.py:50-65, scale-0 disparity only (scales 1-3 are computed by the reference but unused)."""
    x = feats[-1]
    for i in range(4, -1, -1):
        p0, p1 = f"decoder.{decoder_index(i, 0)}.conv.conv", f"decoder.{decoder_index(i, 1)}.conv.conv"
        y = _conv3x3_reflect(x, P, p0)
        if capture is not None:
            capture[p0] = y.detach()
        x = F.elu(y)
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if i > 0:
            x = torch.cat([x, feats[i - 1]], 1)
        y = _conv3x3_reflect(x, P, p1)
        if capture is not None:
            capture[p1] = y.detach()
        x = F.elu(y)
    y = _conv3x3_reflect(x, P, "decoder.10.conv")
    if capture is not None:
        capture["decoder.10.conv"] = y.detach()
    return torch.sigmoid(y)


def estimate_depth(images, P, buffers, feed_size, train=True, capture=None):
    """monodepth2_model.py:63-89: (...,3,H,W) -> (...,H,W) depth = 1 / bicubic(disp0)."""
    shape = images.shape
    C, H, W = shape[-3:]
    x = bicubic_resize(images.reshape(-1, C, H, W), feed_size[0], feed_size[1])
    if capture is not None:
        capture["resized"] = x.detach()
    disp = decoder_forward(encoder_forward(x, P, buffers, train, capture), P, capture)
    if capture is not None:
        capture["disp0"] = disp.detach()
    disp = bicubic_resize(disp, H, W)
    depth = disp.reciprocal()
    return depth.reshape(shape[:-3] + depth.shape[-2:])
