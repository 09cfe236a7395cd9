"""Architecture table of MiDaS v2 (MidasNet: ResNeXt-101 32x8d trunk + 4 feature-fusion blocks + output head).

Transcribes as data
  monodepth/midas_v2/blocks.py:7-57        _make_encoder / _make_resnet_backbone / _make_scratch
                                            (trunk = torchvision ResNet(Bottleneck, [3,4,23,3], groups=32, width_per_group=8),
                                            pulled through torch.hub in the reference, blocks.py:27-29)
  monodepth/midas_v2/midas_net.py:15-47    refinenet4..1 (FeatureFusionBlock), output_conv
and the key names of `MidasNet().state_dict()` (666 entries), so checkpoints (model-f46da743.pt) load unchanged.
"""

# (name, inplanes, planes, blocks, stride); for 32x8d: conv width = output channels = 4 * planes
STAGES = [("layer1", 64, 64, 3, 1), ("layer2", 256, 128, 4, 2), ("layer3", 512, 256, 23, 2), ("layer4", 1024, 512, 3, 2)]
GROUPS = 32
FEATURES = 256
NORM_MEAN = (0.485, 0.456, 0.406)      # midas_v2_model.py:46-49 (ImageNet RGB statistics, applied to the BGR input as is)
NORM_STD = (0.229, 0.224, 0.225)


def block_prefix(stage, b):
    """pretrained.layer1 is Sequential(conv1, bn1, relu, maxpool, layer1) (blocks.py:15-17): its blocks sit at index 4."""
    return f"pretrained.layer1.4.{b}" if stage == "layer1" else f"pretrained.{stage}.{b}"


def state_dict_shapes():
    out = {}

    def bn(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)
        out[p + ".running_mean"] = (c,)
        out[p + ".running_var"] = (c,)
        out[p + ".num_batches_tracked"] = ()

    out["pretrained.layer1.0.weight"] = (64, 3, 7, 7)
    bn("pretrained.layer1.1", 64)
    for name, inplanes, planes, blocks, stride in STAGES:
        width = planes * 4
        for b in range(blocks):
            p = block_prefix(name, b)
            cin = inplanes if b == 0 else width
            out[p + ".conv1.weight"] = (width, cin, 1, 1)
            bn(p + ".bn1", width)
            out[p + ".conv2.weight"] = (width, width // GROUPS, 3, 3)
            bn(p + ".bn2", width)
            out[p + ".conv3.weight"] = (width, width, 1, 1)
            bn(p + ".bn3", width)
            if b == 0:
                out[p + ".downsample.0.weight"] = (width, cin, 1, 1)
                bn(p + ".downsample.1", width)
    for i, c in enumerate((256, 512, 1024, 2048)):
        out[f"scratch.layer{i + 1}_rn.weight"] = (FEATURES, c, 3, 3)
    for r in (4, 3, 2, 1):
        for u in (1, 2):
            for c in (1, 2):
                out[f"scratch.refinenet{r}.resConfUnit{u}.conv{c}.weight"] = (FEATURES, FEATURES, 3, 3)
                out[f"scratch.refinenet{r}.resConfUnit{u}.conv{c}.bias"] = (FEATURES,)
    out["scratch.output_conv.0.weight"] = (128, FEATURES, 3, 3)
    out["scratch.output_conv.0.bias"] = (128,)
    out["scratch.output_conv.2.weight"] = (32, 128, 3, 3)
    out["scratch.output_conv.2.bias"] = (32,)
    out["scratch.output_conv.4.weight"] = (1, 32, 1, 1)
    out["scratch.output_conv.4.bias"] = (1,)
    return out


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def dead_parameter(key):
    """refinenet4.resConfUnit1 is constructed but never called (FeatureFusionBlock.forward with a single input,
    midas_net.py:70): Adam owns it, it never receives a gradient."""
    return key.startswith("scratch.refinenet4.resConfUnit1.")
