"""Architecture table of the monodepth2 depth network (ResNet-18 encoder + DepthDecoder).

Transcribes as data
  monodepth/monodepth2/networks/resnet_encoder.py:60-98  (torchvision resnet18: conv1 7x7/2, maxpool 3x3/2,
                                                          layer1..4 of two BasicBlocks, widths 64/128/256/512)
  monodepth/monodepth2/networks/depth_decoder.py:16-65   (num_ch_dec = 16,32,64,128,256; ("upconv", i, 0/1), ("dispconv", s))
and the key names of `ResnetEncoder(18).state_dict()` ("encoder.*") + `DepthDecoder.state_dict()` ("decoder.N.*"),
i.e. encoder.pth and depth.pth of the stock checkpoint merged, so weights move both ways unchanged.
"""

NUM_CH_ENC = [64, 64, 128, 256, 512]
NUM_CH_DEC = [16, 32, 64, 128, 256]
LAYERS = [("layer1", 64, 64, 1), ("layer2", 64, 128, 2), ("layer3", 128, 256, 2), ("layer4", 256, 512, 2)]


def decoder_index(i, j):
    """Position of ("upconv", i, j) in DepthDecoder.decoder (a ModuleList in insertion order, depth_decoder.py:31-48)."""
    return (4 - i) * 2 + j


def upconv_key(i, j):
    return f"decoder.{decoder_index(i, j)}.conv.conv"


def block_has_downsample(cin, cout, stride, b):
    return b == 0 and (stride != 1 or cin != cout)


def state_dict_shapes():
    """Ordered {key: shape}: encoder then decoder, each in its own state_dict order."""
    out = {}

    def bn(p, c):
        out[p + ".weight"] = (c,)
        out[p + ".bias"] = (c,)
        out[p + ".running_mean"] = (c,)
        out[p + ".running_var"] = (c,)
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
            if block_has_downsample(cin, cout, stride, b):
                out[p + ".downsample.0.weight"] = (cout, cin, 1, 1)
                bn(p + ".downsample.1", cout)
    out["encoder.fc.weight"] = (1000, 512)
    out["encoder.fc.bias"] = (1000,)
    for i in range(4, -1, -1):
        cin0 = NUM_CH_ENC[4] if i == 4 else NUM_CH_DEC[i + 1]
        cin1 = NUM_CH_DEC[i] + (NUM_CH_ENC[i - 1] if i > 0 else 0)
        for j, cin in ((0, cin0), (1, cin1)):
            out[upconv_key(i, j) + ".weight"] = (NUM_CH_DEC[i], cin, 3, 3)
            out[upconv_key(i, j) + ".bias"] = (NUM_CH_DEC[i],)
    for s in range(4):
        out[f"decoder.{10 + s}.conv.weight"] = (1, NUM_CH_DEC[s], 3, 3)
        out[f"decoder.{10 + s}.conv.bias"] = (1,)
    return out


def is_buffer(key):
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def dead_parameter(key):
    """Parameters Adam owns but that never receive a gradient: the classifier head (never called,
    resnet_encoder.py:87-98) and the scale 1..3 disparity heads (outputs unused, monodepth2_model.py:78)."""
    return key.startswith(("encoder.fc.", "decoder.11.", "decoder.12.", "decoder.13."))
