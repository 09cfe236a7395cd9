"""Architecture table of the mannequin-challenge hourglass (HourglassModel(3)).

Transcribes the layer configuration of
monodepth/mannequin_challenge/models/hourglass.py:58-181 (Channels1-4, HourglassModel) as data:
  ("inc", cin, cfg) with cfg = [[o0], [k, a, b] x 3]   -- `inception` (:19-55)
  ("pool",) / ("up",)                                   -- nn.AvgPool2d(2) / nn.UpsamplingBilinear2d(2)
  ("chan", branch0_ops, branch1_ops)                    -- ChannelsN: list[0](x) + list[1](x)
and the reference's state_dict key names, so checkpoints move both ways unchanged.
"""

_A = [[16], [3, 64, 16], [7, 64, 16], [11, 64, 16]]
_B = [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]]
_C = [[32], [3, 64, 32], [7, 64, 32], [11, 64, 32]]
_D = [[64], [3, 32, 64], [5, 32, 64], [7, 32, 64]]
_E = [[64], [3, 32, 64], [5, 32, 64], [7, 32, 64]]
_F = [[64], [3, 64, 64], [7, 64, 64], [11, 64, 64]]
_G = [[32], [3, 32, 32], [5, 32, 32], [7, 32, 32]]
_B2 = [[32], [3, 64, 32], [5, 64, 32], [7, 64, 32]]
_A2 = [[16], [3, 32, 16], [7, 32, 16], [11, 32, 16]]


def structure():
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


def state_dict_shapes():
    """Ordered {key: shape} identical to HourglassModel(3).state_dict() (781 entries)."""
    out = {}

    def conv(prefix, cin, cout, k):
        out[prefix + ".weight"] = (cout, cin, k, k)
        out[prefix + ".bias"] = (cout,)

    def bn(prefix, c, affine):
        if affine:
            out[prefix + ".weight"] = (c,)
            out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,)
        out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    def walk(node, prefix):
        if node[0] == "inc":
            _, cin, cfg = node
            conv(f"{prefix}.convs.0.0", cin, cfg[0][0], 1)
            bn(f"{prefix}.convs.0.1", cfg[0][0], False)
            for i in range(1, len(cfg)):
                k, a, b = cfg[i]
                conv(f"{prefix}.convs.{i}.0", cin, a, 1)
                bn(f"{prefix}.convs.{i}.1", a, False)
                conv(f"{prefix}.convs.{i}.3", a, b, k)
                bn(f"{prefix}.convs.{i}.4", b, False)
        elif node[0] == "chan":
            for bi, ops in enumerate(node[1:]):
                for oi, op in enumerate(ops):
                    walk(op, f"{prefix}.list.{bi}.{oi}")

    conv("seq.0", 3, 128, 7)
    bn("seq.1", 128, True)
    walk(structure(), "seq.3")
    conv("uncertainty_layer.0", 64, 1, 3)
    conv("pred_layer", 64, 1, 3)
    return out
