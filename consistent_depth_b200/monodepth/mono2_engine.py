"""Static execution plan of the monodepth2 network (ResNet-18 encoder + DepthDecoder) on the sm_100a conv engine.

Replaces, for a fixed (frames, H, W, feed size), `Monodepth2Model.estimate_depth` (monodepth/monodepth2_model.py:63-89),
`ResnetEncoder.forward` (monodepth2/networks/resnet_encoder.py:87-98), `DepthDecoder.forward`
(monodepth2/networks/depth_decoder.py:50-65) and their autograd backward:

  * every nn.Conv2d is the tcgen05 implicit-GEMM kernel of the hourglass engine (conv_tc.cu / conv_wgrad.cu);
    channel counts above 256 run as 256-wide chunks of the same kernel;
  * the three kinds of stride-2 convolution (conv1 7x7, layerN.0.conv1 3x3, layerN.0.downsample 1x1) are the stride-1
    kernel + a pick of pixels (2y,2x) (1x1: pick first, then convolve); their gradients are zero-stuffed back;
  * BatchNorm2d(train)+ReLU is normalise-on-load in the consumer; statistics come from the conv epilogue
    (cvd_conv_fwd_bn) or cvd_bn_stats after a stride-2 pick; the BasicBlock tail relu(bn2(.) + identity) is one pass;
  * ReflectionPad2d(1), ELU, nearest x2 upsampling and torch.cat of the decoder are ONE gather per source into the
    padded input buffer of the next 3x3 conv (cvd_gather_pad_fwd); the conv then runs "same" on the padded grid and
    only its interior is ever consumed; gradient buffers of those conv outputs keep an all-zero border;
  * bicubic resize in (+ input normalisation) and out (+ reciprocal), sigmoid: dedicated small kernels (mono2_ops.cu).

Parameters live in ONE flat fp32 buffer with a matching flat gradient buffer (+ loss tail slot), exactly like the
hourglass engine, so `FineTuneStep` (CUDA graph, fused Adam, one NCCL all-reduce) drives either model unchanged.
"""
import torch

from .. import ops
from . import mono2_arch as arch


def _numel(s):
    n = 1
    for v in s:
        n *= v
    return n


class Mono2Params:
    """Flat parameter / gradient / BN-buffer storage of a Monodepth2Model (same interface as McParams)."""

    def __init__(self, device, shapes=None, is_buffer=None, counter_key="encoder.bn1.num_batches_tracked"):
        self.dev = torch.device(device)
        self.sd_shapes = shapes if shapes is not None else arch.state_dict_shapes()
        is_buffer = is_buffer or arch.is_buffer
        self.counter_key = counter_key
        self.pmap, off = {}, 0
        self.bmap, boff = {}, 0
        for k, s in self.sd_shapes.items():
            if k.endswith("num_batches_tracked"):
                continue
            if is_buffer(k):
                self.bmap[k] = (boff, s)
                boff += _numel(s)
            else:
                off = (off + 3) // 4 * 4                  # 16-byte alignment of every tensor
                self.pmap[k] = (off, s)
                off += _numel(s)
        self.n_flat = (off + 3) // 4 * 4
        self.flat = torch.zeros(self.n_flat, device=self.dev)
        self.grad_store = torch.zeros(self.n_flat + 4, device=self.dev)     # [gradient | loss]: one all-reduce
        self.grad_flat = self.grad_store[:self.n_flat]
        self.loss_slot = self.grad_store[self.n_flat:self.n_flat + 1]
        self.buf_flat = torch.zeros(boff, device=self.dev)
        for k, (o, s) in self.bmap.items():
            if k.endswith("running_var"):
                self.buf_flat[o:o + s[0]] = 1.0
        self.num_batches_tracked = 0
        self.feed_size = None            # (height, width) entries of the stock encoder.pth (monodepth2_model.py:35-36)

    def _p(self, key, flat=None, n=None, shape=None):
        o, s = self.pmap[key]
        return (flat if flat is not None else self.flat)[o:o + (n or _numel(shape or s))].view(shape or s)

    def _g(self, key, shape=None):
        return self._p(key, self.grad_flat, shape=shape)

    def _rb(self, key, n=None):
        o, s = self.bmap[key]
        return self.buf_flat[o:o + (n or s[0])]

    def named_parameters(self):
        return [(k, self._p(k)) for k in self.pmap]

    def load_state_dict(self, sd):
        for k in self.pmap:
            self._p(k).copy_(torch.as_tensor(sd[k], dtype=torch.float32).reshape(self.pmap[k][1]))
        for k, (o, s) in self.bmap.items():
            self.buf_flat[o:o + s[0]].copy_(torch.as_tensor(sd[k], dtype=torch.float32))
        if self.counter_key in sd:
            self.num_batches_tracked = int(sd[self.counter_key])
        if "height" in sd and "width" in sd:
            self.feed_size = (int(sd["height"]), int(sd["width"]))

    def state_dict(self):
        out = {}
        for k in self.sd_shapes:
            if k in self.pmap:
                out[k] = self._p(k).detach().clone()
            elif k in self.bmap:
                o, s = self.bmap[k]
                out[k] = self.buf_flat[o:o + s[0]].detach().clone()
            else:
                out[k] = torch.tensor(self.num_batches_tracked, dtype=torch.long)
        return out


PAIR_WG = __import__("os").environ.get("CVD_PAIR_WG", "1") != "0"   # run a conv's wgrad and dgrad as parallel graph branches


class _BN:
    """One BatchNorm2d (affine): parameters, running statistics and the per-channel arrays the kernels exchange."""

    def __init__(self, eng, prefix, C):
        z = eng._z
        self.C = C
        self.a, self.b, self.rstd, self.mean, self.bw = z(C), z(C), z(C), z(C), z(C, 4)
        self.gamma, self.beta = eng._p(prefix + ".weight"), eng._p(prefix + ".bias")
        self.dgamma, self.dbeta = eng._g(prefix + ".weight"), eng._g(prefix + ".bias")
        self.rm, self.rv = eng._rb(prefix + ".running_mean"), eng._rb(prefix + ".running_var")
        # one statistics block per 64-channel chunk: chunked launches (grouped conv, Cout > 256) reduce concurrently
        self.scratch = ops.bn_scratch(eng.dev, max(256, 4 * C))
        self.fused = ops.make_bn(self.scratch, self.a, self.b, self.rstd, self.mean, self.gamma, self.beta, self.rm, self.rv)


class _Act:
    """A materialised activation (N,h,w,C) and the buffer of d loss / d activation."""

    def __init__(self, eng, h, w, C):
        self.h, self.w, self.C = h, w, C
        self.buf, self.dbuf = eng._z(eng.N, h, w, C), eng._z(eng.N, h, w, C)
        self.written = False

    def take_written(self):
        """-> accumulate flag for the next writer of dbuf; marks it written."""
        acc = self.written
        self.written = True
        return acc


def _plain(t):
    """cvd_src_t reading tensor t (N,h,w,C) as is."""
    return ops.make_src(ops.View(t, 0))


class Mono2Engine:
    def __init__(self, params, n_frames, H, W, feed, precision=3):
        fh, fw = feed
        assert fh % 32 == 0 and fw % 32 == 0, "monodepth2 feed size must be a multiple of 32 (five stride-2 stages)"
        self.P = params
        self.N, self.H, self.W, self.fh, self.fw = n_frames, H, W, fh, fw
        self.dev, self.prec = params.dev, precision
        self.train_mode = True
        self._p, self._g, self._rb = params._p, params._g, params._rb
        self._build()

    def _z(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, device=self.dev, dtype=dtype)

    # ------------------------------------------------------------------ plan helpers
    def _packed(self, cin, cout, k):
        return torch.empty(ops.packed_bytes(cin, cout, k, self.prec), dtype=torch.uint8, device=self.dev)

    def _conv(self, src, wkey, bias_key, dst, cin, cout, k, h, w, bn=None):
        """Forward conv src -> dst (tensor, plain).  bn: _BN whose train-mode statistics are fused into the epilogue
        (stride-2 convs take theirs from cvd_bn_stats after the pick instead: _bn_stats)."""
        Wt = self._p(wkey)
        bias = self._p(bias_key) if bias_key else None
        pk = self._packed(cin, cout, k)
        self.pack_fwd.append((Wt, pk, False))
        d = ops.make_dst(ops.View(dst, 0))
        N, prec = self.N, self.prec
        self.raw_outputs[wkey[:-7]] = dst
        self.fwd.append(lambda: ops.conv(src, pk, bias, d, N, h, w, cin, cout, k, prec, 0,
                                         bn=bn.fused if (bn is not None and self.train_mode) else None))
        if bn is not None:
            self._bn_eval(bn)

    def _bn_stats(self, bn, y):
        npix = y.numel() // y.shape[-1]

        def run():
            if self.train_mode:
                ops.bn_stats(y, 0, bn.C, npix, bn.scratch, bn.a, bn.b, bn.rstd, bn.mean, bn.gamma, bn.beta, bn.rm, bn.rv)
        self.fwd.append(run)
        self._bn_eval(bn)

    def _bn_eval(self, bn):
        def run():
            if not self.train_mode:               # eval(): running statistics (depth_fine_tuning.py:182)
                av = bn.gamma * torch.rsqrt(bn.rv + 1e-5)
                bn.a.copy_(av)
                bn.b.copy_(bn.beta - bn.rm * av)
        self.fwd.append(run)

    def _wgrad(self, gsrc, xsrc, wkey, cin, cout, k, h, w):
        dw, N, prec = self._g(wkey), self.N, self.prec
        f = lambda: ops.conv_wgrad(gsrc, xsrc, dw, N, h, w, cin, cout, k, prec)
        f.wgrad_of = gsrc
        self.bwd.append(f)

    def _dgrad(self, gsrc, wkey, dx, cin, cout, k, h, w, accumulate):
        """dx (tensor, N,h,w,cin) (+)= conv^T(g): a forward conv with GEMM-cin = cout and the flipped/transposed weights."""
        Wt = self._p(wkey)
        pk = self._packed(cout, cin, k)
        self.pack_bwd.append((Wt, pk, True))
        d = ops.make_dst(ops.View(dx, 0))
        N, prec, flags = self.N, self.prec, (ops.FLAG_ACCUM if accumulate else 0)
        f = lambda: ops.conv(gsrc, pk, None, d, N, h, w, cout, cin, k, prec, flags)
        last = self.bwd[-1] if self.bwd else None
        if PAIR_WG and getattr(last, "wgrad_of", None) is gsrc:
            # the weight gradient and the input gradient of one conv read the same gradient view and write different
            # tensors: two branches of the graph (on the small maps neither fills the GPU alone)
            self.bwd[-1] = ("par", [[f], [last]])
        else:
            self.bwd.append(f)

    def _bn_reduce(self, bn, y, dy, relu):
        npix = y.numel() // y.shape[-1]
        self.bwd.append(lambda: ops.bn_bwd_reduce(y, 0, bn.C, dy, npix, bn.scratch, bn.a, bn.b, bn.rstd, bn.mean, bn.bw, relu,
                                                  gamma=bn.gamma, beta=bn.beta, dgamma=bn.dgamma, dbeta=bn.dbeta))

    def _scratch_view(self, *shape):
        return self.scratch_full[:_numel(shape)].view(*shape)

    # ------------------------------------------------------------------ plan
    def _build(self):
        N, fh, fw, H, W = self.N, self.fh, self.fw, self.H, self.W
        z = self._z
        self.fwd, self.pack_fwd, self.pack_bwd, self.raw_outputs = [], [], [], {}
        self.img4 = z(N, fh, fw, 4)
        self.scratch_full = z(N * fh * fw * 64)           # stride-1 output of a stride-2 conv before the pick
        self.depth, self.disp0, self.ddisp0 = z(N, H, W), z(N, fh, fw), z(N, fh, fw)
        enc_bwd = []                                      # backward op groups of the encoder, in forward order

        # --- conv1 7x7/2 (3 -> 64) + bn1 + relu = feature 0 (resnet_encoder.py:89-92)
        h0, w0 = fh // 2, fw // 2
        f0 = _Act(self, h0, w0, 64)                       # buf = RAW conv output; read as relu(a x + b)
        bn0 = _BN(self, "encoder.bn1", 64)
        img = ops.make_src(ops.View(self.img4, 0))
        full0 = self._scratch_view(N, fh, fw, 64)
        self._conv(img, "encoder.conv1.weight", None, full0, 3, 64, 7, fh, fw)
        self.fwd.append(lambda: ops.subsample2(full0, f0.buf))
        self.raw_outputs["encoder.conv1"] = f0.buf        # (the stride-1 scratch is reused by later layers)
        self._bn_stats(bn0, f0.buf)
        G0 = z(N, fh, fw, 64)                             # zero-stuffed gradient of the stride-1 conv output
        self.f0, self.bn0 = f0, bn0

        def conv1_bwd():
            self._bn_reduce(bn0, f0.buf, f0.dbuf, True)
            self.bwd.append(lambda: ops.bnbwd_stuff(f0.buf, f0.dbuf, bn0.a, bn0.b, bn0.bw, True, G0, 2))
            self._wgrad(_plain(G0), img, "encoder.conv1.weight", 3, 64, 7, fh, fw)
        enc_bwd.append(conv1_bwd)

        # --- maxpool 3x3/2 (resnet_encoder.py:93)
        h1, w1 = h0 // 2, w0 // 2
        m = _Act(self, h1, w1, 64)
        amax = z(N, h1, w1, 64, dtype=torch.uint8)
        self.fwd.append(lambda: ops.maxpool_fwd(f0.buf, bn0.a, bn0.b, True, m.buf, amax))
        enc_bwd.append(lambda: self.bwd.append(
            (lambda acc: (lambda: ops.maxpool_bwd(m.dbuf, amax, f0.dbuf, acc)))(f0.take_written())))

        # --- layer1..4 (torchvision BasicBlock x 2 each)
        x, feats = m, [f0]
        for name, cin, cout, stride in arch.LAYERS:
            for b in range(2):
                x = self._block(x, cin if b == 0 else cout, cout, stride if b == 0 else 1, f"encoder.{name}.{b}", enc_bwd)
            feats.append(x)
        self.feats = feats

        # --- decoder (depth_decoder.py:50-65) + adapter tail (monodepth2_model.py:78-82)
        dec_bwd = []
        cur = dict(t=feats[4].buf, pad=0, mode=ops.GATHER_IDENTITY, C=512, h=feats[4].h, w=feats[4].w, act=feats[4], dR=None)
        for i in range(4, -1, -1):
            cur = self._up_level(i, cur, feats, dec_bwd)
        hp, wp = fh + 2, fw + 2
        assert cur["h"] == fh and cur["w"] == fw
        Pd, dPd = z(N, hp, wp, 16), z(N, hp, wp, 16)
        Rd, dRd = z(N, hp, wp, 4), z(N, hp, wp, 4)        # 1 real channel in a 4-channel pixel (128-bit pixel stride)
        self.fwd.append(lambda s=cur: ops.gather_pad_fwd(s["t"], 0, 1, None, None, Pd, 0, 16, 0, ops.GATHER_ELU))
        self._conv(_plain(Pd), "decoder.10.conv.weight", "decoder.10.conv.bias", Rd, 16, 1, 3, hp, wp)
        self.fwd.append(lambda: ops.sigmoid_fwd(Rd, self.disp0))
        self.fwd.append(lambda: ops.disp_to_depth(self.disp0, self.depth))

        # ---------------- backward plan (execution order)
        self.bwd = []
        self.bwd.append(lambda: ops.disp_to_depth_bwd(self.grad_depth, self.depth, self.ddisp0))
        self.bwd.append(lambda: ops.sigmoid_bwd(self.ddisp0, self.disp0, dRd))
        gb = self._g("decoder.10.conv.bias")
        self.bwd.append(lambda: ops.channel_sum(dRd, 0, 1, gb))
        self._wgrad(_plain(dRd), _plain(Pd), "decoder.10.conv.weight", 16, 1, 3, hp, wp)
        self._dgrad(_plain(dRd), "decoder.10.conv.weight", dPd, 16, 1, 3, hp, wp, False)
        self.bwd.append(lambda s=cur: ops.gather_pad_bwd(dPd, 0, s["t"], 0, 1, s["dR"], 0, 1, 16, 0, ops.GATHER_ELU, False))
        for emit in reversed(dec_bwd):
            emit()
        for emit in reversed(enc_bwd):
            emit()
        self.pack_fwd_tab = ops.make_pack_table(self.pack_fwd, self.dev)
        self.pack_bwd_tab = ops.make_pack_table(self.pack_bwd, self.dev)

    def _block(self, x, cin, cout, stride, p, enc_bwd):
        """torchvision BasicBlock: relu(bn2(conv2(relu(bn1(conv1(x))))) + (downsample(x) | x)); x, result: _Act."""
        N = self.N
        h, w = x.h, x.w
        hh, ww = h // stride, w // stride
        down = stride != 1 or cin != cout
        bn1, bn2 = _BN(self, p + ".bn1", cout), _BN(self, p + ".bn2", cout)
        y1, d1, y2 = self._z(N, hh, ww, cout), self._z(N, hh, ww, cout), self._z(N, hh, ww, cout)
        out = _Act(self, hh, ww, cout)
        xs = _plain(x.buf)
        if stride == 1:
            self._conv(xs, p + ".conv1.weight", None, y1, cin, cout, 3, h, w, bn=bn1)
            G = None
        else:
            full = self._scratch_view(N, h, w, cout)
            self._conv(xs, p + ".conv1.weight", None, full, cin, cout, 3, h, w)
            self.fwd.append(lambda: ops.subsample2(full, y1))
            self.raw_outputs[p + ".conv1"] = y1
            self._bn_stats(bn1, y1)
            G = self._z(N, h, w, cout)
        t1 = ops.make_src(ops.View(y1, 0), bn1.a, bn1.b, True)
        self._conv(t1, p + ".conv2.weight", None, y2, cout, cout, 3, hh, ww, bn=bn2)
        if down:
            bnd = _BN(self, p + ".downsample.1", cout)
            xsub, dxsub, yd = self._z(N, hh, ww, cin), self._z(N, hh, ww, cin), self._z(N, hh, ww, cout)
            if stride == 1:
                raise NotImplementedError("1x1 projection without stride does not occur in resnet18")
            self.fwd.append(lambda: ops.subsample2(x.buf, xsub))
            self._conv(_plain(xsub), p + ".downsample.0.weight", None, yd, cin, cout, 1, hh, ww, bn=bnd)
            self.fwd.append(lambda: ops.bn_add_relu(y2, bn2.a, bn2.b, yd, bnd.a, bnd.b, out.buf))
        else:
            self.fwd.append(lambda: ops.bn_add_relu(y2, bn2.a, bn2.b, x.buf, None, None, out.buf))

        def backward():
            B = self.bwd
            if down:
                B.append(lambda: ops.relu_bwd_add(out.dbuf, out.buf, None, False))
                self._bn_reduce(bnd, yd, out.dbuf, False)
                gd = ops.make_src(ops.View(yd, 0), bnd.a, bnd.b, False, dy=ops.View(out.dbuf, 0), bw=bnd.bw)
                self._wgrad(gd, _plain(xsub), p + ".downsample.0.weight", cin, cout, 1, hh, ww)
                self._dgrad(gd, p + ".downsample.0.weight", dxsub, cin, cout, 1, hh, ww, False)
            else:
                acc = x.take_written()
                B.append(lambda: ops.relu_bwd_add(out.dbuf, out.buf, x.dbuf, acc))
            self._bn_reduce(bn2, y2, out.dbuf, False)
            g2 = ops.make_src(ops.View(y2, 0), bn2.a, bn2.b, False, dy=ops.View(out.dbuf, 0), bw=bn2.bw)
            self._wgrad(g2, t1, p + ".conv2.weight", cout, cout, 3, hh, ww)
            self._dgrad(g2, p + ".conv2.weight", d1, cout, cout, 3, hh, ww, False)
            self._bn_reduce(bn1, y1, d1, True)
            if stride == 1:
                g1 = ops.make_src(ops.View(y1, 0), bn1.a, bn1.b, True, dy=ops.View(d1, 0), bw=bn1.bw)
            else:
                B.append(lambda: ops.bnbwd_stuff(y1, d1, bn1.a, bn1.b, bn1.bw, True, G, 2))
                g1 = _plain(G)
            self._wgrad(g1, xs, p + ".conv1.weight", cin, cout, 3, h, w)
            self._dgrad(g1, p + ".conv1.weight", x.dbuf, cin, cout, 3, h, w, x.take_written())
            if down:
                B.append(lambda: ops.stuff2(dxsub, x.dbuf, True))
        enc_bwd.append(backward)
        return out

    def _up_level(self, i, cur, feats, dec_bwd):
        """Decoder level i: ConvBlock(i,0) -> nearest x2 -> cat skip -> ConvBlock(i,1).  cur / result: description of
        the tensor whose ELU (or identity for the encoder feature) is the level's input."""
        N = self.N
        Cd, Cin0 = arch.NUM_CH_DEC[i], cur["C"]
        h, w = cur["h"], cur["w"]
        Cenc = arch.NUM_CH_ENC[i - 1] if i > 0 else 0
        k0, k1 = arch.upconv_key(i, 0), arch.upconv_key(i, 1)
        z = self._z
        P0, dP0 = z(N, h + 2, w + 2, Cin0), z(N, h + 2, w + 2, Cin0)
        R0, dR0 = z(N, h + 2, w + 2, Cd), z(N, h + 2, w + 2, Cd)
        H2, W2 = 2 * h + 2, 2 * w + 2
        P1, dP1 = z(N, H2, W2, Cd + Cenc), z(N, H2, W2, Cd + Cenc)
        R1, dR1 = z(N, H2, W2, Cd), z(N, H2, W2, Cd)
        self.fwd.append(lambda: ops.gather_pad_fwd(cur["t"], 0, cur["pad"], None, None, P0, 0, Cin0, 0, cur["mode"]))
        self._conv(_plain(P0), k0 + ".weight", k0 + ".bias", R0, Cin0, Cd, 3, h + 2, w + 2)
        self.fwd.append(lambda: ops.gather_pad_fwd(R0, 0, 1, None, None, P1, 0, Cd, 1, ops.GATHER_ELU))
        if i > 0:
            f = feats[i - 1]
            if i - 1 == 0:       # feature 0 is stored raw: relu(bn1(.)) on load
                fa, fb, fmode = self.bn0.a, self.bn0.b, ops.GATHER_AFFINE_RELU
            else:
                fa, fb, fmode = None, None, ops.GATHER_IDENTITY
            self.fwd.append(lambda: ops.gather_pad_fwd(f.buf, 0, 0, fa, fb, P1, Cd, Cenc, 0, fmode))
        self._conv(_plain(P1), k1 + ".weight", k1 + ".bias", R1, Cd + Cenc, Cd, 3, H2, W2)

        def backward():
            B = self.bwd
            gb1, gb0 = self._g(k1 + ".bias"), self._g(k0 + ".bias")
            B.append(lambda: ops.channel_sum(dR1, 0, Cd, gb1))
            self._wgrad(_plain(dR1), _plain(P1), k1 + ".weight", Cd + Cenc, Cd, 3, H2, W2)
            self._dgrad(_plain(dR1), k1 + ".weight", dP1, Cd + Cenc, Cd, 3, H2, W2, False)
            B.append(lambda: ops.gather_pad_bwd(dP1, 0, R0, 0, 1, dR0, 0, 1, Cd, 1, ops.GATHER_ELU, False))
            if i > 0:
                f = feats[i - 1]
                fmode = ops.GATHER_AFFINE_RELU if i - 1 == 0 else ops.GATHER_IDENTITY
                acc = f.take_written()
                B.append(lambda: ops.gather_pad_bwd(dP1, Cd, None, 0, 0, f.dbuf, 0, 0, Cenc, 0, fmode, acc))
            B.append(lambda: ops.channel_sum(dR0, 0, Cd, gb0))
            self._wgrad(_plain(dR0), _plain(P0), k0 + ".weight", Cin0, Cd, 3, h + 2, w + 2)
            self._dgrad(_plain(dR0), k0 + ".weight", dP0, Cin0, Cd, 3, h + 2, w + 2, False)
            if cur["dR"] is None:        # level 4: the input is encoder feature 4
                a4 = cur["act"]
                acc = a4.take_written()
                B.append(lambda: ops.gather_pad_bwd(dP0, 0, None, 0, 0, a4.dbuf, 0, 0, Cin0, 0, ops.GATHER_IDENTITY, acc))
            else:
                B.append(lambda: ops.gather_pad_bwd(dP0, 0, cur["t"], 0, 1, cur["dR"], 0, 1, Cin0, 0, ops.GATHER_ELU, False))
        dec_bwd.append(backward)
        return dict(t=R1, pad=1, mode=ops.GATHER_ELU, C=Cd, h=2 * h, w=2 * w, act=None, dR=dR1)

    # ------------------------------------------------------------------ execution
    def _run(self, plan):
        """Plan entries are callables (serial) or ("par", [branch, ...]): independent branches forked onto side streams and
        joined back (parallel branches of the captured CUDA graph), as in McEngine._run."""
        import os
        multi = os.environ.get("CVD_MULTI_STREAM", "1") == "1"
        main = torch.cuda.current_stream() if multi else None
        if not hasattr(self, "side_streams"):
            self.side_streams = []
        for op in plan:
            if not isinstance(op, tuple):
                op()
                continue
            branches = [b for b in op[1] if b]
            if not multi or len(branches) <= 1:
                for br in branches:
                    for f in br:
                        f()
                continue
            while len(self.side_streams) < len(branches) - 1:
                self.side_streams.append(torch.cuda.Stream(device=self.dev))
            for i, br in enumerate(branches[1:]):
                s = self.side_streams[i]
                s.wait_stream(main)
                with torch.cuda.stream(s):
                    for f in br:
                        f()
            for f in branches[0]:
                f()
            for i in range(len(branches) - 1):
                main.wait_stream(self.side_streams[i])

    def forward(self, images):
        """images (N,3,H,W) BGR in [0,1] (CUDA) -> depth (N,H,W) (engine-owned buffer)."""
        assert images.shape == (self.N, 3, self.H, self.W), images.shape
        ops.bicubic_image(images.contiguous(), self.img4)
        ops.pack_batch(self.pack_fwd_tab[0], self.pack_fwd_tab[1], self.prec)
        self._run(self.fwd)
        if self.train_mode:
            self.P.num_batches_tracked += 1
        return self.depth

    def backward(self, grad_depth):
        """grad_depth (N,H,W) = d loss / d depth; accumulates into the flat gradient buffer (zero it first)."""
        self.grad_depth = grad_depth.contiguous()
        ops.pack_batch(self.pack_bwd_tab[0], self.pack_bwd_tab[1], self.prec)
        self._run(self.bwd)
