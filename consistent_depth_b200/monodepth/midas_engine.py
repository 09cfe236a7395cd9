"""Static execution plan of MiDaS v2 (MidasNet) on the sm_100a conv engine.

Replaces, for a fixed (frames, H, W), `MidasV2Model.estimate_depth` (monodepth/midas_v2_model.py:52-69),
`MidasNet.forward` (monodepth/midas_v2/midas_net.py:49-76) with its ResNeXt-101 32x8d trunk (blocks.py:7-29) and
their autograd backward:

  * every nn.Conv2d is the tcgen05 implicit-GEMM kernel; channel counts above 256 run as 256-wide chunks;
  * the grouped 3x3 convs (groups = 32) run as 64-channel chunks of the same dense kernel with the weights expanded
    block-diagonally at packing time (layer4: one group per chunk, no waste; layer1: 8 groups per chunk);
    the weight gradient keeps only the same-group entries (cvd_conv_wgrad_grouped);
  * stride-2 convs: stride-1 kernel + pick of pixels (2y,2x) (1x1: pick first), gradients zero-stuffed back;
  * BatchNorm2d(train)+ReLU is normalise-on-load; statistics come from the conv epilogue or cvd_bn_stats after a pick;
    the Bottleneck tail relu(bn3(.) + identity) is one pass;
  * ResidualConvUnit: its in-place ReLU makes the skip relu(x) (blocks.py:111-117): out = relu(x) is written first
    and conv2 accumulates into it; the ReLUs in front of the convs are applied on load;
  * FeatureFusionBlock: bilinear x2 (align_corners=True) fused with the next block's sum (cvd_up2_bilinear_fwd).

Same flat parameter / gradient storage as the other engines: `FineTuneStep` drives it unchanged.
"""
import torch

from .. import ops
from . import midas_arch as arch
from .mono2_engine import Mono2Engine, Mono2Params, _Act, _BN, _plain

CHUNK = 64          # channel width of one block-diagonal launch of a grouped conv
CHUNK_LAUNCH = __import__("os").environ.get("CVD_MIDAS_CHUNK_LAUNCH", "1") != "0"   # grouped conv: all chunks in one launch
# CHUNK_LAUNCH off: the chunk launches of one grouped conv are independent and forked onto this many graph branches
BRANCHES = int(__import__("os").environ.get("CVD_MIDAS_BRANCHES", "8"))


class MidasParams(Mono2Params):
    def __init__(self, device):
        super().__init__(device, arch.state_dict_shapes(), arch.is_buffer, "pretrained.layer1.1.num_batches_tracked")


def _relu_src(t):
    """cvd_src_t reading relu(t) (ReLU applied on load)."""
    return ops.make_src(ops.View(t, 0), None, None, True)


class MidasEngine(Mono2Engine):
    def __init__(self, params, n_frames, H, W, precision=3):
        assert H % 32 == 0 and W % 32 == 0, "MiDaS needs H, W multiples of 32 (align = 32, midas_v2_model.py:14)"
        self.P = params
        self.N, self.H, self.W = n_frames, H, W
        self.dev, self.prec = params.dev, precision
        self.train_mode = True
        self._p, self._g, self._rb = params._p, params._g, params._rb
        self._build()

    # ------------------------------------------------------------------ helpers
    def _conv_b(self, src, wkey, bkey, dst, cin, cout, k, h, w, accumulate=False):
        """Forward conv with bias, optionally accumulating into dst."""
        Wt, bias = self._p(wkey), (self._p(bkey) if bkey else None)
        pk = self._packed(cin, cout, k)
        self.pack_fwd.append((Wt, pk, False))
        d = ops.make_dst(ops.View(dst, 0))
        N, prec, flags = self.N, self.prec, (ops.FLAG_ACCUM if accumulate else 0)
        self.raw_outputs[wkey[:-7]] = dst
        self.fwd.append(lambda: ops.conv(src, pk, bias, d, N, h, w, cin, cout, k, prec, flags))

    def _chunk_packs(self, Wt, width, gs, flip, table):
        """One contiguous packed buffer for all CHUNK x CHUNK block-diagonal chunks of a grouped weight."""
        nb = ops.packed_bytes(CHUNK, CHUNK, 3, self.prec)
        pk = torch.empty(nb * (width // CHUNK), dtype=torch.uint8, device=self.dev)
        for j in range(width // CHUNK):
            table.append((Wt[j * CHUNK:(j + 1) * CHUNK], pk[j * nb:(j + 1) * nb], flip, CHUNK, gs))
        return pk, nb

    def _grouped_fwd(self, y1, bn1, wkey, dst, width, h, w, bn2):
        """conv2 of a Bottleneck: grouped 3x3 of relu(bn1(y1)) -> dst, as width / CHUNK block-diagonal CHUNK-channel
        convs in ONE chunked launch (blockIdx.y = chunk; CVD_MIDAS_CHUNK_LAUNCH=0: one launch per chunk on BRANCHES streams)."""
        gs = width // arch.GROUPS
        Wt = self._p(wkey)
        N, prec = self.N, self.prec
        nch = width // CHUNK
        if CHUNK_LAUNCH:
            pk, nb = self._chunk_packs(Wt, width, gs, False, self.pack_fwd)
            s = ops.make_src(ops.View(y1, 0), bn1.a, bn1.b, True)
            d = ops.make_dst(ops.View(dst, 0))
            fused = bn2.fused if bn2 is not None else None
            self.fwd.append(lambda: ops.conv_chunks(s, pk, None, d, N, h, w, CHUNK, CHUNK, 3, nch, CHUNK, CHUNK, nb, prec, 0,
                                                    bn=fused if self.train_mode else None))
            if bn2 is not None:
                self._bn_eval(bn2)
            return
        branches = [[] for _ in range(BRANCHES)]
        scratches = [bn2.scratch] + [ops.bn_scratch(self.dev) for _ in range(BRANCHES - 1)] if bn2 is not None else None
        for j in range(nch):
            c0 = j * CHUNK
            pk = self._packed(CHUNK, CHUNK, 3)
            self.pack_fwd.append((Wt[c0:c0 + CHUNK], pk, False, CHUNK, gs))
            s = ops.make_src(ops.View(y1, c0), bn1.a, bn1.b, True)
            d = ops.make_dst(ops.View(dst, c0))
            fused = None
            if bn2 is not None:      # concurrent launches: one statistics scratch per branch
                fused = ops.make_bn(scratches[j % BRANCHES], bn2.a, bn2.b, bn2.rstd, bn2.mean, bn2.gamma[c0:c0 + CHUNK],
                                    bn2.beta[c0:c0 + CHUNK], bn2.rm[c0:c0 + CHUNK], bn2.rv[c0:c0 + CHUNK])
            branches[j % BRANCHES].append(lambda s=s, pk=pk, d=d, fused=fused: ops.conv(
                s, pk, None, d, N, h, w, CHUNK, CHUNK, 3, prec, 0, bn=fused if self.train_mode else None))
        self.fwd.append(("par", branches))
        if bn2 is not None:
            self._bn_eval(bn2)

    def _grouped_bwd(self, g_of_chunk, y1, bn1, wkey, d1, width, h, w):
        """Weight gradient and input gradient (-> d1) of the grouped conv; g_of_chunk(c0) -> cvd_src_t of the chunk."""
        gs = width // arch.GROUPS
        Wt, dW = self._p(wkey), self._g(wkey)
        N, prec = self.N, self.prec
        nch = width // CHUNK
        if CHUNK_LAUNCH:
            g0 = g_of_chunk(0)
            x0 = ops.make_src(ops.View(y1, 0), bn1.a, bn1.b, True)
            pk, nb = self._chunk_packs(Wt, width, gs, True, self.pack_bwd)
            d = ops.make_dst(ops.View(d1, 0))
            self.bwd.append(("par", [
                [lambda: ops.conv_wgrad_grouped_chunks(g0, x0, dW, N, h, w, CHUNK, nch, gs, 3, prec)],
                [lambda: ops.conv_chunks(g0, pk, None, d, N, h, w, CHUNK, CHUNK, 3, nch, CHUNK, CHUNK, nb, prec, 0)]]))
            return
        branches = [[] for _ in range(BRANCHES)]
        for j in range(nch):
            c0 = j * CHUNK
            g = g_of_chunk(c0)
            x = ops.make_src(ops.View(y1, c0), bn1.a, bn1.b, True)
            dw = dW[c0:c0 + CHUNK]
            branches[j % BRANCHES].append(lambda g=g, x=x, dw=dw: ops.conv_wgrad_grouped(g, x, dw, N, h, w, CHUNK, gs, 3, prec))
            pk = self._packed(CHUNK, CHUNK, 3)
            self.pack_bwd.append((Wt[c0:c0 + CHUNK], pk, True, CHUNK, gs))
            d = ops.make_dst(ops.View(d1, c0))
            branches[(j + BRANCHES // 2) % BRANCHES].append(lambda g=g, pk=pk, d=d: ops.conv(g, pk, None, d, N, h, w, CHUNK, CHUNK, 3, prec, 0))
        self.bwd.append(("par", branches))

    # ------------------------------------------------------------------ plan
    def _build(self):
        N, H, W = self.N, self.H, self.W
        z = self._z
        self.fwd, self.pack_fwd, self.pack_bwd, self.raw_outputs = [], [], [], {}
        self.img4 = z(N, H, W, 4)
        # stride-1 output of a stride-2 conv before the pick: conv1 (64 ch at H x W) is the largest
        self.scratch_full = z(N * H * W * 64)
        self.depth = z(N, H, W)
        enc_bwd = []

        # --- trunk stem: conv1 7x7/2 + bn1 + relu + maxpool 3x3/2 (blocks.py:15-17)
        h0, w0 = H // 2, W // 2
        f0 = _Act(self, h0, w0, 64)                       # RAW conv output, read as relu(a x + b)
        bn0 = _BN(self, "pretrained.layer1.1", 64)
        img = ops.make_src(ops.View(self.img4, 0))
        full0 = self._scratch_view(N, H, W, 64)
        self._conv(img, "pretrained.layer1.0.weight", None, full0, 3, 64, 7, H, W)
        self.fwd.append(lambda: ops.subsample2(full0, f0.buf))
        self.raw_outputs["pretrained.layer1.0"] = f0.buf
        self._bn_stats(bn0, f0.buf)
        G0 = z(N, H, W, 64)

        def conv1_bwd():
            self._bn_reduce(bn0, f0.buf, f0.dbuf, True)
            self.bwd.append(lambda: ops.bnbwd_stuff(f0.buf, f0.dbuf, bn0.a, bn0.b, bn0.bw, True, G0, 2))
            self._wgrad(_plain(G0), img, "pretrained.layer1.0.weight", 3, 64, 7, H, W)
        enc_bwd.append(conv1_bwd)
        h1, w1 = h0 // 2, w0 // 2
        m = _Act(self, h1, w1, 64)
        amax = z(N, h1, w1, 64, dtype=torch.uint8)
        self.fwd.append(lambda: ops.maxpool_fwd(f0.buf, bn0.a, bn0.b, True, m.buf, amax))
        enc_bwd.append(lambda: self.bwd.append(
            (lambda acc: (lambda: ops.maxpool_bwd(m.dbuf, amax, f0.dbuf, acc)))(f0.take_written())))

        # --- layer1..4: Bottleneck blocks
        x, taps = m, []
        for name, inplanes, planes, blocks, stride in arch.STAGES:
            width = planes * 4
            for b in range(blocks):
                x = self._bottleneck(x, inplanes if b == 0 else width, width, stride if b == 0 else 1, b == 0,
                                     arch.block_prefix(name, b), enc_bwd)
            taps.append(x)
        self.taps = taps

        # --- decoder
        dec = self._decoder(taps)

        # ---------------- backward plan (execution order)
        self.bwd = []
        dec()
        for emit in reversed(enc_bwd):
            emit()
        self.pack_fwd_tab = ops.make_pack_table(self.pack_fwd, self.dev)
        self.pack_bwd_tab = ops.make_pack_table(self.pack_bwd, self.dev)

    def _bottleneck(self, x, cin, width, stride, down, p, enc_bwd):
        """torchvision Bottleneck (v1.5, groups=32): relu(bn3(conv3(relu(bn2(conv2g(relu(bn1(conv1(x)))))))) + id)."""
        N = self.N
        h, w = x.h, x.w
        hh, ww = h // stride, w // stride
        z = self._z
        bn1, bn2, bn3 = _BN(self, p + ".bn1", width), _BN(self, p + ".bn2", width), _BN(self, p + ".bn3", width)
        y1, d1 = z(N, h, w, width), z(N, h, w, width)
        y2, d2 = z(N, hh, ww, width), z(N, hh, ww, width)
        y3 = z(N, hh, ww, width)
        out = _Act(self, hh, ww, width)
        xs = _plain(x.buf)
        self._conv(xs, p + ".conv1.weight", None, y1, cin, width, 1, h, w, bn=bn1)
        if stride == 1:
            self._grouped_fwd(y1, bn1, p + ".conv2.weight", y2, width, h, w, bn2)
            G = None
        else:
            full = self._scratch_view(N, h, w, width)
            self._grouped_fwd(y1, bn1, p + ".conv2.weight", full, width, h, w, None)
            self.fwd.append(lambda: ops.subsample2(full, y2))
            self._bn_stats(bn2, y2)
            G = z(N, h, w, width)
        self.raw_outputs[p + ".conv2"] = y2
        t2 = ops.make_src(ops.View(y2, 0), bn2.a, bn2.b, True)
        self._conv(t2, p + ".conv3.weight", None, y3, width, width, 1, hh, ww, bn=bn3)
        if down:
            bnd = _BN(self, p + ".downsample.1", width)
            yd = z(N, hh, ww, width)
            if stride == 1:
                xsub, dxsub = x.buf, None
            else:
                xsub, dxsub = z(N, hh, ww, cin), z(N, hh, ww, cin)
                self.fwd.append(lambda: ops.subsample2(x.buf, xsub))
            self._conv(_plain(xsub), p + ".downsample.0.weight", None, yd, cin, width, 1, hh, ww, bn=bnd)
            self.fwd.append(lambda: ops.bn_add_relu(y3, bn3.a, bn3.b, yd, bnd.a, bnd.b, out.buf))
        else:
            self.fwd.append(lambda: ops.bn_add_relu(y3, bn3.a, bn3.b, x.buf, None, None, out.buf))
        self.raw_outputs[p] = out.buf

        def backward():
            B = self.bwd
            if down:
                B.append(lambda: ops.relu_bwd_add(out.dbuf, out.buf, None, False))
                self._bn_reduce(bnd, yd, out.dbuf, False)
                gd = ops.make_src(ops.View(yd, 0), bnd.a, bnd.b, False, dy=ops.View(out.dbuf, 0), bw=bnd.bw)
                self._wgrad(gd, _plain(xsub), p + ".downsample.0.weight", cin, width, 1, hh, ww)
                if stride == 1:
                    self._dgrad(gd, p + ".downsample.0.weight", x.dbuf, cin, width, 1, hh, ww, x.take_written())
                else:
                    self._dgrad(gd, p + ".downsample.0.weight", dxsub, cin, width, 1, hh, ww, False)
            else:
                acc = x.take_written()
                B.append(lambda: ops.relu_bwd_add(out.dbuf, out.buf, x.dbuf, acc))
            self._bn_reduce(bn3, y3, out.dbuf, False)
            g3 = ops.make_src(ops.View(y3, 0), bn3.a, bn3.b, False, dy=ops.View(out.dbuf, 0), bw=bn3.bw)
            self._wgrad(g3, t2, p + ".conv3.weight", width, width, 1, hh, ww)
            self._dgrad(g3, p + ".conv3.weight", d2, width, width, 1, hh, ww, False)
            self._bn_reduce(bn2, y2, d2, True)
            if stride == 1:
                gch = lambda c0: ops.make_src(ops.View(y2, c0), bn2.a, bn2.b, True, dy=ops.View(d2, c0), bw=bn2.bw)
            else:
                B.append(lambda: ops.bnbwd_stuff(y2, d2, bn2.a, bn2.b, bn2.bw, True, G, 2))
                gch = lambda c0: ops.make_src(ops.View(G, c0))
            self._grouped_bwd(gch, y1, bn1, p + ".conv2.weight", d1, width, h, w)
            self._bn_reduce(bn1, y1, d1, True)
            g1 = ops.make_src(ops.View(y1, 0), bn1.a, bn1.b, True, dy=ops.View(d1, 0), bw=bn1.bw)
            self._wgrad(g1, xs, p + ".conv1.weight", cin, width, 1, h, w)
            self._dgrad(g1, p + ".conv1.weight", x.dbuf, cin, width, 1, h, w, x.take_written())
            if down and stride != 1:
                B.append(lambda: ops.stuff2(dxsub, x.dbuf, True))
        enc_bwd.append(backward)
        return out

    def _rcu(self, p, S, h, w, U_init):
        """ResidualConvUnit p on input buffer S -> U (blocks.py:83-117).  U_init() must have written the skip term
        (relu(S), possibly plus another addend) into U before conv2 accumulates.  Returns (o1, D1)."""
        F = arch.FEATURES
        o1, D1 = self._z(self.N, h, w, F), self._z(self.N, h, w, F)
        self._conv_b(_relu_src(S), p + ".conv1.weight", p + ".conv1.bias", o1, F, F, 3, h, w)
        U = U_init()
        self._conv_b(_relu_src(o1), p + ".conv2.weight", p + ".conv2.bias", U, F, F, 3, h, w, accumulate=True)
        self.raw_outputs[p] = U
        return o1, D1

    def _rcu_bwd(self, p, S, dS_target, dU, o1, D1, h, w, before_input_grad=None):
        """Backward of ResidualConvUnit p: dU = d loss / d output (clobbered); dS_target = d loss / d input (written)."""
        F = arch.FEATURES
        B = self.bwd
        gb2, gb1 = self._g(p + ".conv2.bias"), self._g(p + ".conv1.bias")
        B.append(lambda: ops.channel_sum(dU, 0, F, gb2))
        self._wgrad(_plain(dU), _relu_src(o1), p + ".conv2.weight", F, F, 3, h, w)
        self._dgrad(_plain(dU), p + ".conv2.weight", D1, F, F, 3, h, w, False)
        B.append(lambda: ops.relu_bwd_add(D1, o1, None, False))            # d o1 = d relu(o1) * [o1 > 0]
        B.append(lambda: ops.channel_sum(D1, 0, F, gb1))
        self._wgrad(_plain(D1), _relu_src(S), p + ".conv1.weight", F, F, 3, h, w)
        self._dgrad(_plain(D1), p + ".conv1.weight", dU, F, F, 3, h, w, True)   # dU <- d relu(S) = skip + conv1 path
        B.append(lambda: ops.relu_bwd_add(dU, S, dS_target, False))         # d S = d relu(S) * [S > 0]

    def _decoder(self, taps):
        """layerK_rn -> refinenet4..1 -> output_conv -> 1 / relu (midas_net.py:61-76, midas_v2_model.py:67).
        Returns the function that emits the decoder's backward ops."""
        N, H, W, F = self.N, self.H, self.W, arch.FEATURES
        z = self._z
        L, dL = [], []
        for i, t in enumerate(taps):
            Lk, dLk = z(N, t.h, t.w, F), z(N, t.h, t.w, F)
            self._conv_b(_plain(t.buf), f"scratch.layer{i + 1}_rn.weight", None, Lk, t.C, F, 3, t.h, t.w)
            L.append(Lk); dL.append(dLk)
        levels = []          # per refinenet r: dict of buffers
        U_prev = None
        for r in (4, 3, 2, 1):
            t = taps[r - 1]
            h, w = t.h, t.w
            p = f"scratch.refinenet{r}"
            lv = dict(r=r, h=h, w=w, p=p)
            if r == 4:
                S, dS = L[3], dL[3]
            else:
                S, dS = z(N, h, w, F), z(N, h, w, F)
                Lr, Up = L[r - 1], U_prev

                def init_S(S=S, Lr=Lr, Up=Up):
                    self.fwd.append(lambda: ops.up2_bilinear(Up, Lr, True, S, True))      # up2(U_{r+1}) + relu(L_r)
                    return S
                lv["rcu1"] = self._rcu(p + ".resConfUnit1", Lr, h, w, init_S)
            U, dU = z(N, h, w, F), z(N, h, w, F)

            def init_U(S=S, U=U):
                self.fwd.append(lambda: ops.relu_add(S, None, U))
                return U
            lv["rcu2"] = self._rcu(p + ".resConfUnit2", S, h, w, init_U)
            lv.update(S=S, dS=dS, U=U, dU=dU)
            levels.append(lv)
            U_prev = U
        h2, w2 = H // 2, W // 2
        Pth, dPth = z(N, h2, w2, F), z(N, h2, w2, F)
        U1 = levels[-1]["U"]
        self.fwd.append(lambda: ops.up2_bilinear(U1, None, False, Pth, True))
        hd1, dhd1 = z(N, h2, w2, 128), z(N, h2, w2, 128)
        hd1u, dhd1u = z(N, H, W, 128), z(N, H, W, 128)
        hd2, dhd2 = z(N, H, W, 32), z(N, H, W, 32)
        hd3, dhd3 = z(N, H, W, 4), z(N, H, W, 4)
        oc = "scratch.output_conv"
        self._conv_b(_plain(Pth), oc + ".0.weight", oc + ".0.bias", hd1, F, 128, 3, h2, w2)
        self.fwd.append(lambda: ops.up2_bilinear(hd1, None, False, hd1u, False))
        self._conv_b(_plain(hd1u), oc + ".2.weight", oc + ".2.bias", hd2, 128, 32, 3, H, W)
        self._conv_b(_relu_src(hd2), oc + ".4.weight", oc + ".4.bias", hd3, 32, 1, 1, H, W)
        self.fwd.append(lambda: ops.recip_relu(hd3, self.depth))

        def backward():
            B = self.bwd
            B.append(lambda: ops.recip_relu_bwd(self.grad_depth, self.depth, hd3, dhd3))
            g4, g2, g0 = self._g(oc + ".4.bias"), self._g(oc + ".2.bias"), self._g(oc + ".0.bias")
            B.append(lambda: ops.channel_sum(dhd3, 0, 1, g4))
            self._wgrad(_plain(dhd3), _relu_src(hd2), oc + ".4.weight", 32, 1, 1, H, W)
            self._dgrad(_plain(dhd3), oc + ".4.weight", dhd2, 32, 1, 1, H, W, False)
            B.append(lambda: ops.relu_bwd_add(dhd2, hd2, None, False))
            B.append(lambda: ops.channel_sum(dhd2, 0, 32, g2))
            self._wgrad(_plain(dhd2), _plain(hd1u), oc + ".2.weight", 128, 32, 3, H, W)
            self._dgrad(_plain(dhd2), oc + ".2.weight", dhd1u, 128, 32, 3, H, W, False)
            B.append(lambda: ops.up2_bilinear_bwd(dhd1u, dhd1, False, False))
            B.append(lambda: ops.channel_sum(dhd1, 0, 128, g0))
            self._wgrad(_plain(dhd1), _plain(Pth), oc + ".0.weight", F, 128, 3, h2, w2)
            self._dgrad(_plain(dhd1), oc + ".0.weight", dPth, F, 128, 3, h2, w2, False)
            dU1 = levels[-1]["dU"]
            B.append(lambda: ops.up2_bilinear_bwd(dPth, dU1, True, False))
            for idx in range(len(levels) - 1, -1, -1):          # refinenet1, 2, 3, 4
                lv = levels[idx]
                r, h, w, p = lv["r"], lv["h"], lv["w"], lv["p"]
                o1, D1 = lv["rcu2"]
                self._rcu_bwd(p + ".resConfUnit2", lv["S"], lv["dS"], lv["dU"], o1, D1, h, w)
                if r < 4:
                    up = levels[idx - 1]                          # refinenet r+1 feeds this level
                    dS, dUp = lv["dS"], up["dU"]
                    B.append(lambda dS=dS, dUp=dUp: ops.up2_bilinear_bwd(dS, dUp, True, False))   # before dS is clobbered
                    o1b, D1b = lv["rcu1"]
                    self._rcu_bwd(p + ".resConfUnit1", L[r - 1], dL[r - 1], lv["dS"], o1b, D1b, h, w)
                t = taps[r - 1]
                wk = f"scratch.layer{r}_rn.weight"
                self._wgrad(_plain(dL[r - 1]), _plain(t.buf), wk, t.C, F, 3, h, w)
                self._dgrad(_plain(dL[r - 1]), wk, t.dbuf, t.C, F, 3, h, w, t.take_written())
        return backward

    # ------------------------------------------------------------------ execution
    def forward(self, images):
        """images (N,3,H,W) BGR in [0,1] (CUDA) -> depth (N,H,W) (engine-owned buffer)."""
        assert images.shape == (self.N, 3, self.H, self.W), images.shape
        ops.image_normalize(images.contiguous(), self.img4, arch.NORM_MEAN, arch.NORM_STD)
        ops.pack_batch(self.pack_fwd_tab[0], self.pack_fwd_tab[1], self.prec)
        self._run(self.fwd)
        if self.train_mode:
            self.P.num_batches_tracked += 1
        return self.depth
