"""Static execution plan of the mannequin-challenge hourglass on the sm_100a conv engine.

Replaces `HourglassModel.forward` (monodepth/mannequin_challenge/models/hourglass.py:175-181) and
its autograd backward for a fixed (frames, H, W): every nn.Conv2d becomes one tcgen05 implicit-GEMM
launch (the four 1x1 convs of an `inception` block (:27,39) are ONE GEMM with N = o0+a1+a2+a3),
BatchNorm2d(train)+ReLU never materialises (statistics kernel after the conv, normalise-on-load in the
consumer), torch.cat is a channel view, and AvgPool / UpsamplingBilinear2d+add are single fused passes.

Buffers (all NHWC fp32, allocated once):
  per inception: raw conv outputs  buf[N,h,w, o0 | a1 a2 a3 | b1 b2 b3]   (block output = view with a gap)
                 gradient buffer   dbuf of the same layout (d loss / d post-activation)
                 per-channel arrays a, b (BN scale/shift), rstd, mean, bw (backward constants)
  parameters:    ONE flat fp32 buffer (+ matching flat gradient buffer) so Adam is one launch and the
                 multi-GPU gradient exchange is one NCCL all-reduce; the 1x1 weights of a block are
                 adjacent in it, which is what makes the fused GEMM's weight matrix contiguous.
The plan is a list of pre-bound C-ABI calls; the whole training step replays inside one CUDA graph.
"""
import torch

from .. import ops
from . import mc_arch


class _T:
    """Activation handle: a channel view of a buffer + how to read it (BN scale/shift, ReLU) + its gradient view."""

    def __init__(self, buf, off=0, n0=0, gap=0, C=None, a=None, b=None, relu=False, dbuf=None):
        self.buf, self.off, self.n0, self.gap = buf, off, n0, gap
        self.C = C if C is not None else buf.shape[-1]
        self.a, self.b, self.relu, self.dbuf = a, b, relu, dbuf
        self.grad_written = False
        self.rstd = self.mean = self.bw = None

    def view(self):
        return ops.View(self.buf, self.off, self.n0, self.gap)

    def dview(self):
        return ops.View(self.dbuf, self.off, self.n0, self.gap)

    def src(self):
        return ops.make_src(self.view(), self.a, self.b, self.relu)

    def bnbwd_src(self):
        return ops.make_src(self.view(), self.a, self.b, self.relu, dy=self.dview(), bw=self.bw)


class McParams:
    """Flat parameter / gradient / BN-buffer storage shared by every engine (one per input shape) of a model."""

    def __init__(self, device):
        self.dev = torch.device(device)
        self._layout_params()

    def _layout_params(self):
        shapes = mc_arch.state_dict_shapes()
        self.sd_shapes = shapes
        order = []          # (key, shape) in flat order
        bufs = []           # running-stat keys in flat order

        def add_conv1():
            order.extend([("seq.0.weight", shapes["seq.0.weight"]), ("seq.0.bias", (128,)),
                          ("seq.1.weight", (128,)), ("seq.1.bias", (128,))])
            bufs.extend([("seq.1.running_mean", (128,)), ("seq.1.running_var", (128,))])

        def walk(node, prefix):
            if node[0] == "inc":
                cfg = node[2]
                for i in range(4):
                    order.append((f"{prefix}.convs.{i}.0.weight", shapes[f"{prefix}.convs.{i}.0.weight"]))
                for i in range(4):
                    order.append((f"{prefix}.convs.{i}.0.bias", shapes[f"{prefix}.convs.{i}.0.bias"]))
                for i in range(1, 4):
                    order.append((f"{prefix}.convs.{i}.3.weight", shapes[f"{prefix}.convs.{i}.3.weight"]))
                for i in range(1, 4):
                    order.append((f"{prefix}.convs.{i}.3.bias", shapes[f"{prefix}.convs.{i}.3.bias"]))
                for stat in ("running_mean", "running_var"):
                    for i in range(4):
                        bufs.append((f"{prefix}.convs.{i}.1.{stat}", shapes[f"{prefix}.convs.{i}.1.{stat}"]))
                for stat in ("running_mean", "running_var"):
                    for i in range(1, 4):
                        bufs.append((f"{prefix}.convs.{i}.4.{stat}", shapes[f"{prefix}.convs.{i}.4.{stat}"]))
            elif node[0] == "chan":
                for bi, branch in enumerate(node[1:]):
                    for oi, op in enumerate(branch):
                        walk(op, f"{prefix}.list.{bi}.{oi}")

        add_conv1()
        walk(mc_arch.structure(), "seq.3")
        order.extend([("pred_layer.weight", (1, 64, 3, 3)), ("pred_layer.bias", (1,))])

        def numel(s):
            n = 1
            for v in s:
                n *= v
            return n

        self.pmap, off = {}, 0
        for k, s in order:
            # every tensor starts 16-byte aligned, EXCEPT that members of a fused group must be adjacent:
            # all group members have sizes that are multiples of 4 floats, so alignment never inserts a gap there
            off = (off + 3) // 4 * 4
            self.pmap[k] = (off, s)
            off += numel(s)
        self.n_flat = (off + 3) // 4 * 4
        self.flat = torch.zeros(self.n_flat, device=self.dev)
        # flat gradient + one tail slot for the (local) loss: a single all-reduce carries both
        self.grad_store = torch.zeros(self.n_flat + 4, device=self.dev)
        self.grad_flat = self.grad_store[:self.n_flat]
        self.loss_slot = self.grad_store[self.n_flat:self.n_flat + 1]
        self.bmap, boff = {}, 0
        for k, s in bufs:
            self.bmap[k] = (boff, s)
            boff += numel(s)
        self.buf_flat = torch.zeros(boff, device=self.dev)
        for k, (o, s) in self.bmap.items():
            if k.endswith("running_var"):
                self.buf_flat[o:o + numel(s)] = 1.0
        self.uncertainty = {"uncertainty_layer.0.weight": torch.zeros(1, 64, 3, 3, device=self.dev),
                            "uncertainty_layer.0.bias": torch.zeros(1, device=self.dev)}
        self.num_batches_tracked = 0

    def _p(self, key, flat=None, n=None, shape=None):
        o, s = self.pmap[key]
        cnt = 1
        for v in (shape or s):
            cnt *= v
        return (flat if flat is not None else self.flat)[o:o + (n or cnt)].view(shape or s)

    def _g(self, key, shape=None):
        return self._p(key, self.grad_flat, shape=shape)

    def _rb(self, key, n=None):
        o, s = self.bmap[key]
        return self.buf_flat[o:o + (n or s[0])]

    def named_parameters(self):
        """Leaf views into the flat buffer, keyed like the reference; uncertainty_layer included (never trained)."""
        out = []
        for k, (o, s) in self.pmap.items():
            out.append((k, self._p(k)))
        return out

    def load_state_dict(self, sd):
        for k in self.pmap:
            self._p(k).copy_(torch.as_tensor(sd[k], dtype=torch.float32).reshape(self.pmap[k][1]))
        for k in self.bmap:
            o, s = self.bmap[k]
            self.buf_flat[o:o + s[0]].copy_(torch.as_tensor(sd[k], dtype=torch.float32))
        for k in self.uncertainty:
            if k in sd:
                self.uncertainty[k].copy_(torch.as_tensor(sd[k], dtype=torch.float32))
        if "seq.1.num_batches_tracked" in sd:
            self.num_batches_tracked = int(sd["seq.1.num_batches_tracked"])

    def state_dict(self):
        out = {}
        for k, s in self.sd_shapes.items():
            if k in self.pmap:
                out[k] = self._p(k).detach().clone()
            elif k in self.bmap:
                o, sh = self.bmap[k]
                out[k] = self.buf_flat[o:o + sh[0]].detach().clone()
            elif k in self.uncertainty:
                out[k] = self.uncertainty[k].detach().clone()
            else:
                out[k] = torch.tensor(self.num_batches_tracked, dtype=torch.long)
        return out


class McEngine:
    def __init__(self, params, n_frames, H, W, precision=3):
        assert H % 16 == 0 and W % 16 == 0, "mannequin-challenge hourglass needs H, W multiples of 16 (align = 16)"
        self.P = params
        self.N, self.H, self.W, self.dev, self.prec = n_frames, H, W, params.dev, precision
        self.train_mode = True
        import os
        self.multi_stream = os.environ.get("CVD_MULTI_STREAM", "1") == "1"
        self.fuse_bn = os.environ.get("CVD_FUSE_BN", "1") == "1"     # BN batch statistics in the conv epilogue
        # fork the side streams BEFORE branch 0 is enqueued so that branch 0 overlaps with the others too
        # (measured on B200, round 2: 139.0 -> 144.8 frame-pairs/s); CVD_FORK_FIRST=0 restores the old order
        self.fork_first = os.environ.get("CVD_FORK_FIRST", "1") == "1"
        # second-generation conv path for the inception convolutions (prep.cu + conv2.cu: operands pre-split once into bf16
        # hi/lo planes, TMA-fed kx-fused tcgen05 conv for forward and dgrad); CVD_CONV2=0 restores the first-generation kernels.
        # bf16x3 parity mode only.
        self.v2 = os.environ.get("CVD_CONV2", "1") == "1" and precision == 3
        self.v2_wgrad = self.v2 and os.environ.get("CVD_WGRAD2", "1") == "1"       # weight gradients on the same operand planes
        # Per-layer dispatch (tools/conv2_microbench.py, B200): the kx-fused kernel wins where the GEMM N of the per-tap
        # kernel is small (<= 32 output channels: 2.2x on 64->16 11x11) and for 1x1 convs; with N >= 64 the per-tap kernel
        # already runs the tensor pipe at ~76 % and has no idle window lanes, so those k x k convs stay on it.
        self.v2_nmax = int(os.environ.get("CVD_CONV2_NMAX", "32"))
        self.v2_all = os.environ.get("CVD_CONV2_ALL", "0") == "1"                  # A/B: every inception conv on the new kernels
        # Weight gradients feed nothing but Adam, so they leave the backward's critical path: every wgrad is enqueued on
        # one of two side streams as soon as its operands exist and only joined at the end of backward(); inside the CUDA
        # graph they overlap with the dgrad / BatchNorm-backward chain, in particular with the low-resolution levels whose
        # kernels fill a fraction of the SMs.  Needs per-block gradient-operand planes (no shared scratch).
        self.wg_async = os.environ.get("CVD_WGRAD_ASYNC", "1") == "1"
        self.wg_streams = []
        self.side_streams = []
        self.pmap, self.grad_flat = params.pmap, params.grad_flat
        self._p, self._g, self._rb = params._p, params._g, params._rb
        self._build_plan()

    # ------------------------------------------------------------------ plan
    def _zeros(self, *shape):
        return torch.zeros(*shape, device=self.dev)

    def _packed(self, cin, cout, k):
        return torch.empty(ops.packed_bytes(cin, cout, k, self.prec), dtype=torch.uint8, device=self.dev)

    # second-generation path helpers ------------------------------------------
    def _z_of(self, t, h, w):
        """Operand planes of activation handle t (post BN+ReLU), prepared ONCE per forward however many convs read it."""
        if getattr(t, "z", None) is None:
            t.z = ops.z_alloc(self.N, t.C, h, w, self.dev)
            src, C, z = t.src(), t.C, t.z
            self.fwd.append(lambda: ops.prep_operand(src, C, z))
        return t.z

    def _conv2(self, z, zoff, wkey, bkey, dst, cin, cout, k, h, w, wshape=None, bn=None):
        N = self.N
        Wt = self._p(wkey, shape=wshape)
        bias = self._p(bkey, n=cout, shape=(cout,))
        pk = torch.empty(ops.conv2_packed_bytes(cin, cout, k), dtype=torch.uint8, device=self.dev)
        self.pack2_fwd.append((Wt, pk, False))
        self.raw_outputs[wkey[:-7]] = (dst.buf, dst.off, cout)
        d = ops.make_dst(dst.view())
        t, rm, rv, gamma, beta, si = bn
        bns = ops.make_bn(self.conv_scratch[si], t.a, t.b, t.rstd, t.mean, gamma, beta, rm, rv)
        self.fwd.append(lambda: ops.conv2(z, zoff, pk, bias, d, N, h, w, cin, cout, k, 0, bns if self.train_mode else None))

    def _use_conv2(self, k_gemm, n_gemm, k):
        """Forward / dgrad dispatch of a k x k conv with GEMM K = k_gemm input and N = n_gemm output channels (measured,
        tools/conv2_microbench.py on B200): few output channels or wide filters -> kx-fused kernel."""
        if self.v2_all:
            return True
        return n_gemm <= 16 or (n_gemm <= self.v2_nmax and k >= 7) or (k >= 11 and k_gemm >= 64 and n_gemm <= k_gemm)

    def _use_wgrad2(self, cin, cout, k):
        """Weight-gradient dispatch: the TMA-fed kernel everywhere except more output than input channels (32 -> 64: 8 G
        chunks x 3..7 ky rows make many passes over small maps)."""
        return self.v2_all or not (cin < cout)

    def _gz_planes(self, C, h, w, slot):
        if self.wg_async:                # read by deferred weight gradients: one set of planes per block
            return ops.z_alloc(self.N, C, h, w, self.dev)
        key = (C, h, w, slot)
        if key not in self._gz:
            self._gz[key] = ops.z_alloc(self.N, C, h, w, self.dev)
        return self._gz[key]

    def _build_plan(self):
        N, H, W = self.N, self.H, self.W
        self.fwd, self.recs = [], []
        self.raw_outputs = {}
        self.pack_fwd, self.pack_bwd = [], []
        self.pack2_fwd, self.pack2_bwd = [], []
        self._gz = {}                    # transient gradient-operand planes shared by all inception blocks, by size
        self.scratch = ops.bn_scratch(self.dev)
        self.conv_scratch = [ops.bn_scratch(self.dev) for _ in range(3)]   # one per concurrently running conv
        self.eval_affine = []            # (a, b, lo, hi, running_mean, running_var, gamma, beta) for eval mode
        self.img4 = self._zeros(N, H, W, 4)
        self.depth = self._zeros(N, H, W, 1)
        self.dld4 = self._zeros(N, H, W, 4)
        img = _T(self.img4, C=3)

        # --- conv1 7x7 (3 -> 128) + BN(affine) + ReLU (hourglass.py:164-166)
        r0 = self._zeros(N, H, W, 128)
        t0 = _T(r0, C=128, a=self._zeros(128), b=self._zeros(128), relu=True, dbuf=self._zeros(N, H, W, 128))
        t0.rstd, t0.mean, t0.bw = self._zeros(128), self._zeros(128), self._zeros(128, 4)
        fuse = self.fuse_bn
        self._conv(img, "seq.0.weight", "seq.0.bias", _T(r0), 3, 128, 7, N, H, W,
                   bn=(t0, self._rb("seq.1.running_mean", 128), self._rb("seq.1.running_var", 128),
                       self._p("seq.1.weight"), self._p("seq.1.bias"), 0) if fuse else None)
        self._stats(t0, 0, 128, N * H * W, "seq.1.running_mean", "seq.1.running_var", "seq.1.weight", "seq.1.bias",
                    fused=fuse)
        self.recs.append(("conv1", img, t0))

        z = self._chan(t0, mc_arch.structure(), "seq.3", H, W)

        # --- pred layer 3x3 (64 -> 1), exp -> depth (hourglass.py:173,178; mannequin_challenge_model.py:66)
        self._conv(z, "pred_layer.weight", "pred_layer.bias", _T(self.depth), 64, 1, 3, N, H, W, flags=ops.FLAG_EXP)
        self.recs.append(("pred", z))
        self._emit_backward()
        self.pack_fwd_tab = ops.make_pack_table(self.pack_fwd, self.dev)
        self.pack_bwd_tab = ops.make_pack_table(self.pack_bwd, self.dev)
        self.pack2_fwd_tab = ops.make_pack2_table(self.pack2_fwd, self.dev) if self.pack2_fwd else None
        self.pack2_bwd_tab = ops.make_pack2_table(self.pack2_bwd, self.dev) if self.pack2_bwd else None

    # forward helpers ------------------------------------------------------
    def _conv(self, x, wkey, bkey, dst, cin, cout, k, N, h, w, flags=0, wshape=None, bn=None):
        Wt = self._p(wkey, shape=wshape)
        bias = self._p(bkey, n=cout, shape=(cout,))
        pk = self._packed(cin, cout, k)
        prec = self.prec
        self.pack_fwd.append((Wt, pk, False))
        self.raw_outputs[wkey[:-7]] = (dst.buf, dst.off, cout)       # conv prefix -> where its raw output lives
        s, d = x.src(), ops.make_dst(dst.view())
        bns = None
        if bn is not None:     # (tensor record, running_mean, running_var, gamma, beta, scratch index): fused batch statistics
            t, rm, rv, gamma, beta, si = bn
            bns = ops.make_bn(self.conv_scratch[si], t.a, t.b, t.rstd, t.mean, gamma, beta, rm, rv)
        self.fwd.append(lambda: ops.conv(s, pk, bias, d, N, h, w, cin, cout, k, prec, flags,
                                         bn=bns if self.train_mode else None))
        return pk

    def _stats(self, t, lo, cnt, npix, rm_key, rv_key, g_key=None, b_key=None, fused=False):
        """Normalisation constants of channels [lo, lo+cnt) of t.  fused: the producing convs computed the
        train-mode batch statistics in their epilogues already (cvd_conv_fwd_bn); only eval mode needs work."""
        rm, rv = self._rb(rm_key, cnt), self._rb(rv_key, cnt)
        gamma = self._p(g_key) if g_key else None
        beta = self._p(b_key) if b_key else None
        buf, a, b, rstd, mean, scratch = t.buf, t.a, t.b, t.rstd, t.mean, self.scratch

        def run():
            if self.train_mode:
                if not fused:
                    ops.bn_stats(buf, lo, cnt, npix, scratch, a, b, rstd, mean, gamma, beta, rm, rv)
            else:                                   # eval(): running statistics (depth_fine_tuning.py:182)
                av = torch.rsqrt(rv + 1e-5)
                if gamma is not None:
                    av = av * gamma
                a[lo:lo + cnt] = av
                b[lo:lo + cnt] = (beta if beta is not None else 0) - rm * av
        self.fwd.append(run)

    def _inception(self, x, prefix, cfg, h, w):
        N = self.N
        o0 = cfg[0][0]
        ks, As, Bs = [c[0] for c in cfg[1:]], [c[1] for c in cfg[1:]], [c[2] for c in cfg[1:]]
        A, Bt = sum(As), sum(Bs)
        Ct = o0 + A + Bt
        buf, dbuf = self._zeros(N, h, w, Ct), self._zeros(N, h, w, Ct)
        arr = lambda: self._zeros(Ct)
        a, b, rstd, mean, bw = arr(), arr(), arr(), arr(), self._zeros(Ct, 4)

        def sub(off, C):
            t = _T(buf, off=off, C=C, a=a, b=b, relu=True, dbuf=dbuf)
            t.rstd, t.mean, t.bw = rstd, mean, bw
            return t

        one = sub(0, o0 + A)                                   # fused 1x1 output (o0 | a1 a2 a3)
        cin = x.C
        fuse = self.fuse_bn
        rm1 = self._rb(f"{prefix}.convs.0.1.running_mean", o0 + A)
        rv1 = self._rb(f"{prefix}.convs.0.1.running_var", o0 + A)
        v2 = self.v2 and fuse
        if v2:
            self._conv2(self._z_of(x, h, w), 0, f"{prefix}.convs.0.0.weight", f"{prefix}.convs.0.0.bias", _T(buf), cin, o0 + A, 1,
                        h, w, wshape=(o0 + A, cin, 1, 1), bn=(one, rm1, rv1, None, None, 0))
        else:
            self._conv(x, f"{prefix}.convs.0.0.weight", f"{prefix}.convs.0.0.bias", _T(buf), cin, o0 + A, 1, N, h, w,
                       wshape=(o0 + A, cin, 1, 1), bn=(one, rm1, rv1, None, None, 0) if fuse else None)
        self._stats(one, 0, o0 + A, N * h * w, f"{prefix}.convs.0.1.running_mean", f"{prefix}.convs.0.1.running_var",
                    fused=fuse)
        rmk = self._rb(f"{prefix}.convs.1.4.running_mean", Bt)
        rvk = self._rb(f"{prefix}.convs.1.4.running_var", Bt)
        mids, outs = [], []
        aoff, boff = o0, o0 + A
        zmid = None
        if v2:                                                 # the three k x k convs read ONE prepared tensor (a1 | a2 | a3)
            zmid = self._z_of(sub(o0, A), h, w)
        main_list, branches = self.fwd, []
        for i in range(3):                                     # the three k x k convs are independent: parallel branches
            mid = sub(aoff, As[i])
            mid.zsrc = (zmid, (aoff - o0) // 8)                # where this conv's input lives in the prepared planes
            self.fwd = []
            ko = boff - (o0 + A)
            if v2 and self._use_conv2(As[i], Bs[i], ks[i]):
                self._conv2(zmid, (aoff - o0) // 8, f"{prefix}.convs.{i + 1}.3.weight", f"{prefix}.convs.{i + 1}.3.bias",
                            _T(buf, off=boff), As[i], Bs[i], ks[i], h, w, bn=(one, rmk[ko:ko + Bs[i]], rvk[ko:ko + Bs[i]], None, None, i))
            else:
                self._conv(mid, f"{prefix}.convs.{i + 1}.3.weight", f"{prefix}.convs.{i + 1}.3.bias", _T(buf, off=boff),
                           As[i], Bs[i], ks[i], N, h, w,
                           bn=(one, rmk[ko:ko + Bs[i]], rvk[ko:ko + Bs[i]], None, None, i) if fuse else None)
            branches.append(self.fwd)
            mids.append(mid)
            outs.append(sub(boff, Bs[i]))
            aoff += As[i]
            boff += Bs[i]
        self.fwd = main_list
        self.fwd.append(("par", branches))
        kout = sub(o0 + A, Bt)
        self._stats(kout, o0 + A, Bt, N * h * w, f"{prefix}.convs.1.4.running_mean", f"{prefix}.convs.1.4.running_var",
                    fused=fuse)
        out = _T(buf, off=0, n0=o0, gap=A, C=o0 + Bt, a=a, b=b, relu=True, dbuf=dbuf)
        out.rstd, out.mean, out.bw = rstd, mean, bw
        self.recs.append(("inc", x, prefix, cfg, h, w, one, mids, outs, kout))
        return out

    def _pool(self, x, h, w):
        N = self.N
        p = _T(self._zeros(N, h // 2, w // 2, x.C), C=x.C, dbuf=self._zeros(N, h // 2, w // 2, x.C))
        xv, a, b, relu, pb, C = x.view(), x.a, x.b, x.relu, p.buf, x.C
        self.fwd.append(lambda: ops.pool_fwd(xv, a, b, relu, pb, N, h, w, C))
        self.recs.append(("pool", x, p, h, w))
        return p

    def _merge(self, skip, inner, h, w):
        N = self.N
        assert skip.C == inner.C
        z = _T(self._zeros(N, h, w, skip.C), C=skip.C, dbuf=self._zeros(N, h, w, skip.C))
        v1, a1, b1, v2, a2, b2, zb, C = skip.view(), skip.a, skip.b, inner.view(), inner.a, inner.b, z.buf, skip.C
        self.fwd.append(lambda: ops.merge_up_fwd(v1, a1, b1, v2, a2, b2, zb, N, h, w, C))
        self.recs.append(("merge", skip, inner, z, h, w))
        return z

    def _chan(self, x, node, prefix, h, w):
        res = []
        for bi, branch in enumerate(node[1:]):
            t, hh, ww, up = x, h, w, False
            for oi, op in enumerate(branch):
                name = f"{prefix}.list.{bi}.{oi}"
                if op[0] == "pool":
                    t = self._pool(t, hh, ww)
                    hh, ww = hh // 2, ww // 2
                elif op[0] == "inc":
                    t = self._inception(t, name, op[2], hh, ww)
                elif op[0] == "chan":
                    t = self._chan(t, op, name, hh, ww)
                elif op[0] == "up":
                    up = True
            res.append((t, up))
        skip = [t for t, up in res if not up][0]
        inner = [t for t, up in res if up][0]
        return self._merge(skip, inner, h, w)

    # backward plan ----------------------------------------------------------
    def _emit_backward(self):
        self.bwd = []
        N, prec, scratch = self.N, self.prec, self.scratch
        for rec in reversed(self.recs):
            kind = rec[0]
            if kind == "pred":
                z = rec[1]
                H, W = self.H, self.W
                Wt = self._p("pred_layer.weight")
                pkt = self._packed(1, 64, 3)
                self.pack_bwd.append((Wt, pkt, True))
                dld = _T(self.dld4, C=1)
                gs, xs = dld.src(), z.src()
                dW, db = self._g("pred_layer.weight"), self._g("pred_layer.bias")
                dld4, depth = self.dld4, self.depth
                self.bwd.append(lambda: ops.dlogdepth(self.grad_depth, depth, dld4, db))
                self.bwd.append(("wg", [lambda gs=gs, xs=xs, dW=dW: ops.conv_wgrad(gs, xs, dW, N, H, W, 64, 1, 3, prec)]))
                d = ops.make_dst(z.dview())
                fl = ops.FLAG_ACCUM if z.grad_written else 0
                self.bwd.append(lambda gs=gs, pkt=pkt, d=d, fl=fl: ops.conv(gs, pkt, None, d, N, H, W, 1, 64, 3, prec, fl))
                z.grad_written = True
            elif kind == "merge":
                _, skip, inner, z, h, w = rec
                assert z.grad_written
                dz, v2, v1, acc1, C = z.dbuf, inner.dview(), skip.dview(), skip.grad_written, z.C
                self.bwd.append(lambda dz=dz, v2=v2, v1=v1, acc1=acc1, h=h, w=w, C=C:
                                ops.merge_up_bwd(dz, v2, v1, acc1, N, h, w, C))
                skip.grad_written = inner.grad_written = True
            elif kind == "pool":
                _, x, p, h, w = rec
                assert p.grad_written
                dp, xv, acc, C = p.dbuf, x.dview(), x.grad_written, x.C
                self.bwd.append(lambda dp=dp, xv=xv, acc=acc, h=h, w=w, C=C: ops.pool_bwd(dp, xv, acc, N, h, w, C))
                x.grad_written = True
            elif kind == "inc":
                _, x, prefix, cfg, h, w, one, mids, outs, kout = rec
                o0 = cfg[0][0]
                ks, As, Bs = [c[0] for c in cfg[1:]], [c[1] for c in cfg[1:]], [c[2] for c in cfg[1:]]
                A, Bt = sum(As), sum(Bs)
                npix = N * h * w
                buf, dbuf = one.buf, one.dbuf
                a, b, rstd, mean, bw = one.a, one.b, one.rstd, one.mean, one.bw
                dbk = self.grad_flat[self.pmap[f"{prefix}.convs.1.3.bias"][0]:][:Bt]
                self.bwd.append(lambda buf=buf, dbuf=dbuf, a=a, b=b, rstd=rstd, mean=mean, bw=bw, dbk=dbk, lo=o0 + A, cnt=Bt, npix=npix:
                                ops.bn_bwd_reduce(buf, lo, cnt, dbuf, npix, scratch, a, b, rstd, mean, bw, True, dbias=dbk))
                kbranches = []
                v2 = self.v2 and self.fuse_bn
                if v2:      # gradient wrt the k x k convs' raw outputs (BN+ReLU backward), prepared once for the three dgrads
                    gzk = self._gz_planes(Bt, h, w, 0)
                    gsrc_all = kout.bnbwd_src()
                    self.bwd.append(lambda gsrc_all=gsrc_all, gzk=gzk, Bt=Bt: ops.prep_operand(gsrc_all, Bt, gzk))
                boffs = [sum(Bs[:i]) for i in range(3)]
                for i in range(3):
                    main_bwd, self.bwd = self.bwd, []
                    Wt = self._p(f"{prefix}.convs.{i + 1}.3.weight")
                    gs, xs = outs[i].bnbwd_src(), mids[i].src()
                    dW = self._g(f"{prefix}.convs.{i + 1}.3.weight")
                    if v2 and self.v2_wgrad and self._use_wgrad2(As[i], Bs[i], ks[i]):
                        zmid, xo = mids[i].zsrc
                        self.bwd.append(lambda zmid=zmid, xo=xo, gzk=gzk, go=boffs[i] // 8, gs=gs, xs=xs, dW=dW, ci=As[i], co=Bs[i], k=ks[i], h=h, w=w:
                                        ops.conv2_wgrad(zmid, xo, gzk, go, dW, N, h, w, ci, co, k) or
                                        ops.conv_wgrad(gs, xs, dW, N, h, w, ci, co, k, prec))
                    else:
                        self.bwd.append(lambda gs=gs, xs=xs, dW=dW, ci=As[i], co=Bs[i], k=ks[i], h=h, w=w:
                                        ops.conv_wgrad(gs, xs, dW, N, h, w, ci, co, k, prec))
                    d = ops.make_dst(mids[i].dview())
                    if v2 and self._use_conv2(Bs[i], As[i], ks[i]):
                        pkt = torch.empty(ops.conv2_packed_bytes(Bs[i], As[i], ks[i]), dtype=torch.uint8, device=self.dev)
                        self.pack2_bwd.append((Wt, pkt, True))
                        self.bwd.append(lambda gzk=gzk, zo=boffs[i] // 8, pkt=pkt, d=d, ci=Bs[i], co=As[i], k=ks[i], h=h, w=w:
                                        ops.conv2(gzk, zo, pkt, None, d, N, h, w, ci, co, k, 0, None))
                    else:
                        pkt = self._packed(Bs[i], As[i], ks[i])
                        self.pack_bwd.append((Wt, pkt, True))
                        self.bwd.append(lambda gs=gs, pkt=pkt, d=d, ci=Bs[i], co=As[i], k=ks[i], h=h, w=w:
                                        ops.conv(gs, pkt, None, d, N, h, w, ci, co, k, prec, 0))
                    kbranches.append([[self.bwd[0]], [self.bwd[1]]])
                    self.bwd = main_bwd
                if self.wg_async:
                    # the three weight gradients leave the critical path; the three dgrads stay parallel branches
                    self.bwd.append(("wg", [pair[0][0] for pair in kbranches]))
                    self.bwd.append(("par", [pair[1] for pair in kbranches]))
                else:
                    # wgrad and dgrad of each of the three convs: six independent kernels
                    self.bwd.append(("par", [b for pair in kbranches for b in pair]))
                db1 = self.grad_flat[self.pmap[f"{prefix}.convs.0.0.bias"][0]:][:o0 + A]
                self.bwd.append(lambda buf=buf, dbuf=dbuf, a=a, b=b, rstd=rstd, mean=mean, bw=bw, db1=db1, cnt=o0 + A, npix=npix:
                                ops.bn_bwd_reduce(buf, 0, cnt, dbuf, npix, scratch, a, b, rstd, mean, bw, True, dbias=db1))
                cin = x.C
                W1 = self._p(f"{prefix}.convs.0.0.weight", shape=(o0 + A, cin, 1, 1))
                dW1 = self._g(f"{prefix}.convs.0.0.weight", shape=(o0 + A, cin, 1, 1))
                gs, xs = one.bnbwd_src(), x.src()
                wg = (lambda gs=gs, xs=xs, dW1=dW1, cin=cin, co=o0 + A, h=h, w=w:
                      ops.conv_wgrad(gs, xs, dW1, N, h, w, cin, co, 1, prec))
                if v2:                                             # gradient wrt the fused 1x1 outputs, prepared once
                    gz1 = self._gz_planes(o0 + A, h, w, 1)
                    self.bwd.append(lambda gs=gs, gz1=gz1, c1=o0 + A: ops.prep_operand(gs, c1, gz1))
                    if self.v2_wgrad:
                        wg = (lambda xz=x.z, gz1=gz1, gs=gs, xs=xs, dW1=dW1, cin=cin, co=o0 + A, h=h, w=w:
                              ops.conv2_wgrad(xz, 0, gz1, 0, dW1, N, h, w, cin, co, 1) or
                              ops.conv_wgrad(gs, xs, dW1, N, h, w, cin, co, 1, prec))
                if x.dbuf is not None:
                    d = ops.make_dst(x.dview())
                    fl = ops.FLAG_ACCUM if x.grad_written else 0
                    if v2:
                        pkt = torch.empty(ops.conv2_packed_bytes(o0 + A, cin, 1), dtype=torch.uint8, device=self.dev)
                        self.pack2_bwd.append((W1, pkt, True))
                        dg = (lambda gz1=gz1, pkt=pkt, d=d, fl=fl, ci=o0 + A, co=cin, h=h, w=w:
                              ops.conv2(gz1, 0, pkt, None, d, N, h, w, ci, co, 1, fl, None))
                    else:
                        pkt = self._packed(o0 + A, cin, 1)
                        self.pack_bwd.append((W1, pkt, True))
                        dg = (lambda gs=gs, pkt=pkt, d=d, fl=fl, ci=o0 + A, co=cin, h=h, w=w:
                              ops.conv(gs, pkt, None, d, N, h, w, ci, co, 1, prec, fl))
                    if self.wg_async:
                        self.bwd.append(("wg", [wg]))
                        self.bwd.append(dg)
                    else:
                        self.bwd.append(("par", [[wg], [dg]]))
                    x.grad_written = True
                else:
                    self.bwd.append(("wg", [wg]))
            elif kind == "conv1":
                _, img, t0 = rec
                H, W = self.H, self.W
                assert t0.grad_written
                dg, dbt, dbias = self._g("seq.1.weight"), self._g("seq.1.bias"), self._g("seq.0.bias")
                gamma, beta = self._p("seq.1.weight"), self._p("seq.1.bias")
                self.bwd.append(lambda t0=t0: ops.bn_bwd_reduce(t0.buf, 0, 128, t0.dbuf, N * H * W, scratch, t0.a, t0.b, t0.rstd,
                                                                t0.mean, t0.bw, True, gamma, beta, dg, dbt, dbias))
                gs, xs, dW = t0.bnbwd_src(), img.src(), self._g("seq.0.weight")
                self.bwd.append(("wg", [lambda gs=gs, xs=xs, dW=dW: ops.conv_wgrad(gs, xs, dW, N, H, W, 3, 128, 7, prec)]))

    # ------------------------------------------------------------------ execution
    def _run(self, plan):
        """Plan entries are callables (serial, current stream) or ("par", [branch, ...]): independent branches
        forked onto side streams and joined back — inside the CUDA graph they become parallel branches, which
        keeps the SMs busy on the low-resolution hourglass levels whose kernels have fewer CTAs than the GPU has SMs."""
        main = None
        wg_used = 0
        for op in plan:
            if not isinstance(op, tuple):
                op()
                continue
            if op[0] == "wg":            # deferred weight gradients (see __init__): side stream, joined after the plan
                if not (self.multi_stream and self.wg_async):
                    for f in op[1]:
                        f()
                    continue
                if main is None:
                    main = torch.cuda.current_stream()
                while len(self.wg_streams) < 2:
                    self.wg_streams.append(torch.cuda.Stream(device=self.dev))
                for f in op[1]:
                    sw = self.wg_streams[wg_used % 2]
                    wg_used += 1
                    sw.wait_stream(main)
                    with torch.cuda.stream(sw):
                        f()
                continue
            branches = op[1]
            if not self.multi_stream or len(branches) == 1:
                for br in branches:
                    for f in br:
                        f()
                continue
            if main is None:
                main = torch.cuda.current_stream()
            while len(self.side_streams) < len(branches) - 1:
                self.side_streams.append(torch.cuda.Stream(device=self.dev))
            if not self.fork_first:
                for f in branches[0]:
                    f()
            for i, br in enumerate(branches[1:]):
                s = self.side_streams[i]
                s.wait_stream(main)
                with torch.cuda.stream(s):
                    for f in br:
                        f()
            if self.fork_first:
                for f in branches[0]:
                    f()
            for i in range(len(branches) - 1):
                main.wait_stream(self.side_streams[i])
        if wg_used:
            for sw in self.wg_streams:
                main.wait_stream(sw)

    def forward(self, images):
        """images (N,3,H,W) BGR in [0,1] (CUDA, contiguous) -> depth (N,H,W) (engine-owned buffer)."""
        assert images.shape == (self.N, 3, self.H, self.W), images.shape
        ops.image_to_nhwc4(images.contiguous(), self.img4, self.N, self.H, self.W)
        ops.pack_batch(self.pack_fwd_tab[0], self.pack_fwd_tab[1], self.prec)
        if self.pack2_fwd_tab is not None:
            ops.conv2_pack_batch(*self.pack2_fwd_tab)
        self._run(self.fwd)
        if self.train_mode:
            self.P.num_batches_tracked += 1
        return self.depth.view(self.N, self.H, self.W)

    def backward(self, grad_depth):
        """grad_depth (N,H,W): d loss / d depth.  Accumulates into grad_flat (zero it first, as opt.zero_grad does)."""
        self.grad_depth = grad_depth.contiguous()
        ops.pack_batch(self.pack_bwd_tab[0], self.pack_bwd_tab[1], self.prec)
        if self.pack2_bwd_tab is not None:
            ops.conv2_pack_batch(*self.pack2_bwd_tab)
        self._run(self.bwd)
