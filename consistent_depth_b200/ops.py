"""Thin tensor-level wrappers over the conv-engine C-ABI (include/cvd.h).

Every function takes CUDA fp32 tensors in the engine's NHWC layout plus channel
*views*; nothing here has a CPU implementation.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import cvd_dst_t, cvd_src_t

XF_AFFINE, XF_BNBWD = 0, 1
FLAG_ACCUM, FLAG_EXP = 1, 2


class View:
    """Channel view of an NHWC buffer: logical channel c -> physical off + c + (c >= n0 ? gap : 0)."""
    __slots__ = ("t", "off", "n0", "gap")

    def __init__(self, t, off=0, n0=0, gap=0):
        self.t, self.off, self.n0, self.gap = t, off, n0, gap

    @property
    def c_total(self):
        return self.t.shape[-1]


def make_src(x, a=None, b=None, relu=False, dy=None, bw=None):
    s = cvd_src_t()
    s.x = x.t.data_ptr()
    s.c_total, s.c_off, s.n0, s.gap = x.c_total, x.off, x.n0, x.gap
    s.a = a.data_ptr() if a is not None else None
    s.b = b.data_ptr() if b is not None else None
    s.relu = 1 if relu else 0
    if dy is not None:
        s.mode = XF_BNBWD
        s.dy = dy.t.data_ptr()
        s.dy_ctotal, s.dy_coff, s.dy_n0, s.dy_gap = dy.c_total, dy.off, dy.n0, dy.gap
        s.bw = bw.data_ptr()
    else:
        s.mode = XF_AFFINE
    s._keep = (x.t, a, b, dy.t if dy is not None else None, bw)   # keep the tensors alive with the struct
    return s


def make_dst(y):
    d = cvd_dst_t()
    d.y = y.t.data_ptr()
    d.c_total, d.c_off, d.n0, d.gap = y.c_total, y.off, y.n0, y.gap
    d._keep = (y.t,)
    return d


def packed_bytes(cin, cout, k, precision):
    return int(_lib.lib().cvd_conv_packed_bytes(cin, cout, k, precision))


def pack_weights(w_oihw, transpose_flip=False, precision=3, out=None):
    """fp32 OIHW -> streamed bf16 hi/lo core-matrix blobs (uint8 tensor)."""
    cout, cin, k, _ = w_oihw.shape
    n = packed_bytes(cin, cout, k, precision)
    if out is None:
        out = torch.empty(n, dtype=torch.uint8, device=w_oihw.device)
    _lib.check(_lib.lib().cvd_conv_pack_weights(_lib.ptr(w_oihw), cin, cout, k, 1 if transpose_flip else 0,
                                                precision, _lib.ptr(out), _lib.stream()), "cvd_conv_pack_weights")
    return out


def make_bn(scratch, a, b, rstd, mean, gamma=None, beta=None, running_mean=None, running_var=None, eps=1e-5,
            momentum=0.1):
    """cvd_bn_t for conv(..., bn=): a/b/rstd/mean are the destination buffer's per-physical-channel arrays,
    gamma/beta/running_* this conv's own [cout] arrays."""
    ts = (scratch, gamma, beta, running_mean, running_var, a, b, rstd, mean)
    s = _lib.cvd_bn_t(*[_lib.ptr(t).value if t is not None else None for t in ts], eps, momentum)
    s._keep = ts
    return s


def conv(src, packed, bias, dst, N, H, W, cin, cout, k, precision=3, flags=0, bn=None):
    """src: cvd_src_t, dst: cvd_dst_t (from make_src/make_dst); cin/cout in GEMM terms.
    bn (make_bn): also produce the train-mode BatchNorm statistics of the output in the conv epilogue."""
    if bn is not None:
        _lib.check(_lib.lib().cvd_conv_fwd_bn(C.byref(src), _lib.ptr(packed), _lib.ptr(bias), C.byref(dst),
                                              N, H, W, cin, cout, k, precision, flags, C.byref(bn), _lib.stream()),
                   "cvd_conv_fwd_bn")
        return
    _lib.check(_lib.lib().cvd_conv_fwd(C.byref(src), _lib.ptr(packed), _lib.ptr(bias), C.byref(dst),
                                       N, H, W, cin, cout, k, precision, flags, _lib.stream()), "cvd_conv_fwd")


def conv_chunks(src, packed, bias, dst, N, H, W, cin, cout, k, nchunks, src_shift, dst_shift, packed_stride,
                precision=3, flags=0, bn=None):
    """nchunks same-shape convs in one launch: chunk j = views shifted by j*src_shift / j*dst_shift channels, weights
    at packed + j*packed_stride bytes (bn: scratch from bn_scratch(device, 256 * nchunks))."""
    _lib.check(_lib.lib().cvd_conv_fwd_chunks(C.byref(src), _lib.ptr(packed), _lib.ptr(bias), C.byref(dst),
                                              N, H, W, cin, cout, k, precision, flags, C.byref(bn) if bn is not None else None,
                                              nchunks, src_shift, dst_shift, C.c_longlong(packed_stride), _lib.stream()),
               "cvd_conv_fwd_chunks")


def conv_wgrad(gsrc, xsrc, dw, N, H, W, cin, cout, k, precision=3):
    """dw (fp32 OIHW, pre-zeroed or accumulating) += G (x) X."""
    _lib.check(_lib.lib().cvd_conv_wgrad(C.byref(gsrc), C.byref(xsrc), _lib.ptr(dw), N, H, W, cin, cout, k,
                                         precision, _lib.stream()), "cvd_conv_wgrad")


def bn_scratch(device, C=256):
    n = int(_lib.lib().cvd_bn_scratch_bytes(C))
    return torch.zeros((n + 7) // 8, dtype=torch.float64, device=device)


def bn_stats(x, c_off, Cn, npix, scratch, a, b, rstd, mean, gamma=None, beta=None, running_mean=None,
             running_var=None, eps=1e-5, momentum=0.1):
    _lib.check(_lib.lib().cvd_bn_stats(_lib.ptr(x), x.shape[-1], c_off, Cn, C.c_longlong(npix), _lib.ptr(scratch),
                                       _lib.ptr(gamma), _lib.ptr(beta), C.c_float(eps), C.c_float(momentum),
                                       _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(a), _lib.ptr(b),
                                       _lib.ptr(rstd), _lib.ptr(mean), _lib.stream()), "cvd_bn_stats")


def bn_bwd_reduce(x, c_off, Cn, dy, npix, scratch, a, b, rstd, mean, bw, relu=True, gamma=None, beta=None,
                  dgamma=None, dbeta=None, dbias=None, dy_view=None):
    """dy defaults to the same physical layout as x (dy_view = (c_total, c_off, n0, gap, lc0) overrides)."""
    if dy_view is None:
        dy_view = (dy.shape[-1], 0, 0, 0, c_off)
    _lib.check(_lib.lib().cvd_bn_bwd_reduce(_lib.ptr(x), x.shape[-1], c_off, _lib.ptr(dy), dy_view[0], dy_view[1],
                                            dy_view[2], dy_view[3], dy_view[4], _lib.ptr(a), _lib.ptr(b),
                                            _lib.ptr(rstd), _lib.ptr(mean), _lib.ptr(gamma), _lib.ptr(beta),
                                            1 if relu else 0, C.c_longlong(npix), Cn, _lib.ptr(scratch), _lib.ptr(bw),
                                            _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dbias), _lib.stream()),
               "cvd_bn_bwd_reduce")


def pool_fwd(xv, a, b, relu, p, N, H, W, Cn):
    _lib.check(_lib.lib().cvd_pool_fwd(_lib.ptr(xv.t), xv.c_total, xv.off, xv.n0, xv.gap, _lib.ptr(a), _lib.ptr(b),
                                       1 if relu else 0, _lib.ptr(p), N, H, W, Cn, _lib.stream()), "cvd_pool_fwd")


def pool_bwd(dp, dxv, accumulate, N, H, W, Cn):
    """dxv: View of the full-resolution gradient buffer."""
    _lib.check(_lib.lib().cvd_pool_bwd(_lib.ptr(dp), _lib.ptr(dxv.t), dxv.c_total, dxv.off, dxv.n0, dxv.gap,
                                       1 if accumulate else 0, N, H, W, Cn, _lib.stream()), "cvd_pool_bwd")


def merge_up_fwd(x1v, a1, b1, x2v, a2, b2, z, N, H, W, Cn):
    _lib.check(_lib.lib().cvd_merge_up_fwd(_lib.ptr(x1v.t), x1v.c_total, x1v.off, x1v.n0, x1v.gap, _lib.ptr(a1), _lib.ptr(b1),
                                           _lib.ptr(x2v.t), x2v.c_total, x2v.off, x2v.n0, x2v.gap, _lib.ptr(a2), _lib.ptr(b2),
                                           _lib.ptr(z), N, H, W, Cn, _lib.stream()), "cvd_merge_up_fwd")


def merge_up_bwd(dz, dy2v, dy1v, accumulate1, N, H, W, Cn):
    """dz plain (N,H,W,Cn); dy2v: View at half resolution; dy1v: View at full resolution or None."""
    L = _lib.lib()
    if dy1v is None:
        rc = L.cvd_merge_up_bwd(_lib.ptr(dz), _lib.ptr(dy2v.t), dy2v.c_total, dy2v.off, dy2v.n0, dy2v.gap,
                                None, 0, 0, 0, 0, 0, N, H, W, Cn, _lib.stream())
    else:
        rc = L.cvd_merge_up_bwd(_lib.ptr(dz), _lib.ptr(dy2v.t), dy2v.c_total, dy2v.off, dy2v.n0, dy2v.gap,
                                _lib.ptr(dy1v.t), dy1v.c_total, dy1v.off, dy1v.n0, dy1v.gap, 1 if accumulate1 else 0,
                                N, H, W, Cn, _lib.stream())
    _lib.check(rc, "cvd_merge_up_bwd")


def image_to_nhwc4(img, out, N, H, W):
    _lib.check(_lib.lib().cvd_image_to_nhwc4(_lib.ptr(img), _lib.ptr(out), N, H, W, _lib.stream()), "cvd_image_to_nhwc4")


def dlogdepth(grad_depth, depth, out4, dbias=None):
    _lib.check(_lib.lib().cvd_dlogdepth(_lib.ptr(grad_depth), _lib.ptr(depth), _lib.ptr(out4),
                                        C.c_longlong(depth.numel()), _lib.ptr(dbias), _lib.stream()), "cvd_dlogdepth")


def make_pack_table(entries, device):
    """entries: [(w_oihw tensor, packed uint8 tensor, transpose_flip)] -> device descriptor table for pack_batch."""
    import numpy as np
    dt = np.dtype([("w", "<u8"), ("out", "<u8"), ("cin", "<i4"), ("cout", "<i4"), ("k", "<i4"), ("flip", "<i4")])
    arr = np.zeros(len(entries), dtype=dt)
    for i, e in enumerate(entries):
        w, out, flip = e[0], e[1], e[2]
        if len(e) == 3:
            arr[i] = (w.data_ptr(), out.data_ptr(), w.shape[1], w.shape[0], w.shape[2], 1 if flip else 0)
        else:       # (w rows of a grouped weight, packed, flip, chunk width, group size): block-diagonal chunk
            arr[i] = (w.data_ptr(), out.data_ptr(), e[3], e[3], w.shape[2], (1 if flip else 0) | (e[4] << 8))
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    t._keep = [e[0] for e in entries] + [e[1] for e in entries]
    return t, len(entries)


def pack_batch(table, n, precision):
    _lib.check(_lib.lib().cvd_conv_pack_batch(_lib.ptr(table), n, precision, _lib.stream()), "cvd_conv_pack_batch")


# ---------------------------------------------------------------- monodepth2 passes (include/cvd.h, second half)
GATHER_IDENTITY, GATHER_ELU, GATHER_AFFINE_RELU = 0, 1, 2


def bicubic_image(img, out4, mean=0.45, std=0.225):
    """img (N,3,H,W) -> out4 (N,oh,ow,4) = (bicubic(img) - mean) / std."""
    N, _, H, W = img.shape
    _lib.check(_lib.lib().cvd_bicubic_image_fwd(_lib.ptr(img), N, H, W, _lib.ptr(out4), out4.shape[1], out4.shape[2],
                                                C.c_float(mean), C.c_float(1.0 / std), _lib.stream()), "cvd_bicubic_image_fwd")


def disp_to_depth(disp, depth):
    N, fh, fw = disp.shape
    _lib.check(_lib.lib().cvd_disp_to_depth_fwd(_lib.ptr(disp), N, fh, fw, _lib.ptr(depth), depth.shape[-2], depth.shape[-1],
                                                _lib.stream()), "cvd_disp_to_depth_fwd")


def disp_to_depth_bwd(ddepth, depth, ddisp):
    N, fh, fw = ddisp.shape
    _lib.check(_lib.lib().cvd_disp_to_depth_bwd(_lib.ptr(ddepth), _lib.ptr(depth), N, fh, fw, depth.shape[-2], depth.shape[-1],
                                                _lib.ptr(ddisp), _lib.stream()), "cvd_disp_to_depth_bwd")


def sigmoid_fwd(raw_padded, disp):
    N, fh, fw = disp.shape
    _lib.check(_lib.lib().cvd_sigmoid_fwd(_lib.ptr(raw_padded), raw_padded.shape[-1], N, fh, fw, _lib.ptr(disp), _lib.stream()),
               "cvd_sigmoid_fwd")


def sigmoid_bwd(ddisp, disp, draw_padded):
    N, fh, fw = disp.shape
    _lib.check(_lib.lib().cvd_sigmoid_bwd(_lib.ptr(ddisp), _lib.ptr(disp), N, fh, fw, _lib.ptr(draw_padded),
                                          draw_padded.shape[-1], _lib.stream()), "cvd_sigmoid_bwd")


def subsample2(src, dst):
    N, H, W, Cn = src.shape
    _lib.check(_lib.lib().cvd_subsample2(_lib.ptr(src), N, H, W, Cn, _lib.ptr(dst), _lib.stream()), "cvd_subsample2")


def stuff2(src, dst, accumulate):
    N, H, W, Cn = dst.shape
    _lib.check(_lib.lib().cvd_stuff2(_lib.ptr(src), N, H, W, Cn, _lib.ptr(dst), 1 if accumulate else 0, _lib.stream()),
               "cvd_stuff2")


def bnbwd_stuff(x, dy, a, b, bw, relu, dst, stride):
    N, h, w, Cn = x.shape
    _lib.check(_lib.lib().cvd_bnbwd_stuff(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(a), _lib.ptr(b), _lib.ptr(bw), 1 if relu else 0,
                                          N, h, w, Cn, _lib.ptr(dst), dst.shape[1], dst.shape[2], stride, _lib.stream()),
               "cvd_bnbwd_stuff")


def maxpool_fwd(x, a, b, relu, out, argmax):
    N, H, W, Cn = x.shape
    _lib.check(_lib.lib().cvd_maxpool3s2_fwd(_lib.ptr(x), _lib.ptr(a), _lib.ptr(b), 1 if relu else 0, N, H, W, Cn,
                                             _lib.ptr(out), _lib.ptr(argmax), _lib.stream()), "cvd_maxpool3s2_fwd")


def maxpool_bwd(dout, argmax, dx, accumulate):
    N, H, W, Cn = dx.shape
    _lib.check(_lib.lib().cvd_maxpool3s2_bwd(_lib.ptr(dout), _lib.ptr(argmax), N, H, W, Cn, _lib.ptr(dx),
                                             1 if accumulate else 0, _lib.stream()), "cvd_maxpool3s2_bwd")


def bn_add_relu(y, a, b, res, ra, rb, out):
    Cn = y.shape[-1]
    _lib.check(_lib.lib().cvd_bn_add_relu(_lib.ptr(y), _lib.ptr(a), _lib.ptr(b), _lib.ptr(res), _lib.ptr(ra), _lib.ptr(rb),
                                          C.c_longlong(y.numel() // Cn), Cn, _lib.ptr(out), _lib.stream()), "cvd_bn_add_relu")


def relu_bwd_add(dout, out, dres, accumulate):
    _lib.check(_lib.lib().cvd_relu_bwd_add(_lib.ptr(dout), _lib.ptr(out), _lib.ptr(dres), 1 if accumulate else 0,
                                           C.c_longlong(dout.numel()), _lib.stream()), "cvd_relu_bwd_add")


def gather_pad_fwd(src, s_coff, s_pad, a, b, dst, d_coff, Cn, upsample, mode):
    """src (N, hs+2*s_pad, ws+2*s_pad, Cs) -> dst (N, (hs<<up)+2, (ws<<up)+2, Cd) channels [d_coff, d_coff+Cn)."""
    N = src.shape[0]
    hs, ws = src.shape[1] - 2 * s_pad, src.shape[2] - 2 * s_pad
    assert dst.shape[1] == (hs << upsample) + 2 and dst.shape[2] == (ws << upsample) + 2, (src.shape, dst.shape)
    _lib.check(_lib.lib().cvd_gather_pad_fwd(_lib.ptr(src), src.shape[-1], s_coff, s_pad, _lib.ptr(a), _lib.ptr(b),
                                             _lib.ptr(dst), dst.shape[-1], d_coff, N, hs, ws, Cn, upsample, mode,
                                             _lib.stream()), "cvd_gather_pad_fwd")


def gather_pad_bwd(dpad, p_coff, src, s_coff, s_pad, dsrc, ds_coff, ds_pad, Cn, upsample, mode, accumulate):
    N = dsrc.shape[0]
    hs, ws = dsrc.shape[1] - 2 * ds_pad, dsrc.shape[2] - 2 * ds_pad
    assert dpad.shape[1] == (hs << upsample) + 2 and dpad.shape[2] == (ws << upsample) + 2, (dpad.shape, dsrc.shape)
    _lib.check(_lib.lib().cvd_gather_pad_bwd(_lib.ptr(dpad), dpad.shape[-1], p_coff, _lib.ptr(src),
                                             src.shape[-1] if src is not None else 0, s_coff, s_pad, _lib.ptr(dsrc),
                                             dsrc.shape[-1], ds_coff, ds_pad, N, hs, ws, Cn, upsample, mode,
                                             1 if accumulate else 0, _lib.stream()), "cvd_gather_pad_bwd")


def channel_sum(x, c_off, Cn, out):
    _lib.check(_lib.lib().cvd_channel_sum(_lib.ptr(x), x.shape[-1], c_off, Cn, C.c_longlong(x.numel() // x.shape[-1]),
                                          _lib.ptr(out), _lib.stream()), "cvd_channel_sum")


# ---------------------------------------------------------------- MiDaS-v2 passes
def conv_wgrad_grouped(gsrc, xsrc, dw_rows, N, H, W, c, group_size, k, precision=3):
    _lib.check(_lib.lib().cvd_conv_wgrad_grouped(C.byref(gsrc), C.byref(xsrc), _lib.ptr(dw_rows), N, H, W, c, group_size, k,
                                                 precision, _lib.stream()), "cvd_conv_wgrad_grouped")


def conv_wgrad_grouped_chunks(gsrc, xsrc, dw, N, H, W, c, nchunks, group_size, k, precision=3):
    """All chunks of a grouped conv's weight gradient in one launch; gsrc / xsrc: views of chunk 0, dw: whole gradient."""
    _lib.check(_lib.lib().cvd_conv_wgrad_grouped_chunks(C.byref(gsrc), C.byref(xsrc), _lib.ptr(dw), N, H, W, c, nchunks,
                                                        group_size, k, precision, _lib.stream()), "cvd_conv_wgrad_grouped_chunks")


def pack_weights_grouped(w_rows, c, group_size, transpose_flip=False, precision=3):
    """Rows [chunk, chunk+c) of a (Cout, group_size, k, k) grouped weight -> packed dense c x c block-diagonal chunk."""
    k = w_rows.shape[2]
    out = torch.empty(packed_bytes(c, c, k, precision), dtype=torch.uint8, device=w_rows.device)
    _lib.check(_lib.lib().cvd_conv_pack_weights(_lib.ptr(w_rows), c, c, k, (1 if transpose_flip else 0) | (group_size << 8),
                                                precision, _lib.ptr(out), _lib.stream()), "cvd_conv_pack_weights")
    return out


def image_normalize(img, out4, mean, std):
    N, _, H, W = img.shape
    m, s = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    _lib.check(_lib.lib().cvd_image_normalize_nhwc4(_lib.ptr(img), N, H, W, m, s, _lib.ptr(out4), _lib.stream()),
               "cvd_image_normalize_nhwc4")


def relu_add(x, other, out):
    _lib.check(_lib.lib().cvd_relu_add(_lib.ptr(x), _lib.ptr(other), _lib.ptr(out), C.c_longlong(x.numel()), _lib.stream()),
               "cvd_relu_add")


def up2_bilinear(x, r, relu_r, out, align_corners):
    N, h, w, Cn = x.shape
    _lib.check(_lib.lib().cvd_up2_bilinear_fwd(_lib.ptr(x), _lib.ptr(r), 1 if relu_r else 0, N, h, w, Cn,
                                               1 if align_corners else 0, _lib.ptr(out), _lib.stream()), "cvd_up2_bilinear_fwd")


def up2_bilinear_bwd(dout, dx, align_corners, accumulate):
    N, h, w, Cn = dx.shape
    _lib.check(_lib.lib().cvd_up2_bilinear_bwd(_lib.ptr(dout), N, h, w, Cn, 1 if align_corners else 0, _lib.ptr(dx),
                                               1 if accumulate else 0, _lib.stream()), "cvd_up2_bilinear_bwd")


def recip_relu(raw4, depth):
    _lib.check(_lib.lib().cvd_recip_relu_fwd(_lib.ptr(raw4), _lib.ptr(depth), C.c_longlong(depth.numel()), _lib.stream()),
               "cvd_recip_relu_fwd")


def recip_relu_bwd(ddepth, depth, raw4, draw4):
    _lib.check(_lib.lib().cvd_recip_relu_bwd(_lib.ptr(ddepth), _lib.ptr(depth), _lib.ptr(raw4), _lib.ptr(draw4),
                                             C.c_longlong(depth.numel()), _lib.stream()), "cvd_recip_relu_bwd")


# ---------------------------------------------------------------- second-generation conv path (prep.cu, conv2.cu)
def z_alloc(N, C, H, W, device):
    """Operand planes of a C-channel tensor: [2 (hi | lo)][N][ceil16(C)/8][H*W][8] bf16 (see prep.cu)."""
    c8 = (C + 15) // 16 * 2
    return torch.empty(2, N, c8, H * W, 8, dtype=torch.bfloat16, device=device)


def prep_operand(src, Cn, z, zc8_off=0):
    """Transform + bf16 hi/lo split of the Cn logical channels of `src` (cvd_src_t from make_src) into chunks
    [zc8_off, zc8_off + ceil16(Cn)/8) of the planes z (from z_alloc)."""
    N, c8, HW = z.shape[1], z.shape[2], z.shape[3]
    _lib.check(_lib.lib().cvd_prep_operand(C.byref(src), int(Cn), C.c_longlong(N), C.c_longlong(HW), _lib.ptr(z), c8, zc8_off, 3,
                                           _lib.stream()), "cvd_prep_operand")


def conv2_tap_groups(cout_gemm, k):
    G, ng = C.c_int(0), C.c_int(0)
    _lib.check(_lib.lib().cvd_conv2_tap_groups(cout_gemm, k, C.byref(G), C.byref(ng)), "cvd_conv2_tap_groups")
    return G.value, ng.value


def conv2_packed_bytes(cin_gemm, cout_gemm, k):
    return int(_lib.lib().cvd_conv2_packed_bytes(cin_gemm, cout_gemm, k))


def make_pack2_table(entries, device):
    """entries: [(w_oihw tensor, packed uint8 tensor, flip)] -> device descriptor table for conv2_pack_batch."""
    import numpy as np
    dt = np.dtype([("w", "<u8"), ("out", "<u8"), ("cin", "<i4"), ("cout", "<i4"), ("k", "<i4"), ("flip", "<i4"),
                   ("G", "<i4"), ("ng", "<i4")])
    arr = np.zeros(len(entries), dtype=dt)
    for i, (w, out, flip) in enumerate(entries):
        cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
        G, ng = conv2_tap_groups(cin if flip else cout, k)
        arr[i] = (w.data_ptr(), out.data_ptr(), cin, cout, k, 1 if flip else 0, G, ng)
    t = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
    t._keep = [e[0] for e in entries] + [e[1] for e in entries]
    return t, len(entries)


def conv2_pack_batch(table, n):
    _lib.check(_lib.lib().cvd_conv2_pack_batch(_lib.ptr(table), n, _lib.stream()), "cvd_conv2_pack_batch")


def conv2_pack(w_oihw, flip=False):
    """One conv's weights -> conv2 tiles (tests / one-off use; engines pack all convs in one launch)."""
    cout, cin, k, _ = w_oihw.shape
    gi, go = (cout, cin) if flip else (cin, cout)
    out = torch.empty(conv2_packed_bytes(gi, go, k), dtype=torch.uint8, device=w_oihw.device)
    tab, n = make_pack2_table([(w_oihw, out, flip)], w_oihw.device)
    conv2_pack_batch(tab, n)
    return out


def conv2(z, zc8_off, packed, bias, dst, N, H, W, cin, cout, k, flags=0, bn=None):
    """TMA-fed kx-fused conv: z operand planes (z_alloc/prep_operand), dst cvd_dst_t; cin/cout in GEMM terms."""
    _lib.check(_lib.lib().cvd_conv2_fwd(_lib.ptr(z), z.shape[2], zc8_off, _lib.ptr(packed), _lib.ptr(bias), C.byref(dst),
                                        N, H, W, cin, cout, k, flags, C.byref(bn) if bn is not None else None,
                                        _lib.stream()), "cvd_conv2_fwd")


def conv2_wgrad(xz, x_off, gz, g_off, dw, N, H, W, cin, cout, k):
    """dw (fp32 OIHW, pre-zeroed or accumulating) += G (x) X on operand planes.  Returns False when the shape is not
    covered by the second-generation kernel (the caller then runs conv_wgrad on the fp32 views)."""
    rc = _lib.lib().cvd_conv2_wgrad(_lib.ptr(xz), xz.shape[2], x_off, _lib.ptr(gz), gz.shape[2], g_off, _lib.ptr(dw),
                                    N, H, W, cin, cout, k, _lib.stream())
    if rc == 2:
        return False
    _lib.check(rc, "cvd_conv2_wgrad")
    return True
