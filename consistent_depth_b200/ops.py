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
    return s


def make_dst(y):
    d = cvd_dst_t()
    d.y = y.t.data_ptr()
    d.c_total, d.c_off, d.n0, d.gap = y.c_total, y.off, y.n0, y.gap
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


def conv(src, packed, bias, dst, N, H, W, cin, cout, k, precision=3, flags=0):
    """src: cvd_src_t, dst: cvd_dst_t (from make_src/make_dst); cin/cout in GEMM terms."""
    _lib.check(_lib.lib().cvd_conv_fwd(C.byref(src), _lib.ptr(packed), _lib.ptr(bias), C.byref(dst),
                                       N, H, W, cin, cout, k, precision, flags, _lib.stream()), "cvd_conv_fwd")
