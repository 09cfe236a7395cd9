"""ctypes binding of libcvd_sm100.so (the C-ABI declared in include/cvd.h).

PyTorch is used only for device memory and streams: tensors are passed as raw
device pointers (`tensor.data_ptr()`) and the current CUDA stream handle.
The library is built in-tree by `__graft_entry__.build()`; if it is missing the
import of any kernel-backed op fails loudly — there is no fallback path.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVD_LIB", os.path.join(_HERE, "lib", "libcvd_sm100.so"))


class CvdError(RuntimeError):
    pass


class cvd_src_t(C.Structure):
    _fields_ = [("x", C.c_void_p), ("dy", C.c_void_p), ("a", C.c_void_p), ("b", C.c_void_p),
                ("bw", C.c_void_p),
                ("c_total", C.c_int), ("c_off", C.c_int), ("n0", C.c_int), ("gap", C.c_int),
                ("dy_ctotal", C.c_int), ("dy_coff", C.c_int), ("dy_n0", C.c_int), ("dy_gap", C.c_int),
                ("relu", C.c_int), ("mode", C.c_int)]


class cvd_dst_t(C.Structure):
    _fields_ = [("y", C.c_void_p), ("c_total", C.c_int), ("c_off", C.c_int), ("n0", C.c_int), ("gap", C.c_int)]


class cvd_bn_t(C.Structure):
    _fields_ = [("scratch", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p),
                ("a", C.c_void_p), ("b", C.c_void_p), ("rstd", C.c_void_p), ("mean", C.c_void_p),
                ("eps", C.c_float), ("momentum", C.c_float)]


_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises CvdError if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise CvdError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). consistent_depth_b200 has no CPU / eager fallback.")
        _lib = C.CDLL(LIB_PATH)
        _lib.cvd_last_error.restype = C.c_char_p
        _lib.cvd_launch_count.restype = C.c_longlong
        _lib.cvd_consistency_workspace_bytes.restype = C.c_size_t
        _lib.cvd_conv_packed_bytes.restype = C.c_size_t
        _lib.cvd_bn_scratch_bytes.restype = C.c_size_t
        _lib.cvd_conv2_packed_bytes.restype = C.c_size_t
    return _lib


def check(rc, what=""):
    if rc != 0:
        raise CvdError(f"{what}: {lib().cvd_last_error().decode()}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL). Tensor must be CUDA + contiguous."""
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise CvdError("expected a CUDA tensor: consistent_depth_b200 kernels have no CPU path")
    if not t.is_contiguous():
        raise CvdError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().cvd_launch_count())
