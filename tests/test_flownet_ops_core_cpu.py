"""CPU: the per-element arithmetic of the FlowNet2 custom-op kernels (csrc/flownet_ops_core.h, shared with flownet_ops.cu)
compiled with gcc and checked against oracle/flownet_ops_oracle.py (which works on explicit zero-padded copies like the
reference kernels; the CUDA code works on the unpadded tensors with bounds tests -- two independent formulations)."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import flownet_ops_oracle as fo
from oracle import synth

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    out = str(tmp_path_factory.mktemp("emul") / "libflownet_host.so")
    inc = os.path.join(os.path.dirname(HERE), "consistent_depth_b200", "csrc")
    subprocess.check_call([gcc, "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I", inc, "-o", out,
                           os.path.join(HERE, "host_emul", "flownet_ops_host.c"), "-lm"])
    return C.CDLL(out)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("pad,K,md,s1,s2,H,W", [(20, 1, 20, 1, 2, 12, 17),      # FlowNetC.py:28-31
                                                (4, 3, 4, 1, 1, 9, 11), (6, 1, 4, 2, 2, 10, 14)])
def test_correlation_core(host_lib, pad, K, md, s1, s2, H, W):
    B, Cc = 2, 24
    a, b = synth.uniform(1, 1, (B, Cc, H, W), -1, 1), synth.uniform(1, 2, (B, Cc, H, W), -1, 1)
    want = fo.correlation(a, b, pad, K, md, s1, s2)
    out = np.zeros(want.shape, np.float32)
    host_lib.correlation_host(p(a), p(b), p(out), C.c_longlong(out.size), Cc, H, W, want.shape[2], want.shape[3], pad, K, md, s1, s2)
    np.testing.assert_allclose(out, want, rtol=1e-4, atol=1e-6)
    assert want.shape[1] == (2 * (md // s2) + 1) ** 2 and np.abs(want).max() > 1e-3


def test_resample2d_core(host_lib):
    B, Cc, H, W = 2, 5, 11, 13
    x = synth.uniform(2, 1, (B, Cc, H, W), -1, 1)
    flow = synth.normal(2, 2, (B, 2, H, W), 3.0)            # many samples leave the image: border replication path
    want = fo.resample2d(x, flow)
    out = np.zeros(x.shape, np.float32)
    host_lib.resample2d_host(p(x), p(flow), p(out), C.c_longlong(out.size), Cc, H, W)
    np.testing.assert_allclose(out, want, rtol=1e-4, atol=1e-5)
    # zero flow is the identity
    host_lib.resample2d_host(p(x), p(np.zeros_like(flow)), p(out), C.c_longlong(out.size), Cc, H, W)
    np.testing.assert_array_equal(out, x)


def test_channelnorm_core(host_lib):
    B, Cc, H, W = 2, 7, 6, 9
    x = synth.uniform(3, 1, (B, Cc, H, W), -2, 2)
    out = np.zeros((B, 1, H, W), np.float32)
    host_lib.channelnorm_host(p(x), p(out), C.c_longlong(out.size), Cc, H, W)
    np.testing.assert_allclose(out, fo.channelnorm(x), rtol=1e-5)
