"""CPU: the per-element arithmetic of the flow-consistency-mask CUDA kernel (csrc/flow_mask_core.h, a __host__ __device__
function shared with flow_mask.cu) compiled with gcc and checked against the oracle and the reference-generated golden
masks -- the kernel was written without GPU time left, so this is its arithmetic check; only the launch indexing
(one thread per element) is left to tests/test_flowmask_gpu.py."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import flowmask_oracle as fo
from oracle.make_golden import FLOWMASK_CASES

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    out = str(tmp_path_factory.mktemp("emul") / "libflowmask_host.so")
    inc = os.path.join(os.path.dirname(HERE), "consistent_depth_b200", "csrc")
    subprocess.check_call([gcc, "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-I", inc, "-o", out,
                           os.path.join(HERE, "host_emul", "flow_mask_host.c"), "-lm"])
    return C.CDLL(out)


@pytest.mark.parametrize("name", list(FLOWMASK_CASES))
def test_flow_mask_core_matches_oracle_and_reference(host_lib, name):
    seed, H, W, ft, ct = FLOWMASK_CASES[name]
    flows, colors = fo.synthetic_pair(seed, H, W)
    f = np.ascontiguousarray(np.stack([x.transpose(2, 0, 1) for x in flows])[None], np.float32)       # (1,2,2,H,W)
    c = np.ascontiguousarray(np.stack([x.transpose(2, 0, 1) for x in colors])[None], np.float32)      # (1,2,3,H,W)
    m = np.zeros((1, 2, H, W), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    host_lib.flow_mask_host(p(f), p(c), p(m), 1, H, W, C.c_float(ft), C.c_float(ct))
    want = fo.consistent_flow_masks(flows, colors, ft, ct)
    g = np.load(os.path.join(HERE, "golden", "flowmask.npz"))
    for d in range(2):
        got = m[0, d] > 0.5
        assert (got != want[d]).mean() <= 2e-3 and (got != g[f"{name}_mask{d}"]).mean() <= 2e-3
        assert 0.05 < got.mean() < 0.95
