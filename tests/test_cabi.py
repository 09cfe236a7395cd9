"""CPU: the C-ABI library builds, loads and exports every symbol include/cvd.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    return ctypes.CDLL(g.LIB)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cvd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("cvd_consistency_fwd_bwd", "cvd_mask_sums", "cvd_adam_flat", "cvd_version", "cvd_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    missing = [s for s in declared_symbols() if not hasattr(built_lib, s)]
    assert not missing, f"declared in include/cvd.h but not exported: {missing}"


def test_version_and_error_string(built_lib):
    assert built_lib.cvd_version() == 100
    built_lib.cvd_last_error.restype = ctypes.c_char_p
    assert isinstance(built_lib.cvd_last_error(), bytes)


def test_argument_validation_needs_no_gpu(built_lib):
    # NULL pointers are rejected before any CUDA call is made
    rc = built_lib.cvd_mask_sums(None, None, 1, 4, 4, None, None)
    assert rc != 0
    built_lib.cvd_last_error.restype = ctypes.c_char_p
    assert b"null" in built_lib.cvd_last_error()


def test_bn_scratch_holds_one_block_per_256_channels(built_lib):
    """Chunked launches (blockIdx.y / .z = 256-channel chunk) reduce their chunks concurrently, each into its own
    [2 x 256 sums | ticket] block of the statistics scratch; the engines size every BatchNorm's scratch for one block per
    64-channel chunk of a grouped conv (midas_engine.CHUNK)."""
    built_lib.cvd_bn_scratch_bytes.restype = ctypes.c_size_t
    block = (2 * 256 + 1) * 8
    assert built_lib.cvd_bn_scratch_bytes(1) == block and built_lib.cvd_bn_scratch_bytes(256) == block
    assert built_lib.cvd_bn_scratch_bytes(257) == 2 * block and built_lib.cvd_bn_scratch_bytes(2048) == 8 * block
    from consistent_depth_b200.monodepth import midas_arch
    from consistent_depth_b200.monodepth.midas_engine import CHUNK
    widths = {s[0] for k, s in midas_arch.state_dict_shapes().items() if len(s) == 4 and s[1] * 32 == s[0]}
    for w in widths:            # _BN allocates cvd_bn_scratch_bytes(max(256, 4 * C)): >= one block per CHUNK channels
        assert built_lib.cvd_bn_scratch_bytes(max(256, 4 * w)) >= (w // CHUNK) * block


def test_chunked_entry_points_validate_arguments_without_gpu(built_lib):
    built_lib.cvd_last_error.restype = ctypes.c_char_p
    assert built_lib.cvd_conv_fwd_chunks(None, None, None, None, 1, 8, 8, 64, 64, 3, 3, 0, None, 4, 64, 64,
                                         ctypes.c_longlong(0), None) != 0
    assert b"null" in built_lib.cvd_last_error()
    assert built_lib.cvd_conv_wgrad_grouped_chunks(None, None, None, 1, 8, 8, 64, 4, 8, 3, 3, None) != 0
    assert b"null" in built_lib.cvd_last_error()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "consistent_depth_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports oracle/"


def test_ops_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from consistent_depth_b200 import _lib
    from consistent_depth_b200.utils.geometry import fused_consistency
    z = torch.zeros(1, 2, 4, 4)
    with pytest.raises((_lib.CvdError, RuntimeError, AssertionError)):
        fused_consistency(z, [z, z], [z[:, :1], z[:, :1]], torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 4), 1.0, 0.1)
