"""CPU: the static execution plans of the monodepth2 and MiDaS engines, built and replayed against a recording stand-in
for the C-ABI library (no kernels run).  Checks the host-side plumbing that a GPU is not needed for: the plan builds for
the fixture shapes, forward / backward / eval replay without Python errors, every conv of the architecture table is
launched, and every trainable conv weight receives a weight-gradient launch whose destination is that tensor's slice of
the flat gradient buffer (grouped / >256-channel convs: all their chunks)."""
import collections
import ctypes as C

import pytest
import torch


class _Recorder:
    def __init__(self, real):
        self.real, self.calls, self.args = real, collections.Counter(), collections.defaultdict(list)

    def __getattr__(self, name):
        if name in ("cvd_bn_scratch_bytes", "cvd_conv_packed_bytes", "cvd_conv2_packed_bytes", "cvd_correlation_out_size"):
            return getattr(self.real, name)

        def f(*a):
            self.calls[name] += 1
            self.args[name].append(a)
            return 0
        return f


@pytest.fixture
def fake_lib(monkeypatch):
    from consistent_depth_b200 import _lib
    rec = _Recorder(_lib.lib())
    monkeypatch.setattr(_lib, "lib", lambda: rec)
    monkeypatch.setattr(_lib, "ptr", lambda t: C.c_void_p(0) if t is None else C.c_void_p(t.data_ptr()))
    monkeypatch.setattr(_lib, "stream", lambda: C.c_void_p(0))
    monkeypatch.setenv("CVD_MULTI_STREAM", "0")
    return rec


def _wgrad_targets(rec):
    out = collections.Counter()
    for name, idx in (("cvd_conv_wgrad", 2), ("cvd_conv_wgrad_grouped", 2), ("cvd_conv_wgrad_grouped_chunks", 2),
                      ("cvd_conv2_wgrad", 6)):
        for a in rec.args[name]:
            out[a[idx].value] += 1
    return out


def _check_weight_coverage(P, arch, rec, chunk_rows=None):
    targets = _wgrad_targets(rec)
    for k, (off, shape) in P.pmap.items():
        if len(shape) != 4 or arch.dead_parameter(k):
            continue
        g = P._g(k)
        if chunk_rows and shape[1] * 32 == shape[0]:            # grouped (Cout, Cout/32, 3, 3): per-chunk launches only
            rows = range(0, shape[0], chunk_rows)
            assert all(targets[g[r:r + 1].data_ptr()] == 1 for r in rows), k
        else:
            assert targets[g.data_ptr()] == 1, k

    dead = [P._g(k).data_ptr() for k in P.pmap if arch.dead_parameter(k)]
    assert not any(targets[p] for p in dead)


def test_monodepth2_plan(fake_lib):
    from consistent_depth_b200.monodepth import mono2_arch
    from consistent_depth_b200.monodepth.mono2_engine import Mono2Engine, Mono2Params
    P = Mono2Params("cpu")
    e = Mono2Engine(P, 2, 24, 40, (64, 96))
    depth = e.forward(torch.rand(2, 3, 24, 40))
    assert depth.shape == (2, 24, 40)
    n_convs = sum(1 for k, s in mono2_arch.state_dict_shapes().items() if len(s) == 4 and not mono2_arch.dead_parameter(k))
    assert n_convs == 31 and fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == n_convs
    assert fake_lib.calls["cvd_conv_fwd_bn"] == 16           # stride-1 convs followed by BatchNorm: statistics in the epilogue
    assert fake_lib.calls["cvd_bn_stats"] == 4               # the four stride-2 convs followed by BatchNorm
    e.backward(torch.rand(2, 24, 40))
    assert fake_lib.calls["cvd_conv_wgrad"] == n_convs and fake_lib.calls["cvd_bn_bwd_reduce"] == 20
    assert fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == 2 * n_convs - 1    # no dgrad into the image
    _check_weight_coverage(P, mono2_arch, fake_lib)
    e.train_mode = False
    fake_lib.calls.clear()
    e.forward(torch.rand(2, 3, 24, 40))
    assert fake_lib.calls["cvd_conv_fwd_bn"] == 0 and fake_lib.calls["cvd_bn_stats"] == 0            # eval: running statistics
    assert P.num_batches_tracked == 1


def test_midas_plan(fake_lib):
    from consistent_depth_b200.monodepth import midas_arch
    from consistent_depth_b200.monodepth.midas_engine import CHUNK, MidasEngine, MidasParams
    P = MidasParams("cpu")
    e = MidasEngine(P, 2, 64, 96)
    assert e.forward(torch.rand(2, 3, 64, 96)).shape == (2, 64, 96)
    shapes = midas_arch.state_dict_shapes()
    dense = [k for k, s in shapes.items() if len(s) == 4 and not midas_arch.dead_parameter(k) and s[1] * 32 != s[0]]
    grouped = [k for k, s in shapes.items() if len(s) == 4 and s[1] * 32 == s[0]]
    chunks = sum(shapes[k][0] // CHUNK for k in grouped)
    assert len(grouped) == 33 and chunks == 508
    # every grouped conv is ONE chunked launch (blockIdx.y = 64-channel chunk) forward, one for dgrad, one for wgrad
    assert fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == len(dense)
    assert fake_lib.calls["cvd_conv_fwd_chunks"] == 33
    assert sum(a[13] for a in fake_lib.args["cvd_conv_fwd_chunks"]) == chunks
    e.backward(torch.rand(2, 64, 96))
    assert fake_lib.calls["cvd_conv_fwd_chunks"] == 66
    assert fake_lib.calls["cvd_conv_wgrad"] == len(dense) and fake_lib.calls["cvd_conv_wgrad_grouped_chunks"] == 33
    assert sum(a[7] for a in fake_lib.args["cvd_conv_wgrad_grouped_chunks"]) == chunks
    assert fake_lib.calls["cvd_bn_bwd_reduce"] == 104                                              # 104 BatchNorms
    _check_weight_coverage(P, midas_arch, fake_lib)


def test_flownet2_op_modules_plumbing(fake_lib):
    """Correlation / Resample2d / ChannelNorm mirrors: constructor surface of the reference modules, output shapes
    (cvd_correlation_out_size is host-only code and runs for real), argument order of the launches."""
    from consistent_depth_b200.third_party.flownet2.networks import ChannelNorm, Correlation, Resample2d
    a, b = torch.zeros(2, 16, 12, 17), torch.zeros(2, 16, 12, 17)
    out = Correlation(pad_size=20, kernel_size=1, max_displacement=20, stride1=1, stride2=2, corr_multiply=1)(a, b)
    assert out.shape == (2, 441, 12, 17)                                   # FlowNetC.py:28-31 configuration
    assert fake_lib.args["cvd_correlation_fwd"][0][3:12] == (2, 16, 12, 17, 20, 1, 20, 1, 2)
    assert Correlation(pad_size=4, kernel_size=3, max_displacement=4, stride1=2, stride2=1)(a, b).shape == (2, 81, 5, 8)
    assert Resample2d()(a, torch.zeros(2, 2, 12, 17)).shape == a.shape
    assert ChannelNorm()(a).shape == (2, 1, 12, 17)
    with pytest.raises(RuntimeError):
        Resample2d()(a.requires_grad_(True), torch.zeros(2, 2, 12, 17))


def test_mannequin_challenge_plan(fake_lib):
    """The hourglass engine's plan (the bench workload's network): launch counts of one forward + backward and
    weight-gradient coverage -- every conv weight of HourglassModel(3) that trains is the destination of exactly one
    wgrad launch (the four 1x1 convs of an inception block are ONE fused GEMM whose destination is the first of the four
    adjacent weights in the flat gradient buffer)."""
    from consistent_depth_b200.monodepth import mc_arch
    from consistent_depth_b200.monodepth.mc_engine import McEngine, McParams
    P = McParams("cpu")
    e = McEngine(P, 2, 32, 48)
    assert e.forward(torch.rand(2, 3, 32, 48)).shape == (2, 32, 48)
    n_inc = sum(1 for k in mc_arch.state_dict_shapes() if k.endswith(".convs.0.0.weight"))
    assert n_inc == 22
    # Dispatch: conv1 and the pred layer (3 / 1 channels) and the narrow-filter / many-output-channel k x k convs stay on
    # the first-generation per-tap kernel; the fused 1x1 convs and the k x k convs with few output channels or wide filters
    # run the TMA-fed kx-fused kernel.  Operands are prepared once: the block input (unless another block already prepared the same
    # tensor) and the 1x1 outputs a1|a2|a3.
    kk = [s_ for k_, s_ in mc_arch.state_dict_shapes().items() if ".convs." in k_ and k_.endswith(".3.weight")]
    assert len(kk) == 3 * n_inc
    use2 = lambda kg, ng, k: ng <= 16 or (ng <= 32 and k >= 7) or (k >= 11 and kg >= 64 and ng <= kg)   # McEngine._use_conv2
    fwd_v2 = sum(1 for s_ in kk if use2(s_[1], s_[0], s_[2]))      # (Cout, Cin, k, k): forward GEMM K = Cin, N = Cout
    dgrad_v2 = sum(1 for s_ in kk if use2(s_[0], s_[1], s_[2]))    # dgrad GEMM K = Cout, N = Cin
    assert 0 < fwd_v2 < len(kk) and 0 < dgrad_v2 < len(kk)
    assert fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == 2 + len(kk) - fwd_v2
    assert fake_lib.calls["cvd_conv2_fwd"] == n_inc + fwd_v2
    assert n_inc < fake_lib.calls["cvd_prep_operand"] <= 2 * n_inc
    assert fake_lib.calls["cvd_conv2_pack_batch"] == 1
    assert fake_lib.calls["cvd_bn_stats"] == 0                 # statistics come from the conv epilogues
    n_prep_fwd, n_v1_fwd = fake_lib.calls["cvd_prep_operand"], fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"]
    e.backward(torch.rand(2, 32, 48))
    # weight gradients: the fused 1x1 convs and the wide / few-channel k x k convs on the operand planes (cvd_conv2_wgrad),
    # the rest (and conv1 / pred) on the fp32 views
    wg_v2 = sum(1 for s_ in kk if not s_[1] < s_[0])             # McEngine._use_wgrad2
    assert 0 < wg_v2 < len(kk)
    assert fake_lib.calls["cvd_conv2_wgrad"] == n_inc + wg_v2 and fake_lib.calls["cvd_conv_wgrad"] == 2 + len(kk) - wg_v2
    # backward: one gradient-operand preparation per BatchNorm group (k x k outputs, 1x1 outputs); conv1 has no input gradient
    assert fake_lib.calls["cvd_prep_operand"] - n_prep_fwd == 2 * n_inc
    assert fake_lib.calls["cvd_conv2_fwd"] == n_inc + fwd_v2 + n_inc + dgrad_v2
    assert fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == n_v1_fwd + 1 + len(kk) - dgrad_v2
    targets = _wgrad_targets(fake_lib)
    for k, (off, shape) in P.pmap.items():
        if len(shape) != 4:
            continue
        g = P._g(k)
        if ".convs." in k and k.endswith(".0.weight") and not k.endswith(".convs.0.0.weight"):
            assert targets[g.data_ptr()] == 0, k               # 1x1 of branches 1..3: inside the fused GEMM of convs.0.0
        else:
            assert targets[g.data_ptr()] == 1, k


def test_mannequin_challenge_plan_first_generation_kernels(fake_lib, monkeypatch):
    """CVD_CONV2=0: every conv on the first-generation kernel (the fallback kept for A/B measurements)."""
    from consistent_depth_b200.monodepth.mc_engine import McEngine, McParams
    monkeypatch.setenv("CVD_CONV2", "0")
    e = McEngine(McParams("cpu"), 2, 32, 48)
    e.forward(torch.rand(2, 3, 32, 48))
    assert fake_lib.calls["cvd_conv_fwd"] + fake_lib.calls["cvd_conv_fwd_bn"] == 1 + 4 * 22 + 1
    assert fake_lib.calls["cvd_conv2_fwd"] == 0 and fake_lib.calls["cvd_prep_operand"] == 0
