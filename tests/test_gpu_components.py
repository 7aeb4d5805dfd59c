"""GPU parity at the component boundary (L2) and the member boundary (L1) through the C ABI
of include/kcnn_capi.h, against the CPU oracle.  Reads like the reference's own (disabled)
nnet-conv-test.cc: build a component from a config line, Propagate, Backprop with
to_update == the component, Write / Read round trip -- but with exact expected values
from the oracle instead of finite-difference bounds.

Tolerance: 1e-5 relative (max-norm) for FP32, 1e-3 for TF32; max pooling bit-exact.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import dev, dev_empty, host, rel_err, assert_bit_exact  # noqa: E402
from kaldi_cnn_b200 import components as kc  # noqa: E402

TOL = {0: 1e-5, 1: 1e-3}
HERE = os.path.dirname(os.path.abspath(__file__))

C1A = ("ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=40 kernel-width=4 stride=1 "
       "group=128 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5")
C1B = ("ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=8 kernel-width=3 stride=1 "
       "group=64 out-height=33 out-width=9 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5")
CPAD = ("ConvolutionComponent in-height=6 in-width=7 in-channel=5 in-pad-height=1 in-pad-width=2 kernel-height=3 "
        "kernel-width=4 stride=1 group=10 out-height=6 out-width=8 learning-rate=0.05 param-stddev=0.1 bias-stddev=0.5")
CTIME = ("ConvolutionComponent in-height=1 in-width=14 in-channel=64 kernel-height=1 kernel-width=3 stride=1 "
         "group=128 out-height=1 out-width=12 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5")
C5 = ("ConvolutionComponent in-height=1 in-width=6 in-channel=256 kernel-height=1 kernel-width=3 stride=1 "
      "group=512 out-height=1 out-width=4 learning-rate=0.02 param-stddev=0.02 bias-stddev=0.5")
CONV_LINES = {"C1a": (C1A, (40, 11, 3, 0, 0, 40, 4, 128), 256), "C1b": (C1B, (40, 11, 3, 0, 0, 8, 3, 64), 32),
              # conv5 of nnet.config at N = 128: the input-gradient GEMM AND the weight-gradient GEMM are both
              # split over K, and they run as two concurrent branches (separate split-K scratch!)
              "conv5": (C5, (1, 6, 256, 0, 0, 1, 3, 512), 128),
              "pad": (CPAD, (6, 7, 5, 1, 2, 3, 4, 10), 9),
              "time": (CTIME, (1, 14, 64, 0, 0, 1, 3, 128), 70)}     # nnet.config-style layer: TMA path + staging


def _get(t):
    return t.detach().cpu().numpy().copy()


@pytest.mark.parametrize("name", list(CONV_LINES))
@pytest.mark.parametrize("math", [0, 1])
def test_convolution_component_two_training_steps(ora, name, math):
    line, (H, W, C, ph, pw, KH, KW, G), N = CONV_LINES[name]
    kc.set_math_mode(math)
    kc.set_rand_seed(42)
    comp = kc.Component.from_string(line)
    assert comp.type == "ConvolutionComponent"
    assert comp.input_dim == H * W * C
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    assert comp.output_dim == OH * OW * G
    # App. C.2: config weight-decay/momentum are swallowed; defaults 0.0002 / 0.9 apply
    wd, mom = comp.weight_decay_momentum()
    assert abs(wd - 0.0002) < 1e-9 and abs(mom - 0.9) < 1e-7
    lr = 0.02 if name != "pad" else 0.05
    lin, bias, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    assert lin.shape == (KH * KW * C, G) and bias.shape == (G,) and not prev.any()
    rng = np.random.default_rng(1234)
    for step in range(2):
        x = rng.standard_normal((N, H * W * C)).astype(np.float32)
        dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
        xd, dyd = dev(x, 4, 0), dev(dy, 0, 0)
        y = comp.propagate(xd)
        y_ref = ora.conv_propagate(lin, None, None, 0, 0, 0, 0, 0, 0, 0, 0) if False else \
            ora.conv_propagate(x, lin, bias, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
        assert rel_err(_get(y), y_ref) <= TOL[math]
        dx = comp.backprop(xd, None, dyd, update=True)
        dx_ref = ora.conv_backprop(dy, lin, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
        assert rel_err(_get(dx), dx_ref) <= TOL[math]
        lin_r, bias_r, prev_r, g_r, bg_r = ora.conv_update(x, dy, lin, bias, prev, H, W, C, ph, pw, KH, KW, G,
                                                           lr, wd, mom, dtype=np.float64)
        scale = max(np.abs(lin_r - lin).max(), 1e-30)      # judge the UPDATE, not the weights' magnitude
        assert np.abs(_get(comp.params(0)) - lin_r).max() <= TOL[math] * scale * 4
        assert rel_err(_get(comp.params(2)), prev_r) <= TOL[math] * 4
        assert np.abs(_get(comp.params(1))[0] - bias_r).max() <= 1e-5 * max(np.abs(bias_r - bias).max(), 1e-30) * 4
        lin, bias, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    kc.set_math_mode(0)


@pytest.mark.parametrize("name", ["conv5", "time", "C1a"])
def test_convolution_deferred_gradient(ora, name):
    """Data-parallel mode (update deferred): Backprop leaves the un-normalised dK, db of
    nnet0/nnet-component-nnet0.cc:763, 775 in the gradient buffers and the right in_deriv."""
    line, (H, W, C, ph, pw, KH, KW, G), N = CONV_LINES[name]
    kc.set_math_mode(1)
    kc.set_rand_seed(42)
    comp = kc.Component.from_string(line)
    comp.set_deferred_update(True)
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    lin, bias, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    rng = np.random.default_rng(77)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
    xd, dyd = dev(x, 4, 0), dev(dy, 0, 0)
    comp.propagate(xd)
    dx = comp.backprop(xd, None, dyd, update=True)
    dx_ref = ora.conv_backprop(dy, lin, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
    assert rel_err(_get(dx), dx_ref) <= 1e-3
    g_r, bg_r = ora.conv_update(x, dy, lin, bias, prev, H, W, C, ph, pw, KH, KW, G, 0.02, 0.0002, 0.9,
                                dtype=np.float64)[3:5]
    assert rel_err(_get(comp.gradient(0)), g_r) <= 1e-3
    assert rel_err(_get(comp.gradient(1))[0], bg_r) <= 1e-5 * 8
    assert np.array_equal(_get(comp.params(0)), lin)          # nothing applied yet
    kc.set_math_mode(0)


def test_backprop_with_another_in_value_does_not_reuse_the_staging_copy(ora):
    """A bare Component must not assume Backprop's in_value is what was last propagated (the two
    device buffers here may even share an address through the caching allocator): the
    channels-last staging copy is reused only under NnetMinibatchUpdater (SetInputPersists)."""
    kc.set_math_mode(1)
    kc.set_rand_seed(7)
    comp = kc.Component.from_string(CTIME)
    H, W, C, ph, pw, KH, KW, G = CONV_LINES["time"][1]
    N = 40
    rng = np.random.default_rng(9)
    x1 = rng.standard_normal((N, W * C)).astype(np.float32)
    x2 = rng.standard_normal((N, W * C)).astype(np.float32)
    dy = rng.standard_normal((N, 12 * G)).astype(np.float32)
    lin, bias, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    wd, mom = comp.weight_decay_momentum()
    comp.propagate(dev(x1))
    comp.backprop(dev(x2), None, dev(dy), update=True)
    lin_r = ora.conv_update(x2, dy, lin, bias, prev, H, W, C, ph, pw, KH, KW, G, 0.02, wd, mom, dtype=np.float64)[0]
    scale = max(np.abs(lin_r - lin).max(), 1e-30)
    assert np.abs(_get(comp.params(0)) - lin_r).max() <= 1e-3 * scale * 4
    # same tensor propagated and back-propagated:
    xd = dev(x1)
    comp.propagate(xd)
    lin = _get(comp.params(0)); prev = _get(comp.params(2)); bias = _get(comp.params(1))[0]
    comp.backprop(xd, None, dev(dy), update=True)
    lin_r = ora.conv_update(x1, dy, lin, bias, prev, H, W, C, ph, pw, KH, KW, G, 0.02, wd, mom, dtype=np.float64)[0]
    scale = max(np.abs(lin_r - lin).max(), 1e-30)
    assert np.abs(_get(comp.params(0)) - lin_r).max() <= 1e-3 * scale * 4
    kc.set_math_mode(0)


def test_convolution_backprop_without_update_leaves_params(ora):
    kc.set_math_mode(0)
    comp = kc.Component.from_string(CPAD)
    H, W, C, ph, pw, KH, KW, G = CONV_LINES["pad"][1]
    before = _get(comp.params(0))
    rng = np.random.default_rng(5)
    x = dev(rng.standard_normal((4, H * W * C)).astype(np.float32))
    dy = dev(rng.standard_normal((4, comp.output_dim)).astype(np.float32))
    comp.backprop(x, None, dy, update=False)
    assert np.array_equal(before, _get(comp.params(0)))


POOL_LINES = [
    ("MaxpoolComponent in-height=1 in-width=8 in-channel=128 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=2",
     (1, 8, 128, 1, 2, 2), 256),
    ("MaxpoolComponent in-height=33 in-width=9 in-channel=64 pool-height-dim=3 pool-width-dim=3 pool-channel-dim=2",
     (33, 9, 64, 3, 3, 2), 16),
    ("MaxpoolComponent in-height=1 in-width=12 in-channel=256 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=1",
     (1, 12, 256, 1, 2, 1), 64),
]


@pytest.mark.parametrize("line,geom,N", POOL_LINES)
@pytest.mark.parametrize("relu_input", [False, True])
def test_maxpool_component_bit_exact(ora, line, geom, N, relu_input):
    H, W, C, ph, pw, pc = geom
    comp = kc.Component.from_string(line)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    if relu_input:
        x = np.maximum(x, 0)          # ties: whole windows of zeros
    y_ref = ora.maxpool_prop(x, H, W, ph, pw, pc)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    xd = dev(x, 4, 0)
    y = comp.propagate(xd)
    assert_bit_exact(_get(y), y_ref, "MaxpoolComponent::Propagate")
    dx = comp.backprop(xd, y, dev(dy), update=False)
    assert_bit_exact(_get(dx), ora.maxpool_backprop(x, y_ref, dy, H, W, ph, pw, pc), "MaxpoolComponent::Backprop")
    # index routing == reference routing when maxima are unique
    if not relu_input:
        comp.set_index_routing(True)
        y2 = comp.propagate(xd)
        assert_bit_exact(_get(y2), y_ref, "index-mode Propagate")
        dx2 = comp.backprop(xd, y2, dev(dy), update=False)
        assert_bit_exact(_get(dx2), _get(dx), "index-mode Backprop")
    # serialisation round trip, both modes
    for binary in (True, False):
        c2 = kc.Component.read(comp.write(binary), binary)
        assert c2.type == "MaxpoolComponent" and c2.output_dim == comp.output_dim
        assert c2.write(binary) == comp.write(binary)


def test_maxpool_overlap_component(ora):
    comp = kc.Component.from_string("MaxpoolComponent in-height=1 in-width=4 in-channel=12 pool-height-dim=1 "
                                    "pool-width-dim=1 pool-channel-dim=3 overlap=true")
    assert comp.output_dim == 4 * 10
    rng = np.random.default_rng(8)
    x = np.maximum(rng.standard_normal((6, 48)), 0).astype(np.float32)
    y_ref = ora.maxpool_prop(x, 1, 4, 1, 1, 3, mode=1)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    xd = dev(x)
    y = comp.propagate(xd)
    assert_bit_exact(_get(y), y_ref, "overlap Propagate")
    dx = comp.backprop(xd, y, dev(dy), update=False)
    assert_bit_exact(_get(dx), ora.maxpool_backprop(x, y_ref, dy, 1, 4, 1, 1, 3, mode=1), "overlap Backprop")


@pytest.mark.parametrize("math", [0, 1])
def test_fully_connected_component_two_steps(ora, math):
    """Propagate / Backprop / UpdateSimple (nnet2/nnet-component.cc:1216-1258, nnet0/nnet-component-nnet0.cc:
    1133-1143) over two steps (the second exercises the momentum) against the oracle.  The update's order of
    operations (AddRowSumMat for the bias, Scale + AddMat + AddMatMat for prev_grad_, AddMat for W) is pinned by
    the oracle's restatement only: those stock Kaldi routines are not in /root/reference, unlike the cnslmat
    kernels the convolution chains are checked against (tests/test_gpu_reference_chains.py)."""
    kc.set_math_mode(math)
    kc.set_rand_seed(43)
    comp = kc.Component.from_string("FullyConnectedComponent input-dim=256 output-dim=1024 learning-rate=0.02 "
                                    "param-stddev=0.01 bias-stddev=1 weight-decay=0.0005 momentum=0.9")
    assert comp.type == "FullyConnectedComponent"
    wd, mom = comp.weight_decay_momentum()
    assert abs(wd - 0.0005) < 1e-9       # FC applies the config values (unlike conv)
    Wm, b, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    assert Wm.shape == (1024, 256) and np.all(b == 1.0) and not prev.any()
    rng = np.random.default_rng(9)
    N = 256
    for _ in range(2):
        x = rng.standard_normal((N, 256)).astype(np.float32)
        dy = rng.standard_normal((N, 1024)).astype(np.float32)
        xd, dyd = dev(x), dev(dy, 4, 0)
        y = comp.propagate(xd)
        assert rel_err(_get(y), ora.fc_propagate(x, Wm, b, dtype=np.float64)) <= TOL[math]
        dx = comp.backprop(xd, None, dyd, update=True)
        assert rel_err(_get(dx), ora.fc_backprop(dy, Wm, dtype=np.float64)) <= TOL[math]
        W_r, b_r, p_r = ora.fc_update(x, dy, Wm, b, prev, 0.02, wd, mom, dtype=np.float64)
        assert np.abs(_get(comp.params(0)) - W_r).max() <= TOL[math] * np.abs(W_r - Wm).max() * 4
        assert rel_err(_get(comp.params(2)), p_r) <= TOL[math] * 4
        assert np.abs(_get(comp.params(1))[0] - b_r).max() <= 1e-5 * np.abs(b_r - b).max() * 4
        Wm, b, prev = _get(comp.params(0)), _get(comp.params(1))[0], _get(comp.params(2))
    kc.set_math_mode(0)


@pytest.mark.parametrize("line", [C1B, "FullyConnectedComponent input-dim=30 output-dim=20 learning-rate=0.02 "
                                  "param-stddev=0.1 bias-stddev=1 weight-decay=0.0005 momentum=0.9"])
@pytest.mark.parametrize("binary", [True, False])
def test_updatable_component_write_read_round_trip(line, binary):
    comp = kc.Component.from_string(line)
    comp.params(2).normal_()               # non-trivial momentum state must survive
    blob = comp.write(binary)
    assert blob.startswith(("<%s> " % comp.type).encode())
    c2 = kc.Component.read(blob, binary)
    assert c2.type == comp.type
    for which in (0, 1, 2):
        a, b = _get(comp.params(which)), _get(c2.params(which))
        if binary:
            assert np.array_equal(a, b)
        else:
            assert np.allclose(a, b, rtol=2e-6, atol=0)
    if binary:
        assert c2.write(True) == blob
    c3 = comp.copy()
    assert np.array_equal(_get(c3.params(2)), _get(comp.params(2)))


def test_config_errors_are_reported_not_fatal():
    with pytest.raises(kc.KcnnError, match="Could not process these elements"):
        kc.Component.from_string(C1A + " bogus-key=3")
    with pytest.raises(kc.KcnnError):
        kc.Component.from_string("MaxpoolComponent in-height=3 in-width=8 in-channel=4 pool-height-dim=2 "
                                 "pool-width-dim=2 pool-channel-dim=2")       # 3 % 2 != 0
    with pytest.raises(kc.KcnnError, match="stride"):
        kc.Component.from_string(C1A.replace("stride=1", "stride=2").replace("out-width=8", "out-width=4"))
    with pytest.raises(kc.KcnnError, match="no such type"):
        kc.Component.from_string("NoSuchComponent dim=3")


def test_l1_members_through_c_api(ora):
    """Conv2D / TpBlock / FlipMat member semantics incl. the 'out is never resized' contract."""
    from kaldi_cnn_b200.components import _lib, _mat, _check
    L = _lib()
    kc.use_current_stream()
    rng = np.random.default_rng(11)
    N, H, W, C, KH, KW, G = 6, 5, 7, 3, 2, 3, 4
    OH, OW = H - KH + 1, W - KW + 1
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    xd, kd = dev(x, 3, 1), dev(k)
    for concat in (1, 0):
        ref = ora.conv2d(x, k, H, W, C, KH, KW, G, concat=bool(concat), dtype=np.float64)
        od = dev_empty(*ref.shape)
        _check(L.kcnn_mat_conv2d(*_mat(xd), *_mat(kd), H, W, C, KH, KW, G, *_mat(od), concat))
        assert rel_err(host(od), ref) <= 1e-5
    bad = dev_empty(N, 5)
    assert L.kcnn_mat_conv2d(*_mat(xd), *_mat(kd), H, W, C, KH, KW, G, *_mat(bad), 1) == -1
    assert b"KALDI_ASSERT" in L.kcnn_last_error()
    od = dev_empty(C, N * H * W)
    _check(L.kcnn_mat_tp_block(*_mat(xd), C, H * W, *_mat(od)))
    assert_bit_exact(host(od), ora.tp_block(x, C, H * W), "TpBlock member")
    fd = dev_empty(KH * KW * G, C)
    _check(L.kcnn_mat_flip_mat(*_mat(kd), KH, KW, C, G, *_mat(fd)))
    assert_bit_exact(host(fd), ora.flip_mat(k, KH, KW, C, G), "FlipMat member")
    wrong = dev_empty(KH * KW * G + 1, C)
    assert L.kcnn_mat_flip_mat(*_mat(kd), KH, KW, C, G, *_mat(wrong)) == -1
