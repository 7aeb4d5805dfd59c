"""CPU: the op-for-op C restatement (oracle/kcnn_oracle_impl.h) against the
independent einsum formulation (oracle/oracle_np.py) in FP64 and FP32.

This is the permanent form of the cross-check SURVEY Appendix A describes.
"""
import numpy as np
import pytest

from oracle import oracle_np as onp

# (N, H, W, C, pad_h, pad_w, KH, KW, G)
SHAPES = [
    (4, 40, 11, 3, 0, 0, 40, 4, 16),    # C1a-like: time-axis conv, OH = 1 (no-flip dgrad branch)
    (3, 40, 11, 3, 0, 0, 8, 3, 8),      # C1b-like: true 2-D (flip branch)
    (5, 1, 18, 16, 0, 0, 1, 3, 12),     # H = 1 layers of nnet.config
    (2, 6, 7, 2, 1, 1, 3, 3, 5),        # zero padding both axes
    (3, 5, 9, 4, 0, 2, 5, 4, 6),        # padding along width only
    (1, 1, 4, 8, 0, 0, 1, 3, 8),        # conv6-like, OW = 2
]


def _data(shape, dtype, seed=0):
    N, H, W, C, ph, pw, KH, KW, G = shape
    rng = np.random.default_rng(seed)
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    x = rng.standard_normal((N, H * W * C)).astype(dtype)
    k = (rng.standard_normal((KH * KW * C, G)) * 0.1).astype(dtype)
    b = rng.standard_normal(G).astype(dtype)
    dy = rng.standard_normal((N, OH * OW * G)).astype(dtype)
    return x, k, b, dy


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_fprop_f64(ora, shape):
    N, H, W, C, ph, pw, KH, KW, G = shape
    x, k, b, _ = _data(shape, np.float64)
    got = ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
    ref = onp.conv_fprop(x, k, b, H, W, C, ph, pw, KH, KW, G)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("branch", [0, 1])
def test_conv_dgrad_both_branches_f64(ora, shape, branch):
    N, H, W, C, ph, pw, KH, KW, G = shape
    x, k, b, dy = _data(shape, np.float64, seed=1)
    got = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G, branch=branch, dtype=np.float64)
    ref = onp.conv_dgrad(dy, k, H, W, C, ph, pw, KH, KW, G)
    assert np.abs(got - ref).max() < 1e-12


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_wgrad_and_sgd_f64(ora, shape):
    N, H, W, C, ph, pw, KH, KW, G = shape
    x, k, b, dy = _data(shape, np.float64, seed=2)
    rng = np.random.default_rng(3)
    prev = rng.standard_normal(k.shape) * 0.01
    lr, wd, mom = 0.02, 0.0005, 0.9
    k2, b2, p2, dk, db = ora.conv_update(x, dy, k, b, prev, H, W, C, ph, pw, KH, KW, G,
                                          lr, wd, mom, dtype=np.float64)
    rdk, rdb = onp.conv_wgrad(x, dy, H, W, C, ph, pw, KH, KW, G)
    assert np.abs(dk - rdk).max() < 1e-11
    assert np.abs(db - rdb).max() < 1e-11
    rk, rb, rp = onp.sgd(k, b, prev, rdk, rdb, N, lr, wd, mom)
    # lr is rounded through float in the reference (learning_rate_/num_sample is BaseFloat)
    assert np.abs(k2 - rk).max() < 1e-7 * max(1.0, np.abs(rdk).max())
    assert np.abs(b2 - rb).max() < 1e-7 * max(1.0, np.abs(rdb).max())
    assert np.abs(p2 - rp).max() < 1e-7 * max(1.0, np.abs(rdk).max())


def test_backprop_branch_rule_matches_survey(ora):
    # nnet.config conv1 and conv6 take the no-flip branch, conv2-5 flip;
    # C1a no-flip, C1b flip.
    assert not ora.conv_backprop_uses_flip(0, 0, 40, 4, 1, 18)
    assert not ora.conv_backprop_uses_flip(0, 0, 1, 3, 1, 2)
    for ow in (16, 14, 12, 4):
        assert ora.conv_backprop_uses_flip(0, 0, 1, 3, 1, ow)
    assert not ora.conv_backprop_uses_flip(0, 0, 40, 4, 1, 8)
    assert ora.conv_backprop_uses_flip(0, 0, 8, 3, 33, 9)


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_f32_close_to_f64(ora, shape):
    """Budget for the FP32 tolerance: the float restatement against the double one."""
    N, H, W, C, ph, pw, KH, KW, G = shape
    x, k, b, dy = _data(shape, np.float32, seed=4)
    y32 = ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G)
    y64 = ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
    assert np.abs(y32 - y64).max() <= 1e-5 * np.abs(y64).max()
    d32 = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G)
    d64 = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
    assert np.abs(d32 - d64).max() <= 1e-5 * np.abs(d64).max()


MP_SHAPES = [
    # (N, H, W, C, ph, pw, pc)
    (4, 1, 8, 128, 1, 2, 2),     # C1a pool
    (3, 33, 9, 64, 3, 3, 2),     # C1b pool 3x3x2
    (2, 1, 12, 256, 1, 2, 1),    # nnet.config time pool
    (2, 1, 1, 40, 1, 1, 5),      # pure intermap pooling
    (1, 4, 6, 6, 2, 3, 3),
]


@pytest.mark.parametrize("shape", MP_SHAPES)
def test_maxpool_bit_exact_vs_reshape(ora, shape):
    N, H, W, C, ph, pw, pc = shape
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    y = ora.maxpool_prop(x, H, W, ph, pw, pc)
    assert np.array_equal(y, onp.maxpool_fwd(x, H, W, C, ph, pw, pc))
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dx = ora.maxpool_backprop(x, y, dy, H, W, ph, pw, pc)
    assert np.array_equal(dx, onp.maxpool_bwd_ties(x, y, dy, H, W, C, ph, pw, pc))


def test_maxpool_ties_route_to_all_equal(ora):
    # post-ReLU zeros: a window of all zeros sends err to every element (cnsl-cu-kernels.cu:302-303)
    N, H, W, C, ph, pw, pc = 2, 1, 8, 16, 1, 2, 2
    rng = np.random.default_rng(6)
    x = np.maximum(rng.standard_normal((N, H * W * C)), 0).astype(np.float32)
    y = ora.maxpool_prop(x, H, W, ph, pw, pc)
    dy = np.ones_like(y)
    dx = ora.maxpool_backprop(x, y, dy, H, W, ph, pw, pc)
    assert dx.sum() > dy.sum()          # more routed elements than outputs
    assert np.array_equal(dx, onp.maxpool_bwd_ties(x, y, dy, H, W, C, ph, pw, pc))


def test_maxpool_sentinel_nan_and_signed_zero(ora):
    H, W, C, ph, pw, pc = 1, 2, 2, 1, 2, 2
    x = np.array([
        [np.nan, 1.0, 2.0, np.nan],          # NaN never wins (val < src is false)
        [-1e30, -2e30, -3e30, -4e30],        # all below the -1e20 sentinel -> -1e20
        [-0.0, 0.0, 0.0, -0.0],              # first seen zero is kept: -0.0
        [0.0, -0.0, -0.0, 0.0],              # first seen zero is kept: +0.0
        [np.nan, np.nan, np.nan, np.nan],    # all NaN -> sentinel
        [-np.inf, np.inf, 1.0, 2.0],
    ], dtype=np.float32)
    y = ora.maxpool_prop(x, H, W, ph, pw, pc)
    assert y.shape == (6, 1)
    assert y[0, 0] == 2.0
    assert y[1, 0] == np.float32(-1e20)
    assert y[2, 0] == 0.0 and np.signbit(y[2, 0])
    assert y[3, 0] == 0.0 and not np.signbit(y[3, 0])
    assert y[4, 0] == np.float32(-1e20)
    assert y[5, 0] == np.inf


def test_permutes_vs_reshape(ora):
    rng = np.random.default_rng(7)
    N, C, bs, G = 5, 3, 7, 4
    x = rng.standard_normal((N, C * bs)).astype(np.float32)
    assert np.array_equal(ora.tp_block(x, C, bs),
                          x.reshape(N, C, bs).transpose(1, 0, 2).reshape(C, N * bs))
    x = rng.standard_normal((N, G * bs)).astype(np.float32)
    assert np.array_equal(ora.tp_inside_block(x, G, bs),
                          x.reshape(N, G, bs).transpose(0, 2, 1).reshape(N * bs, G))
    x = rng.standard_normal((C * bs, G)).astype(np.float32)      # rows (pos, c) -> (c, pos)
    assert np.array_equal(ora.mod_permute_row(x, C, bs),
                          x.reshape(bs, C, G).transpose(1, 0, 2).reshape(C * bs, G))
    KH, KW = 3, 2
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    f = ora.flip_mat(k, KH, KW, C, G)
    ref = k.reshape(C, KH * KW, G)[:, ::-1, :].transpose(2, 1, 0).reshape(G * KH * KW, C)
    assert np.array_equal(f, ref)
    H, W = 4, 5
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    p = ora.pad_zero(x, H, W, C, KH, KW)
    ref = np.pad(x.reshape(N, C, W, H), ((0, 0), (0, 0), (KW - 1, KW - 1), (KH - 1, KH - 1)))
    assert np.array_equal(p, ref.reshape(N, -1))
    m = rng.standard_normal((N, G * bs)).astype(np.float32)
    v = rng.standard_normal(G).astype(np.float32)
    assert np.array_equal(ora.add_mat_rep_vec(m, v, bs), m + np.repeat(v, bs)[None, :])


def test_fc_vs_numpy(ora):
    rng = np.random.default_rng(8)
    N, din, dout = 6, 10, 7
    x = rng.standard_normal((N, din))
    Wm = rng.standard_normal((dout, din)) * 0.1
    b = rng.standard_normal(dout)
    dy = rng.standard_normal((N, dout))
    prev = rng.standard_normal((dout, din)) * 0.01
    y = ora.fc_propagate(x, Wm, b, dtype=np.float64)
    assert np.abs(y - (x @ Wm.T + b)).max() < 1e-12
    dx = ora.fc_backprop(dy, Wm, dtype=np.float64)
    assert np.abs(dx - dy @ Wm).max() < 1e-12
    lr, wd, mom = 0.02, 0.0005, 0.9
    W2, b2, p2 = ora.fc_update(x, dy, Wm, b, prev, lr, wd, mom, dtype=np.float64)
    rW, rb, rp = onp.sgd(Wm, b, prev, dy.T @ x, dy.sum(0), N, lr, wd, mom)
    assert np.abs(W2 - rW).max() < 1e-7 and np.abs(b2 - rb).max() < 1e-7
    assert np.abs(p2 - rp).max() < 1e-7
