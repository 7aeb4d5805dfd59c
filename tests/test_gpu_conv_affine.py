"""GPU parity: the fused implicit-GEMM entry points (convolution fprop / dgrad / wgrad,
affine fprop / dgrad / wgrad, SGD) through the C-ABI, against the CPU oracle.

Tolerances (max |got - ref| / max |ref|):
  KCNN_MATH_FP32_SIMT : 1e-5      KCNN_MATH_TF32_TC : 1e-3
The oracle is evaluated in FP64 (oraD_*) so the bound is on OUR error, and the FP32
oracle itself is required to sit inside the same bound.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import lib, dev, dev_empty, host, rel_err, mdim, ptr, stream  # noqa: E402

TOL = {0: 1e-5, 1: 1e-3}

# (name, N, H, W, C, pad_h, pad_w, KH, KW, G)
CONVS = [
    ("C1a", 256, 40, 11, 3, 0, 0, 40, 4, 128),
    ("C1b", 48, 40, 11, 3, 0, 0, 8, 3, 64),
    ("conv1", 64, 40, 21, 1, 0, 0, 40, 4, 128),
    ("conv2", 64, 1, 18, 128, 0, 0, 1, 3, 128),
    ("conv4", 32, 1, 14, 256, 0, 0, 1, 3, 256),
    ("conv6", 64, 1, 4, 512, 0, 0, 1, 3, 512),
    ("t_pad", 33, 1, 10, 64, 0, 1, 1, 3, 96),       # time-axis + zero padding, sample / map tile tails
    ("t_wide", 20, 1, 18, 32, 0, 2, 1, 5, 160),
    ("t_c40", 17, 1, 12, 40, 0, 0, 1, 3, 64),        # channel count not a multiple of 32 (padded M atoms)
    ("t_c200", 8, 1, 8, 200, 0, 0, 1, 5, 200),
    ("pad2d", 9, 6, 7, 5, 1, 1, 3, 3, 10),
    ("padw", 7, 5, 9, 4, 0, 2, 5, 4, 6),
    ("ragged", 5, 3, 5, 7, 0, 0, 2, 2, 3),
    ("one", 1, 1, 1, 1, 0, 0, 1, 1, 1),
]


def _conv_data(cfg, seed=0):
    _, N, H, W, C, ph, pw, KH, KW, G = cfg
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    k = (rng.standard_normal((KH * KW * C, G)) * 0.05).astype(np.float32)
    b = (rng.standard_normal(G) * 0.5).astype(np.float32)
    dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
    return x, k, b, dy, OH, OW


@pytest.mark.parametrize("cfg", CONVS, ids=[c[0] for c in CONVS])
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_conv_fprop(ora, cfg, math, layout):
    _, N, H, W, C, ph, pw, KH, KW, G = cfg
    x, k, b, _, OH, OW = _conv_data(cfg)
    ref = ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
    pad, off = (0, 0) if layout == "packed" else (5, 3)
    xd, kd, od = dev(x, pad, off), dev(k, pad, off), dev_empty(N, OH * OW * G, pad, off)
    bd = torch.from_numpy(b).cuda()
    lib().cudaF_conv2d_fprop(stream(), math, ptr(xd), mdim(xd), ptr(kd), mdim(kd), ptr(bd), ptr(od), mdim(od),
                             H, W, C, ph, pw, KH, KW, G, 1)
    assert rel_err(host(od), ref) <= TOL[math]
    assert rel_err(ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G), ref) <= 1e-5


@pytest.mark.parametrize("cfg", CONVS[:7], ids=[c[0] for c in CONVS[:7]])
@pytest.mark.parametrize("math", [0, 1])
def test_conv2d_member_semantics_concat_false(ora, cfg, math):
    """CuMatrixBase::Conv2D(..., concat=false): raw [(pos*N + n) x G] matrix (conv2D.cc:199)."""
    _, N, H, W, C, ph, pw, KH, KW, G = cfg
    if ph or pw:
        pytest.skip("Conv2D member has no padding")
    x, k, _, _, OH, OW = _conv_data(cfg, 1)
    ref = ora.conv2d(x, k, H, W, C, KH, KW, G, concat=False, dtype=np.float64)
    xd, kd, od = dev(x, 3, 1), dev(k, 1, 1), dev_empty(OH * OW * N, G, 2, 1)
    lib().cudaF_conv2d_fprop(stream(), math, ptr(xd), mdim(xd), ptr(kd), mdim(kd), None, ptr(od), mdim(od),
                             H, W, C, 0, 0, KH, KW, G, 0)
    assert rel_err(host(od), ref) <= TOL[math]


@pytest.mark.parametrize("cfg", CONVS, ids=[c[0] for c in CONVS])
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_conv_dgrad_matches_both_reference_branches(ora, cfg, math, layout):
    _, N, H, W, C, ph, pw, KH, KW, G = cfg
    x, k, b, dy, OH, OW = _conv_data(cfg, 2)
    ref0 = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G, branch=0, dtype=np.float64)
    ref1 = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G, branch=1, dtype=np.float64)
    assert rel_err(ref0, ref1) < 1e-12
    pad, off = (0, 0) if layout == "packed" else (5, 3)
    dyd, kd, dxd = dev(dy, pad, off), dev(k, pad, off), dev_empty(N, H * W * C, pad, off)
    lib().cudaF_conv2d_dgrad(stream(), math, ptr(dyd), mdim(dyd), ptr(kd), mdim(kd), ptr(dxd), mdim(dxd),
                             H, W, C, ph, pw, KH, KW, G)
    assert rel_err(host(dxd), ref0) <= TOL[math]


@pytest.mark.parametrize("cfg", CONVS, ids=[c[0] for c in CONVS])
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_conv_wgrad_and_bias_grad(ora, cfg, math, layout):
    _, N, H, W, C, ph, pw, KH, KW, G = cfg
    x, k, b, dy, OH, OW = _conv_data(cfg, 3)
    _, _, _, gref, bref = ora.conv_update(x, dy, k, b, np.zeros_like(k), H, W, C, ph, pw, KH, KW, G,
                                          0.02, 0.0005, 0.9, apply=False, dtype=np.float64)
    L = lib()
    pad, off = (0, 0) if layout == "packed" else (5, 3)
    xd, dyd, gd = dev(x, pad, off), dev(dy, pad, off), dev_empty(KH * KW * C, G, pad, off)
    bg = torch.full((G,), float("nan"), device="cuda")
    nbytes = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, ph, pw, KH, KW, G)
    ws = torch.empty(max(nbytes, 4) // 4, dtype=torch.float32, device="cuda")
    L.cudaF_conv2d_wgrad(stream(), math, ptr(xd), mdim(xd), ptr(dyd), mdim(dyd), ptr(gd), mdim(gd), ptr(bg),
                         ptr(ws), H, W, C, ph, pw, KH, KW, G)
    assert rel_err(host(gd), gref) <= TOL[math]
    assert rel_err(bg.cpu().numpy(), bref) <= 1e-5


AFFINES = [(256, 256, 1024), (64, 1056, 1024), (128, 1024, 4096), (33, 70, 130), (1, 1, 1), (512, 4096, 3454)]


@pytest.mark.parametrize("N,din,dout", AFFINES)
@pytest.mark.parametrize("math", [0, 1])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_affine_fprop_dgrad_wgrad(ora, N, din, dout, math, layout):
    rng = np.random.default_rng(4)
    x = rng.standard_normal((N, din)).astype(np.float32)
    Wm = (rng.standard_normal((dout, din)) * 0.05).astype(np.float32)
    b = rng.standard_normal(dout).astype(np.float32)
    dy = rng.standard_normal((N, dout)).astype(np.float32)
    x64, W64, dy64 = x.astype(np.float64), Wm.astype(np.float64), dy.astype(np.float64)
    L = lib()
    pad, off = (0, 0) if layout == "packed" else (5, 3)
    xd, wd, dyd = dev(x, pad, off), dev(Wm, pad, off), dev(dy, pad, off)
    bd = torch.from_numpy(b).cuda()
    yd = dev_empty(N, dout, pad, off)
    L.cudaF_affine_fprop(stream(), math, ptr(xd), mdim(xd), ptr(wd), mdim(wd), ptr(bd), ptr(yd), mdim(yd))
    assert rel_err(host(yd), x64 @ W64.T + b) <= TOL[math]
    dxd = dev_empty(N, din, pad, off)
    L.cudaF_affine_dgrad(stream(), math, ptr(dyd), mdim(dyd), ptr(wd), mdim(wd), ptr(dxd), mdim(dxd))
    assert rel_err(host(dxd), dy64 @ W64) <= TOL[math]
    gd = dev_empty(dout, din, pad, off)
    bg = torch.full((dout,), float("nan"), device="cuda")
    L.cudaF_affine_wgrad(stream(), math, ptr(xd), mdim(xd), ptr(dyd), mdim(dyd), ptr(gd), mdim(gd), ptr(bg))
    assert rel_err(host(gd), dy64.T @ x64) <= TOL[math]
    assert rel_err(bg.cpu().numpy(), dy64.sum(0)) <= 1e-5


def test_sgd_momentum_update_matches_oracle_rounding(ora):
    """prev = m prev - lr wd W + lr dW ; W += prev  (nnet0/nnet-component-nnet0.cc:767-773, 1138-1142), bit for bit.
    What pins this: the reference performs it with stock Kaldi calls (CuMatrix::Scale / AddMat, :769-773) whose
    sources are NOT part of /root/reference, so there is no reference binary to run next to it -- the expected
    values are this NumPy restatement of those calls' arithmetic (one rounding per AddMat, fused multiply-add), the
    same order the oracle's C port uses.  It is the one row where "green" rests on a restatement alone."""
    rng = np.random.default_rng(5)
    R, Cc = 97, 130
    w = rng.standard_normal((R, Cc)).astype(np.float32)
    p = (rng.standard_normal((R, Cc)) * 0.01).astype(np.float32)
    g = rng.standard_normal((R, Cc)).astype(np.float32)
    mom, a_decay, a_grad = np.float32(0.9), np.float32(-3.9e-8), np.float32(7.8125e-5)
    p_ref = p * mom
    p_ref = (p_ref.astype(np.float64) + np.float64(a_decay) * w).astype(np.float32)   # fma
    p_ref = (p_ref.astype(np.float64) + np.float64(a_grad) * g).astype(np.float32)    # fma
    w_ref = w + p_ref
    for pad, off in ((0, 0), (2, 0), (5, 3)):
        wd, pd_, gd = dev(w, pad, off), dev(p, pad, off), dev(g, pad, off)
        lib().cudaF_sgd_momentum_update(stream(), ptr(wd), mdim(wd), ptr(pd_), mdim(pd_), ptr(gd), mdim(gd),
                                        float(mom), float(a_decay), float(a_grad))
        assert np.array_equal(host(pd_), p_ref) and np.array_equal(host(wd), w_ref)
