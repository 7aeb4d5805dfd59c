"""GPU parity at BASELINE.json's FULL sizes (minibatch 512, the layer shapes of
egs/exp/nnet/nnet.config), where the CPU oracle would take minutes: size-independent
properties instead of element-wise comparison.

  * adjoint identities tie the three passes of a layer together:
        <fprop(x) - bias, dy>  ==  <x, dgrad(dy)>  ==  <K, wgrad(x, dy)>
    (exact in exact arithmetic; each side is a TF32 contraction here, so the difference is
    bounded by 1e-3 * ||.|| ||.||, the TF32 tolerance on the Cauchy-Schwarz scale)
  * linearity of the forward pass in its input
  * max-pooling: bit-exact against the reshape / amax formulation (max is exact in any
    order as long as ties are bitwise equal; random normal data has no +-0 ties), and the
    backward routing is bit-exact against the "every element equal to the max" formulation
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import lib, mdim, ptr, stream  # noqa: E402

N = 512
CONVS = [  # (H, W, C, KH, KW, G) of nnet.config (+ intermap variant's conv2)
    (40, 21, 1, 40, 4, 128), (1, 18, 64, 1, 3, 128), (1, 16, 128, 1, 3, 256), (1, 14, 256, 1, 3, 256),
    (1, 6, 256, 1, 3, 512), (1, 4, 512, 1, 3, 512),
    (1, 8, 2000, 1, 5, 2000),      # C3(iii), SURVEY 8d (fbank_conv.sh:251), at its own batch of 256 (below)
]


def _dot(a, b):
    return float((a.double() * b.double()).sum())


@pytest.mark.parametrize("shape", CONVS, ids=["conv%d" % (i + 1) for i in range(6)] + ["c3iii"])
def test_conv_adjoint_identities_and_linearity_full_size(shape):
    H, W, C, KH, KW, G = shape
    N = 256 if C == 2000 else 512
    OH, OW = H - KH + 1, W - KW + 1
    L = lib()
    g = torch.Generator(device="cuda"); g.manual_seed(sum(shape))
    x = torch.randn(N, H * W * C, device="cuda", generator=g)
    x2 = torch.randn(N, H * W * C, device="cuda", generator=g)
    k = torch.randn(KH * KW * C, G, device="cuda", generator=g) * 0.05
    dy = torch.randn(N, OH * OW * G, device="cuda", generator=g)
    zero_b = torch.zeros(G, device="cuda")

    def fprop(inp):
        y = torch.empty(N, OH * OW * G, device="cuda")
        L.cudaF_conv2d_fprop(stream(), 1, ptr(inp), mdim(inp), ptr(k), mdim(k), ptr(zero_b), ptr(y), mdim(y),
                             H, W, C, 0, 0, KH, KW, G, 1)
        return y
    y = fprop(x)
    dx = torch.empty_like(x)
    L.cudaF_conv2d_dgrad(stream(), 1, ptr(dy), mdim(dy), ptr(k), mdim(k), ptr(dx), mdim(dx), H, W, C, 0, 0, KH, KW, G)
    kg = torch.empty_like(k); bg = torch.empty(G, device="cuda")
    nb = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, 0, 0, KH, KW, G)
    ws = torch.empty(max(nb, 4) // 4, device="cuda")
    L.cudaF_conv2d_wgrad(stream(), 1, ptr(x), mdim(x), ptr(dy), mdim(dy), ptr(kg), mdim(kg), ptr(bg), ptr(ws),
                         H, W, C, 0, 0, KH, KW, G)
    torch.cuda.synchronize()
    lhs, mid, rhs = _dot(y, dy), _dot(x, dx), _dot(k, kg)
    scale = float(y.double().norm() * dy.double().norm())
    assert abs(lhs - mid) <= 1e-3 * scale, (lhs, mid, scale)
    assert abs(lhs - rhs) <= 1e-3 * scale, (lhs, rhs, scale)
    # bias gradient = column sums of dy per map (FP32 reduction)
    ref_bg = dy.view(N, G, OH * OW).double().sum(dim=(0, 2))
    assert float((bg.double() - ref_bg).abs().max()) <= 1e-5 * float(ref_bg.abs().max()) * 8
    # linearity in the input
    y2, y12 = fprop(x2), fprop(2.0 * x + x2)
    err = float((y12 - (2.0 * y + y2)).abs().max()) / float(y12.abs().max())
    assert err <= 2e-3, err


@pytest.mark.parametrize("din,dout", [(1024, 4096), (4096, 4096), (4096, 3454)])
def test_affine_adjoint_identities_full_size(din, dout):
    L = lib()
    g = torch.Generator(device="cuda"); g.manual_seed(din + dout)

    def pitched(r, c):
        ld = (c + 3) // 4 * 4
        return torch.randn(r, ld, device="cuda", generator=g)[:, :c]
    x, w, dy = pitched(N, din), pitched(dout, din) * 0.02, pitched(N, dout)
    b0 = torch.zeros(dout, device="cuda")
    y = torch.empty(N, (dout + 3) // 4 * 4, device="cuda")[:, :dout]
    L.cudaF_affine_fprop(stream(), 1, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b0), ptr(y), mdim(y))
    dx = torch.empty(N, din, device="cuda")
    L.cudaF_affine_dgrad(stream(), 1, ptr(dy), mdim(dy), ptr(w), mdim(w), ptr(dx), mdim(dx))
    gw = torch.empty(dout, din, device="cuda"); gb = torch.empty(dout, device="cuda")
    L.cudaF_affine_wgrad(stream(), 1, ptr(x), mdim(x), ptr(dy), mdim(dy), ptr(gw), mdim(gw), ptr(gb))
    torch.cuda.synchronize()
    lhs, mid, rhs = _dot(y, dy), _dot(x, dx), _dot(w, gw)
    scale = float(y.double().norm() * dy.double().norm())
    assert abs(lhs - mid) <= 1e-3 * scale and abs(lhs - rhs) <= 1e-3 * scale, (lhs, mid, rhs, scale)
    ref_gb = dy.double().sum(0)
    assert float((gb.double() - ref_gb).abs().max()) <= 1e-5 * float(ref_gb.abs().max()) * 8
    # fused weight-gradient + SGD step == separate gradient followed by the update formula
    w2, pv = w.clone().contiguous(), torch.randn(dout, din, device="cuda", generator=g) * 0.01
    w_ref = w2.double(); p_ref = pv.double()
    mom, a_decay, a_grad = 0.9, -2e-8, 4e-5
    done = L.cudaF_affine_wgrad_sgd(stream(), 1, ptr(x), mdim(x), ptr(dy), mdim(dy), ptr(w2), mdim(w2), ptr(pv),
                                    mdim(pv), ptr(b0), mom, a_decay, a_grad)
    torch.cuda.synchronize()
    assert done == 1
    p_new = mom * p_ref + a_decay * w_ref + a_grad * gw.double()
    assert float((pv.double() - p_new).abs().max()) <= 1e-3 * float(p_new.abs().max())
    assert float((w2.double() - (w_ref + p_new)).abs().max()) <= 1e-3 * float((p_new).abs().max()) + 1e-7


@pytest.mark.parametrize("geom", [(1, 12, 256, 1, 2, 1, 512), (1, 18, 128, 1, 1, 2, 512), (1, 16, 2000, 1, 2, 1, 8192),
                                  (1, 8, 2000, 1, 2, 10, 8192), (33, 9, 64, 3, 3, 2, 4096), (1, 1, 4000, 1, 1, 5, 8192)])
def test_maxpool_full_size_bit_exact_and_mass_conserving(geom):
    H, W, C, ph, pw, pc, n = geom
    L = lib()
    g = torch.Generator(device="cuda"); g.manual_seed(sum(geom))
    x = torch.randn(n, H * W * C, device="cuda", generator=g)
    OH, OW, OC = H // ph, W // pw, C // pc
    y = torch.empty(n, OH * OW * OC, device="cuda")
    L.cudaF_maxpool_prop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), H, W, ph, pw, pc, 0)
    ref = x.view(n, OC, pc, OW, pw, OH, ph).amax(dim=(2, 4, 6)).reshape(n, -1)
    assert torch.equal(y.view(torch.int32), ref.view(torch.int32))
    dy = torch.randn(n, OH * OW * OC, device="cuda", generator=g)
    dx = torch.full((n, H * W * C), float("nan"), device="cuda")
    L.cudaF_maxpool_backprop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), ptr(dy), mdim(dy), ptr(dx), mdim(dx),
                               H, W, ph, pw, pc, 0, 1)
    torch.cuda.synchronize()
    # reference-exact routing (cnsl-cu-kernels.cu:302-303): err goes to EVERY window element equal to
    # the pooled value (a handful of exact ties do occur among 10^8 random floats), zero elsewhere
    win_x = x.view(n, OC, pc, OW, pw, OH, ph)
    expect = torch.where(win_x == ref.view(n, OC, 1, OW, 1, OH, 1), dy.view(n, OC, 1, OW, 1, OH, 1),
                         torch.zeros((), device="cuda")).reshape(n, -1)
    assert torch.equal(dx.view(torch.int32), expect.view(torch.int32))
    # and the derivative mass of each window is conserved up to its tie count
    ties = (win_x == ref.view(n, OC, 1, OW, 1, OH, 1)).sum(dim=(2, 4, 6)).reshape(n, -1)
    assert int(ties.min()) >= 1
