"""Parity against the REFERENCE ITSELF for the L0 rows of SURVEY 8(a) (a2 - a9): the reference's own
CUDA kernels (src/cnslmat/cnsl-cu-kernels.cu, compiled unmodified for sm_100a by oracle/Makefile
into oracle/_ref/libcnsl_ref_kernels.so) run on the same inputs as the product's launchers and
as the CPU oracle; all three must agree bit for bit.  This is what pins the oracle for these
rows: its restatement is checked against outputs of the reference, not only against itself.

The reference launches every one of these kernels with 16 x 16 blocks over the (cols, rows) of
the output matrix (conv2D.cc:221-222, 259-260, 305-306, 365-366, 405-406, 442-443, 482-483) --
of out_deriv for Maxpool_backprop (:579-580) -- on the legacy default stream; the product's
legacy launchers take the same arguments and ignore the launch shape.

Skipped when oracle/_ref is absent (it is built where /root/reference exists and travels to the
GPU box with the other built libraries).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import lib, dev, dev_empty, host, assert_bit_exact, mdim, ptr, stream  # noqa: E402
from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.capi import Dim3  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcnsl_ref_kernels.so")
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            pytest.skip("oracle/_ref/libcnsl_ref_kernels.so not built (make -C oracle ref needs /root/reference)")
        lib()                                        # the product first: it owns the CUDA context set-up
        R = ctypes.CDLL(REF_SO)
        for name, args in capi._PROTOS.items():      # same prototypes: the product kept the reference's ABI
            if name.startswith("cudaF_") and hasattr(R, name) and args and args[0] is Dim3:
                getattr(R, name).argtypes = args
                getattr(R, name).restype = None
        _ref = R
    return _ref


def shape_of(t):
    """(Gr, Bl) the reference uses for an output matrix t: 16 x 16 threads over (cols, rows)."""
    r, c = t.shape
    return Dim3((c + 15) // 16, (r + 15) // 16, 1), Dim3(16, 16, 1)


def both(name, out_like, call):
    """Run launcher `name` of the reference and of the product through `call(fn, Gr, Bl)`;
    returns (reference result, product result) as numpy arrays."""
    R, L = ref_lib(), lib()
    gr, bl = shape_of(out_like)
    torch.cuda.synchronize()
    got_ref = call(getattr(R, name), gr, bl)
    torch.cuda.synchronize()
    L.kcnn_set_stream(stream())
    try:
        got_ours = call(getattr(L, name), gr, bl)
        torch.cuda.synchronize()
    finally:
        L.kcnn_set_stream(None)
    return got_ref, got_ours


def _inputs(kind, N, cols, rng):
    x = rng.standard_normal((N, cols)).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0)
    elif kind == "quant":
        x = np.round(x * 4) / 4
    return x.astype(np.float32)


# (N, H, W, C, ph, pw, pc)
POOLS = [
    (64, 1, 8, 128, 1, 2, 2),      # C1a
    (16, 33, 9, 64, 3, 3, 2),      # C1b 3x3x2
    (32, 1, 12, 256, 1, 2, 1),     # nnet.config time pool
    (16, 1, 8, 2000, 1, 2, 10),    # C4 run_conv.sh:56-57 (intermap)
    (3, 4, 6, 6, 2, 3, 3),         # small
]


@pytest.mark.parametrize("shape", POOLS)
@pytest.mark.parametrize("kind", ["randn", "relu", "quant"])
def test_reference_maxpool_prop_and_backprop(ora, shape, kind):
    N, H, W, C, ph, pw, pc = shape
    rng = np.random.default_rng(1000 * sum(shape) + len(kind))
    x = _inputs(kind, N, H * W * C, rng)
    y_ora = ora.maxpool_prop(x, H, W, ph, pw, pc)
    dy = rng.standard_normal(y_ora.shape).astype(np.float32)
    dx_ora = ora.maxpool_backprop(x, y_ora, dy, H, W, ph, pw, pc)
    xd = dev(x, 4, 0)

    def fwd(fn, gr, bl):
        yd = dev_empty(N, y_ora.shape[1], 4, 0)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(yd), mdim(yd), H, W, ph, pw, pc)
        torch.cuda.synchronize()
        return host(yd)

    y_ref, y_ours = both("cudaF_maxpool_prop", dev_empty(N, y_ora.shape[1]), fwd)
    assert_bit_exact(y_ref, y_ora, "reference kernel vs oracle: maxpool_prop")
    assert_bit_exact(y_ours, y_ref, "product vs reference kernel: maxpool_prop")

    yd = dev(y_ora, 4, 0)
    dyd = dev(dy, 4, 0)             # the reference indexes out_deriv with out_value's stride (.cu:293)

    def bwd(fn, gr, bl):
        dxd = dev_empty(N, x.shape[1], 4, 0, fill=0.0)      # MaxpoolComponent::Backprop zero-fills first (:889)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(dyd), mdim(dyd), ptr(dxd), mdim(dxd), H, W, ph, pw, pc)
        torch.cuda.synchronize()
        return host(dxd)

    dx_ref, dx_ours = both("cudaF_maxpool_backprop", dyd, bwd)
    assert_bit_exact(dx_ref, dx_ora, "reference kernel vs oracle: maxpool_backprop")
    assert_bit_exact(dx_ours, dx_ref, "product vs reference kernel: maxpool_backprop")


@pytest.mark.parametrize("N,C,bs", [(64, 3, 440), (7, 5, 1), (33, 128, 8)])
def test_reference_tp_block(ora, N, C, bs):
    x = np.random.default_rng(1).standard_normal((N, C * bs)).astype(np.float32)
    want = ora.tp_block(x, C, bs)
    xd = dev(x, 3, 0)

    def run(fn, gr, bl):
        od = dev_empty(C, N * bs, 5, 0)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(od), mdim(od), bs)
        torch.cuda.synchronize()
        return host(od)

    r, o = both("cudaF_tp_block", dev_empty(C, N * bs), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: tp_block")
    assert_bit_exact(o, r, "product vs reference kernel: tp_block")


@pytest.mark.parametrize("N,G,bs", [(64, 128, 8), (5, 64, 297), (9, 3, 1)])
def test_reference_tp_inside_block(ora, N, G, bs):
    x = np.random.default_rng(2).standard_normal((N, G * bs)).astype(np.float32)
    want = ora.tp_inside_block(x, G, bs)
    xd = dev(x, 3, 0)

    def run(fn, gr, bl):
        od = dev_empty(N * bs, G, 5, 0)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(od), mdim(od), bs)
        torch.cuda.synchronize()
        return host(od)

    r, o = both("cudaF_tp_inside_block", dev_empty(N * bs, G), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: tp_inside_block")
    assert_bit_exact(o, r, "product vs reference kernel: tp_inside_block")


@pytest.mark.parametrize("C,bs,G", [(3, 160, 128), (128, 3, 256), (7, 5, 3)])
def test_reference_mod_permute_row(ora, C, bs, G):
    x = np.random.default_rng(3).standard_normal((C * bs, G)).astype(np.float32)
    want = ora.mod_permute_row(x, C, bs)
    xd = dev(x, 3, 0)

    def run(fn, gr, bl):
        od = dev_empty(C * bs, G, 5, 0)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(od), mdim(od), bs, C)
        torch.cuda.synchronize()
        return host(od)

    r, o = both("cudaF_mod_permute_row", dev_empty(C * bs, G), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: mod_permute_row")
    assert_bit_exact(o, r, "product vs reference kernel: mod_permute_row")


@pytest.mark.parametrize("KH,KW,C,G", [(40, 4, 3, 128), (1, 3, 128, 256), (2, 3, 5, 7)])
def test_reference_flip_mat(ora, KH, KW, C, G):
    k = np.random.default_rng(4).standard_normal((KH * KW * C, G)).astype(np.float32)
    want = ora.flip_mat(k, KH, KW, C, G)
    kd = dev(k, 3, 0)

    def run(fn, gr, bl):
        fd = dev_empty(KH * KW * G, C, 2, 0)
        fn(gr, bl, ptr(kd), mdim(kd), KH, KW, G, ptr(fd), mdim(fd))
        torch.cuda.synchronize()
        return host(fd)

    r, o = both("cudaF_flip_mat", dev_empty(KH * KW * G, C), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: flip_mat")
    assert_bit_exact(o, r, "product vs reference kernel: flip_mat")


@pytest.mark.parametrize("N,H,W,C,KH,KW", [(16, 40, 11, 3, 2, 2), (4, 1, 8, 128, 1, 3), (3, 33, 9, 4, 8, 3)])
def test_reference_pad_zero(ora, N, H, W, C, KH, KW):
    x = np.random.default_rng(5).standard_normal((N, H * W * C)).astype(np.float32)
    want = ora.pad_zero(x, H, W, C, KH, KW)
    xd = dev(x, 3, 0)

    def run(fn, gr, bl):
        pd = dev_empty(N, want.shape[1], 1, 0)
        fn(gr, bl, ptr(xd), mdim(xd), H, W, KH, KW, ptr(pd), mdim(pd))
        torch.cuda.synchronize()
        return host(pd)

    r, o = both("cudaF_pad_zero", dev_empty(N, want.shape[1]), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: pad_zero")
    assert_bit_exact(o, r, "product vs reference kernel: pad_zero")


@pytest.mark.parametrize("N,G,rep", [(64, 128, 8), (7, 3, 5), (16, 64, 297)])
def test_reference_add_mat_rep_vec(ora, N, G, rep):
    rng = np.random.default_rng(6)
    m = rng.standard_normal((N, G * rep)).astype(np.float32)
    v = rng.standard_normal(G).astype(np.float32)
    want = ora.add_mat_rep_vec(m, v, rep)
    vd = torch.from_numpy(v).cuda()

    def run(fn, gr, bl):
        md = dev(m, 6, 0)
        fn(gr, bl, ptr(vd), rep, ptr(md), mdim(md))
        torch.cuda.synchronize()
        return host(md)

    r, o = both("cudaF_add_mat_rep_vec", dev_empty(N, G * rep), run)
    assert_bit_exact(r, want, "reference kernel vs oracle: add_mat_rep_vec")
    assert_bit_exact(o, r, "product vs reference kernel: add_mat_rep_vec")


def test_reference_conv2d_chain(ora):
    """Rows a1 / a10 (forward): the reference's GPU Conv2D rebuilt from ITS kernels -- span_row_to_convmat,
    cuBLAS SGEMM (Kaldi's AddMatMat), convmat_to_out, AddMatRepVec (conv2D.cc:60-185,
    nnet0/nnet-component-nnet0.cc:423-446) -- against the oracle and the product's fused
    cudaF_conv2d_fprop in both math modes."""
    ref_lib()
    from tests.ref_conv_check import check
    cases = [("C1a", 32, 40, 11, 3, 40, 4, 128), ("time", 16, 1, 14, 64, 1, 3, 128), ("2d", 8, 12, 9, 3, 5, 3, 64)]
    for name, e_ref, e_fp32, e_tf32, _ in check(cases):
        assert e_ref <= 1e-5, ("reference chain vs oracle", name, e_ref)
        assert e_fp32 <= 1e-5, ("product FP32 vs reference chain", name, e_fp32)
        assert e_tf32 <= 1e-3, ("product TF32 vs reference chain", name, e_tf32)
