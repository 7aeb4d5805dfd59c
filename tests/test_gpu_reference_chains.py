"""Rows a10 / a11 / a12 pinned against the REFERENCE'S OWN kernels: ConvolutionComponent::Propagate with
padding, ::Backprop (both dgrad branches) and the gradient of ::Update (nnet0/nnet-component-nnet0.cc:423-446,
461-544, 738-777) rebuilt from the unmodified cnsl-cu-kernels.cu (oracle/_ref) + cuBLAS SGEMM in the host order
of the reference, against the oracle and the product's fused forward / dgrad / wgrad; plus the overlap /
overlap2D max-pool forward kernels.  Validated on B200 in round 1 (then marked xfail-non-strict because it
had been written without a GPU); now ordinary tests, at the C1a / C1b sizes of SURVEY 8d too.
"""
import os

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcnsl_ref_kernels.so")


def test_reference_backprop_and_gradient_chains(ora):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    from tests.ref_conv_check import check_backward
    cases = [("C1a", 256, 40, 11, 3, 40, 4, 128),          # no-flip dgrad branch
             ("C1b", 256, 40, 11, 3, 8, 3, 64),            # flip branch: pads out_deriv to 40 MB
             ("time", 16, 1, 14, 64, 1, 3, 128), ("2d", 8, 12, 9, 3, 5, 3, 64)]
    for name, d_ref, d_fp32, d_tf32, w_ref, w_fp32, w_tf32 in check_backward(cases):
        assert d_ref <= 1e-5, ("dgrad: reference chain vs oracle", name, d_ref)
        assert d_fp32 <= 1e-5, ("dgrad: product FP32 vs reference chain", name, d_fp32)
        assert d_tf32 <= 1e-3, ("dgrad: product TF32 vs reference chain", name, d_tf32)
        assert w_ref <= 1e-5, ("wgrad: reference chain vs oracle", name, w_ref)
        assert w_fp32 <= 1e-5, ("wgrad: product FP32 vs reference chain", name, w_fp32)
        assert w_tf32 <= 1e-3, ("wgrad: product TF32 vs reference chain", name, w_tf32)


# (N, H, W, C, pc, mode): the overlap shapes of tests/test_gpu_l0_pool_permute.py; mode 1 = overlap
# (stride-1 pooling across channels), 2 = overlap2D (channels as a 2-D map) -- SURVEY 8f-4
OVERLAPS = [(16, 1, 8, 12, 3, 1), (8, 2, 3, 9, 4, 1), (16, 1, 8, 16, 2, 2), (8, 2, 2, 25, 3, 2)]


@pytest.mark.parametrize("shape", OVERLAPS)
def test_reference_overlap_maxpool_forward(ora, shape):
    """Maxpool_prop with overlap / overlap2D (conv2D.cc:485-489, cnsl-cu-kernels.cu:310-350, 405-450): the
    reference's forward kernels next to the oracle and the product, bit for bit.  (The reference's
    overlap BACKWARD kernels race -- .cu:396-397, 497-498 -- and are not a usable checker.)"""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    import numpy as np
    import torch
    from tests.gpu_util import dev, dev_empty, host, assert_bit_exact, mdim, ptr
    from tests.test_gpu_reference_kernels import both
    N, H, W, C, pc, mode = shape
    x = np.maximum(np.random.default_rng(13).standard_normal((N, H * W * C)), 0).astype(np.float32)
    y_ora = ora.maxpool_prop(x, H, W, 1, 1, pc, mode=mode)
    xd = dev(x, 3, 0)
    name = "cudaF_maxpoolchannel_overlap_prop" if mode == 1 else "cudaF_maxpoolchannel_overlap2D_prop"

    def fwd(fn, gr, bl):
        yd = dev_empty(N, y_ora.shape[1], 3, 0)
        fn(gr, bl, ptr(xd), mdim(xd), ptr(yd), mdim(yd), H, W, 1, 1, pc)
        torch.cuda.synchronize()
        return host(yd)

    y_ref, y_ours = both(name, dev_empty(N, y_ora.shape[1]), fwd)
    assert_bit_exact(y_ref, y_ora, "reference kernel vs oracle: " + name)
    assert_bit_exact(y_ours, y_ref, "product vs reference kernel: " + name)


def test_reference_padded_propagate_chain(ora):
    """ConvolutionComponent::Propagate with in-pad-height / in-pad-width (nnet0/nnet-component-nnet0.cc:
    430-435): PaddingZero(pad + 1) then Conv2D on the padded tensor, from the reference's kernels,
    against the oracle and the product's fused forward (padding as TMA / index arithmetic)."""
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    import numpy as np
    import torch
    from kaldi_cnn_b200 import capi
    from kaldi_cnn_b200.capi import mdim, ptr, stream
    from tests.ref_conv_check import RefOps, load_reference, reference_conv_propagate
    L = capi.lib()
    R = load_reference()
    ops = RefOps(R)
    was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for (N, H, W, C, ph, pw, KH, KW, G) in [(9, 6, 7, 5, 1, 2, 3, 4, 10), (16, 1, 14, 64, 0, 1, 1, 3, 128)]:
            rng = np.random.default_rng(21)
            x = rng.standard_normal((N, H * W * C)).astype(np.float32)
            k = (rng.standard_normal((KH * KW * C, G)) * 0.05).astype(np.float32)
            b = rng.standard_normal(G).astype(np.float32)
            xd, kd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(b).cuda()
            Hp, Wp = H + 2 * ph, W + 2 * pw
            OH, OW = Hp - KH + 1, Wp - KW + 1
            padded = ops.pad_zero(xd, H, W, C, ph + 1, pw + 1)
            ref = reference_conv_propagate(R, padded, kd, bd, Hp, Wp, C, KH, KW, G)
            torch.cuda.synchronize()
            want = ora.conv_propagate(x, k, b, H, W, C, ph, pw, KH, KW, G, dtype=np.float64)
            scale = float(np.abs(want).max())
            assert float(np.abs(ref.cpu().numpy() - want).max()) / scale <= 1e-5
            for math, tol in ((0, 1e-5), (1, 1e-3)):
                out = torch.empty(N, OH * OW * G, device="cuda")
                L.cudaF_conv2d_fprop(stream(), math, ptr(xd), mdim(xd), ptr(kd), mdim(kd), ptr(bd), ptr(out), mdim(out),
                                     H, W, C, ph, pw, KH, KW, G, 1)
                torch.cuda.synchronize()
                assert float((out - ref).abs().max()) / scale <= tol, (math, N, H, W, C, ph, pw)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = was
