"""STAGED (runs last, cannot fail the suite): rows a11 / a12 -- ConvolutionComponent::Backprop and the
gradient of ::Update (nnet0/nnet-component-nnet0.cc:461-544, 738-777) -- rebuilt from the reference's
own kernels (oracle/_ref) + cuBLAS SGEMM, against the oracle and the product's fused dgrad / wgrad.

Written after the round-1 GPU budget was spent.  What IS verified: every kernel call pattern
(tests/test_gpu_reference_kernels.py, green on B200), the host order and every matrix shape of the
chains (tests/test_ref_chain_flow.py, CPU).  What is not: this file on a GPU.  Hence
xfail(strict=False): an XPASS at round end is the validation; a failure shows as xfailed and is
round 2's first job.  Drop the marker once it has passed on a B200.
"""
import os

import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="staged: not yet validated on a GPU")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcnsl_ref_kernels.so")


def test_reference_backprop_and_gradient_chains(ora):
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref not built")
    from tests.ref_conv_check import check_backward
    cases = [("C1a", 32, 40, 11, 3, 40, 4, 128), ("time", 16, 1, 14, 64, 1, 3, 128), ("2d", 8, 12, 9, 3, 5, 3, 64)]
    for name, d_ref, d_fp32, d_tf32, w_ref, w_fp32, w_tf32 in check_backward(cases):
        assert d_ref <= 1e-5, ("dgrad: reference chain vs oracle", name, d_ref)
        assert d_fp32 <= 1e-5, ("dgrad: product FP32 vs reference chain", name, d_fp32)
        assert d_tf32 <= 1e-3, ("dgrad: product TF32 vs reference chain", name, d_tf32)
        assert w_ref <= 1e-5, ("wgrad: reference chain vs oracle", name, w_ref)
        assert w_fp32 <= 1e-5, ("wgrad: product FP32 vs reference chain", name, w_fp32)
        assert w_tf32 <= 1e-3, ("wgrad: product TF32 vs reference chain", name, w_tf32)
