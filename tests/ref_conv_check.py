"""The reference's GPU Conv2D, rebuilt from ITS OWN kernels, next to the product's (NOT YET RUN ON A
GPU: written at the end of round 1 when the GPU budget was spent -- validate before trusting).

    python tests/ref_conv_check.py            # on a GPU box, after `make -C oracle ref` here

CuMatrixBase::Conv2D on the GPU (src/cnslmat/conv2D.cc:60-185) is
    span_row_to_convmat  (reference kernel, im2col)           conv2D.cc:100-108
    AddMatMat            (cuBLAS SGEMM through Kaldi)          conv2D.cc:138-139
    copy_rows_at         (reference kernel; one split here)    conv2D.cc:141-153
    convmat_to_out       (reference kernel, col2im)            conv2D.cc:172-185
and ConvolutionComponent::Propagate adds AddMatRepVec (nnet0/nnet-component-nnet0.cc:423-446).
The three kernels come from oracle/_ref/libcnsl_ref_kernels.so (the unmodified reference file
compiled by oracle/Makefile); the SGEMM is the same cuBLAS routine, called through torch.mm with
TF32 off.  Compared with cudaF_conv2d_fprop (FP32 SIMT: 1e-5, TF32 tensor core: 1e-3) and with
the CPU oracle, and timed: this chain is "the reference recompiled for sm_100a", the GPU
baseline the fused TMA / tcgen05 path replaces.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.capi import Dim3, mdim, ptr, stream  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcnsl_ref_kernels.so")

# (name, N, H, W, C, KH, KW, G): C1a, C1b, conv4 of nnet.config
CASES = [("C1a", 256, 40, 11, 3, 40, 4, 128), ("C1b", 64, 40, 11, 3, 8, 3, 64), ("conv4", 512, 1, 14, 256, 1, 3, 256)]


def grid(rows, cols):
    return Dim3((cols + 15) // 16, (rows + 15) // 16, 1), Dim3(16, 16, 1)


def reference_conv_propagate(R, x, kern, bias, H, W, C, KH, KW, G):
    """x [N x H*W*C], kern [KH*KW*C x G], bias [G] (CUDA tensors) -> out [N x OH*OW*G]."""
    N = x.shape[0]
    OH, OW = H - KH + 1, W - KW + 1
    span = torch.empty(OH * OW * N, KH * KW * C, device="cuda")
    g, b = grid(*span.shape)
    R.cudaF_span_row_to_convmat(g, b, ptr(x), mdim(x), ptr(span), mdim(span), H, W, C, KH, KW, 0)
    conv = torch.mm(span, kern)                                   # AddMatMat(1.0, span, kNoTrans, kernel, kNoTrans, 1.0) on zeros
    out = torch.empty(N, OH * OW * G, device="cuda")
    g, b = grid(*conv.shape)
    R.cudaF_convmat_to_out(g, b, ptr(conv), mdim(conv), ptr(out), mdim(out), OH, OW, N)
    g, b = grid(*out.shape)
    R.cudaF_add_mat_rep_vec(g, b, ptr(bias), OH * OW, ptr(out), mdim(out))
    return out


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def load_reference():
    if not os.path.exists(REF_SO):
        return None
    R = ctypes.CDLL(REF_SO)
    for name, args in capi._PROTOS.items():
        if name.startswith("cudaF_") and hasattr(R, name) and args and args[0] is Dim3:
            getattr(R, name).argtypes = args
            getattr(R, name).restype = None
    return R


def check(cases, timing=False, oracle_dtype=np.float64):
    """[(name, err reference-vs-oracle, err product FP32-vs-reference, err product TF32-vs-reference,
    timings or None)], errors max-norm relative."""
    L = capi.lib()
    R = load_reference()
    if R is None:
        raise RuntimeError("build oracle/_ref first: make -C oracle ref (needs /root/reference)")
    from oracle import oracle as ora
    ora.build()
    tf32_was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False          # Kaldi's AddMatMat is cublasSgemm: plain FP32
    rows = []
    try:
        for name, N, H, W, C, KH, KW, G in cases:
            rng = np.random.default_rng(1234)
            x = rng.standard_normal((N, H * W * C)).astype(np.float32)
            k = (rng.standard_normal((KH * KW * C, G)) * 0.05).astype(np.float32)
            bias = rng.standard_normal(G).astype(np.float32)
            xd, kd, bd = torch.from_numpy(x).cuda(), torch.from_numpy(k).cuda(), torch.from_numpy(bias).cuda()
            OH, OW = H - KH + 1, W - KW + 1
            ref = reference_conv_propagate(R, xd, kd, bd, H, W, C, KH, KW, G)
            torch.cuda.synchronize()
            want = ora.conv_propagate(x, k, bias, H, W, C, 0, 0, KH, KW, G, dtype=oracle_dtype)
            scale = float(np.abs(want).max())
            e_ref = float(np.abs(ref.cpu().numpy().astype(np.float64) - want).max()) / scale
            errs, times = [], {}
            for math in (0, 1):
                out = torch.empty(N, OH * OW * G, device="cuda")
                call = lambda: L.cudaF_conv2d_fprop(stream(), math, ptr(xd), mdim(xd), ptr(kd), mdim(kd), ptr(bd),  # noqa: E731
                                                    ptr(out), mdim(out), H, W, C, 0, 0, KH, KW, G, 1)
                call()
                torch.cuda.synchronize()
                errs.append(float((out - ref).abs().max()) / scale)
                if timing:
                    times["product_math%d_us" % math] = timed(call) * 1e3
            if timing:
                times["reference_chain_us"] = timed(lambda: reference_conv_propagate(R, xd, kd, bd, H, W, C, KH, KW, G)) * 1e3
            rows.append((name, e_ref, errs[0], errs[1], times if timing else None))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32_was
    return rows


def main():
    ok = True
    for name, e_ref, e0, e1, times in check(CASES, timing=True):
        ok = ok and e_ref <= 1e-5 and e0 <= 1e-5 and e1 <= 1e-3
        print("%-6s reference(GPU kernels + cuBLAS) vs oracle(f64) %.2e | product FP32 vs reference %.2e | "
              "product TF32 vs reference %.2e | %s" % (name, e_ref, e0, e1, json.dumps(times)), flush=True)
    print("ref_conv_check", "ok" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
