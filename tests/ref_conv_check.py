"""The reference's GPU Conv2D, rebuilt from ITS OWN kernels, next to the product's.  The forward chain (check(), used by
tests/test_gpu_reference_kernels.py::test_reference_conv2d_chain) is green on B200; the timing legs of main() and
everything under --backward were written after the round-1 GPU budget was spent: validate before trusting.

    python tests/ref_conv_check.py            # on a GPU box, after `make -C oracle ref` here
    python tests/ref_conv_check.py --backward # Backprop / Update chains (round 2: never run yet)

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


def out_shape(op, rows, *a):
    """Shape of the matrix each L0 member produces from a [rows x .] input (conv2D.cc asserts /
    Resize calls); shared by RefOps (allocation) and the CPU flow test (checked against the oracle)."""
    if op == "tp_block":                    # (C, bs):            [rows x C*bs] -> [C x rows*bs]
        C, bs = a
        return C, rows * bs
    if op == "tp_inside_block":             # (G, bs):            [rows x G*bs] -> [rows*bs x G]
        G, bs = a
        return rows * bs, G
    if op == "flip_mat":                    # (KH, KW, C, G):     [KH*KW*C x G] -> [KH*KW*G x C]
        KH, KW, C, G = a
        return KH * KW * G, C
    if op == "pad_zero":                    # (H, W, C, KH, KW):  pads KH-1 / KW-1 per side
        H, W, C, KH, KW = a
        return rows, (H + 2 * (KH - 1)) * (W + 2 * (KW - 1)) * C
    if op == "conv2d":                      # (H, W, C, KH, KW, G, concat)
        H, W, C, KH, KW, G, concat = a
        OH, OW = H - KH + 1, W - KW + 1
        return (rows, OH * OW * G) if concat else (OH * OW * rows, G)
    raise ValueError(op)


class RefOps:
    """The reference's L0 kernels as matrix -> matrix functions on CUDA tensors (fresh contiguous
    outputs), with the launch shapes of conv2D.cc (16 x 16 threads over the output)."""

    def __init__(self, R):
        self.R = R

    @staticmethod
    def _new(op, x, *a):
        return torch.empty(*out_shape(op, x.shape[0], *a), device="cuda")

    def tp_block(self, x, C, bs):                       # conv2D.cc:348-386   out[c, n*bs+p] = x[n, c*bs+p]
        out = self._new("tp_block", x, C, bs)
        g, b = grid(*out.shape)
        self.R.cudaF_tp_block(g, b, ptr(x), mdim(x), ptr(out), mdim(out), bs)
        return out

    def tp_inside_block(self, x, G, bs):                # :388-426            out[n*bs+p, g] = x[n, g*bs+p]
        out = self._new("tp_inside_block", x, G, bs)
        g, b = grid(*out.shape)
        self.R.cudaF_tp_inside_block(g, b, ptr(x), mdim(x), ptr(out), mdim(out), bs)
        return out

    def flip_mat(self, k, KH, KW, C, G):                # :244-287            [KH*KW*C x G] -> [KH*KW*G x C]
        out = self._new("flip_mat", k, KH, KW, C, G)
        g, b = grid(*out.shape)
        self.R.cudaF_flip_mat(g, b, ptr(k), mdim(k), KH, KW, G, ptr(out), mdim(out))
        return out

    def pad_zero(self, x, H, W, C, KH, KW):             # :289-344            pads KH-1 / KW-1 per side
        out = self._new("pad_zero", x, H, W, C, KH, KW)
        g, b = grid(*out.shape)
        self.R.cudaF_pad_zero(g, b, ptr(x), mdim(x), H, W, KH, KW, ptr(out), mdim(out))
        return out

    def mod_permute_row(self, x, C, bs):                # :429-463            row i -> row (i % C)*bs + i / C
        out = torch.empty_like(x)
        g, b = grid(*out.shape)
        self.R.cudaF_mod_permute_row(g, b, ptr(x), mdim(x), ptr(out), mdim(out), bs, C)
        return out

    def conv2d(self, x, kern, H, W, C, KH, KW, G, concat):   # :44-201, one split
        n = x.shape[0]
        OH, OW = H - KH + 1, W - KW + 1
        span = torch.empty(OH * OW * n, KH * KW * C, device="cuda")
        g, b = grid(*span.shape)
        self.R.cudaF_span_row_to_convmat(g, b, ptr(x), mdim(x), ptr(span), mdim(span), H, W, C, KH, KW, 0)
        conv = torch.mm(span, kern)
        if not concat:
            return conv
        out = self._new("conv2d", x, H, W, C, KH, KW, G, True)
        g, b = grid(*conv.shape)
        self.R.cudaF_convmat_to_out(g, b, ptr(conv), mdim(conv), ptr(out), mdim(out), OH, OW, n)
        return out


def reference_conv_backprop(ops, out_deriv, lin, H, W, C, ph, pw, KH, KW, G, branch=-1):
    """ConvolutionComponent::Backprop, input-derivative part (nnet0/nnet-component-nnet0.cc:461-540),
    chained from the reference kernels in the order oracle/kcnn_oracle_impl.h restates.
    NOT YET RUN ON A GPU."""
    N = out_deriv.shape[0]
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    ks, osz = KH * KW, OH * OW
    pkh, pkw = KH + 2 * (OH - ph - 1), KW + 2 * (OW - pw - 1)
    poh, pow_ = OH + 2 * (KH - ph - 1), OW + 2 * (KW - pw - 1)
    flip = (not (pkh * pkw < poh * pow_)) if branch < 0 else bool(branch)          # :489-497
    if not flip:                                                                    # :499-528
        od_tp = ops.tp_inside_block(out_deriv, G, osz)                              # [os*N x G]
        flip_od = ops.flip_mat(od_tp, OH, OW, N, G)                                 # [os*G x N]
        lin_tp = lin.t().contiguous()                                               # AddMat(kTrans) :516  [G x ks*C]
        lin_tp2 = ops.tp_block(lin_tp, C, ks)                                       # [C x ks*G]
        pad_k = ops.pad_zero(lin_tp2, KH, KW, G, OH - ph, OW - pw)                  # [C x pkh*pkw*G]
        tmp = ops.conv2d(pad_k, flip_od, pkh, pkw, G, OH, OW, N, True)              # [C x H*W*N]
        return ops.tp_block(tmp, N, H * W)                                          # :525  [N x H*W*C]
    pad_od = ops.pad_zero(out_deriv, OH, OW, G, KH - ph, KW - pw)                   # :530-538
    flip_k = ops.flip_mat(lin, KH, KW, C, G)                                        # [ks*G x C]
    return ops.conv2d(pad_od, flip_k, poh, pow_, G, KH, KW, C, True)


def reference_conv_gradient(ops, in_value, out_deriv, H, W, C, ph, pw, KH, KW, G):
    """The un-normalised weight and bias gradients of ConvolutionComponent::Update (:745-765, 775).
    NOT YET RUN ON A GPU."""
    N = in_value.shape[0]
    Hp, Wp = H + 2 * ph, W + 2 * pw
    OH, OW = Hp - KH + 1, Wp - KW + 1
    x = ops.pad_zero(in_value, H, W, C, ph + 1, pw + 1) if (ph or pw) else in_value   # :751-757
    iv_tmp = ops.tp_block(x, C, Hp * Wp)                                             # [C x N*Hp*Wp]
    od_tmp = ops.tp_inside_block(out_deriv, G, OH * OW)                              # [os*N x G]
    lp_tmp = ops.conv2d(iv_tmp, od_tmp, Hp, Wp, N, OH, OW, G, False)                 # :763  [ks*C x G], rows (r*C + c)
    return ops.mod_permute_row(lp_tmp, C, KH * KW), od_tmp.sum(0)                    # :765, :775


def check_backward(cases):
    """[(name, dgrad: reference chain vs oracle, product FP32 vs chain, product TF32 vs chain,
    wgrad: the same three)] -- for round 2; every piece here is unvalidated."""
    L = capi.lib()
    ops = RefOps(load_reference())
    from oracle import oracle as ora
    ora.build()
    tf32_was = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    rows = []
    try:
        for name, N, H, W, C, KH, KW, G in cases:
            rng = np.random.default_rng(77)
            OH, OW = H - KH + 1, W - KW + 1
            x = rng.standard_normal((N, H * W * C)).astype(np.float32)
            dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
            k = (rng.standard_normal((KH * KW * C, G)) * 0.05).astype(np.float32)
            xd, dyd, kd = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda(), torch.from_numpy(k).cuda()
            dx_ref = reference_conv_backprop(ops, dyd, kd, H, W, C, 0, 0, KH, KW, G)
            gw_ref, gb_ref = reference_conv_gradient(ops, xd, dyd, H, W, C, 0, 0, KH, KW, G)
            torch.cuda.synchronize()
            dx_o = ora.conv_backprop(dy, k, H, W, C, 0, 0, KH, KW, G, dtype=np.float64)
            gw_o = ora.conv_update(x, dy, k, np.zeros(G, np.float32), np.zeros_like(k), H, W, C, 0, 0, KH, KW, G,
                                   0.02, 0.0, 0.0, apply=False, dtype=np.float64)[3]
            rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))  # noqa: E731
            row = [name, rel(dx_ref.cpu().numpy(), dx_o)]
            for math in (0, 1):
                dx = torch.empty(N, H * W * C, device="cuda")
                L.cudaF_conv2d_dgrad(stream(), math, ptr(dyd), mdim(dyd), ptr(kd), mdim(kd), ptr(dx), mdim(dx),
                                     H, W, C, 0, 0, KH, KW, G)
                torch.cuda.synchronize()
                row.append(rel(dx.cpu().numpy(), dx_ref.cpu().numpy().astype(np.float64)))
            row.append(rel(gw_ref.cpu().numpy(), gw_o))
            for math in (0, 1):
                gw = torch.empty(KH * KW * C, G, device="cuda")
                gb = torch.empty(G, device="cuda")
                nbytes = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, 0, 0, KH, KW, G)
                ws = torch.empty(max(int(nbytes) // 4, 1), device="cuda")
                L.cudaF_conv2d_wgrad(stream(), math, ptr(xd), mdim(xd), ptr(dyd), mdim(dyd), ptr(gw), mdim(gw), ptr(gb),
                                     ptr(ws), H, W, C, 0, 0, KH, KW, G)
                torch.cuda.synchronize()
                row.append(rel(gw.cpu().numpy(), gw_ref.cpu().numpy().astype(np.float64)))
            rows.append(tuple(row))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32_was
    return rows


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
    if "--backward" in sys.argv:
        for row in check_backward([c for c in CASES if c[0] != "conv4"] + [("time", 64, 1, 14, 64, 1, 3, 128)]):
            print("%-6s dgrad: chain vs oracle %.2e  FP32 vs chain %.2e  TF32 vs chain %.2e | "
                  "wgrad: chain vs oracle %.2e  FP32 vs chain %.2e  TF32 vs chain %.2e" % row, flush=True)
        return
    ok = True
    for name, e_ref, e0, e1, times in check(CASES, timing=True):
        ok = ok and e_ref <= 1e-5 and e0 <= 1e-5 and e1 <= 1e-3
        print("%-6s reference(GPU kernels + cuBLAS) vs oracle(f64) %.2e | product FP32 vs reference %.2e | "
              "product TF32 vs reference %.2e | %s" % (name, e_ref, e0, e1, json.dumps(times)), flush=True)
    print("ref_conv_check", "ok" if ok else "FAILED")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
