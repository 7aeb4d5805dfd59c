"""The reference's own L0 kernels, recompiled for sm_100a, timed next to the product's on the same
buffers (NOT YET RUN ON A GPU: written at the end of round 1 when the GPU budget was spent; the calls
are those of tests/test_gpu_reference_kernels.py, which is green -- validate the numbers in round 2).

    python tests/ref_kernel_timing.py > gpurun_out/ref_kernel_timing.json      # on a GPU box

"Recompiled kernels are the baseline": for every HBM-bound member of the path
this prints algorithmic bytes, the median CUDA-event time (L2 flushed between
iterations) and GB/s of (i) the reference kernel with the reference's launch shape (16 x 16
threads over the output, conv2D.cc:221-580) and (ii) the product's launcher.  Lives under tests/
because it loads oracle/_ref (the checker), which nothing outside tests/ may do.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.capi import Dim3, mdim, ptr, stream  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libcnsl_ref_kernels.so")


def main():
    if not os.path.exists(REF_SO):
        raise SystemExit("build oracle/_ref first: make -C oracle ref (needs /root/reference)")
    capi.require_gpu()
    L = capi.lib()
    R = ctypes.CDLL(REF_SO)
    for name, args in capi._PROTOS.items():
        if name.startswith("cudaF_") and hasattr(R, name) and args and args[0] is Dim3:
            getattr(R, name).argtypes = args
            getattr(R, name).restype = None
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []

    def shape(t):
        r, c = t.shape
        return Dim3((c + 15) // 16, (r + 15) // 16, 1), Dim3(16, 16, 1)

    def med_ms(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(15):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    def both(name, byts, out_like, call):
        """call(lib, Gr, Bl) launches `name`-like work on either library (same legacy signature)."""
        gr, bl = shape(out_like)
        torch.cuda.synchronize()
        ref_ms = med_ms(lambda: call(R, gr, bl))          # legacy default stream == torch's current stream here
        L.kcnn_set_stream(stream())
        try:
            our_ms = med_ms(lambda: call(L, gr, bl))
        finally:
            L.kcnn_set_stream(None)
        rows.append({"launch": name, "algorithmic_bytes": byts, "reference_ms": ref_ms, "ours_ms": our_ms,
                     "reference_gbs": byts / (ref_ms * 1e-3) / 1e9, "ours_gbs": byts / (our_ms * 1e-3) / 1e9,
                     "speedup": ref_ms / our_ms})

    rnd = lambda r, c: torch.randn(r, c, device="cuda")  # noqa: E731
    emp = lambda r, c: torch.empty(r, c, device="cuda")  # noqa: E731

    # ---- max pooling at the C4 shapes and the model's own time pool
    for (H, W, C, ph, pw, pc, n) in ((1, 16, 2000, 1, 2, 1, 8192), (1, 8, 2000, 1, 2, 10, 8192),
                                     (1, 1, 4000, 1, 1, 5, 8192), (33, 9, 64, 3, 3, 2, 4096), (1, 12, 256, 1, 2, 1, 512)):
        ind, outd = H * W * C, (H // ph) * (W // pw) * (C // pc)
        x, y, dy, dx = rnd(n, ind), emp(n, outd), rnd(n, outd), torch.zeros(n, ind, device="cuda")
        tag = "%dx%dx%d pool %dx%dx%d N=%d" % (H, W, C, ph, pw, pc, n)
        both("maxpool_prop " + tag, 4 * n * (ind + outd), y,
             lambda lib, g, b: lib.cudaF_maxpool_prop(g, b, ptr(x), mdim(x), ptr(y), mdim(y), H, W, ph, pw, pc))
        # reference-exact backward WITHOUT the zero fill (the reference zero-fills in MaxpoolComponent::Backprop,
        # nnet0/nnet-component-nnet0.cc:889, a separate pass that neither side is charged for here):
        # reads in, out_value, out_deriv, writes one routed element per window
        both("maxpool_backprop(no zero fill) " + tag, 4 * n * (ind + 3 * outd), dy,
             lambda lib, g, b: lib.cudaF_maxpool_backprop(g, b, ptr(x), mdim(x), ptr(y), mdim(y), ptr(dy), mdim(dy),
                                                          ptr(dx), mdim(dx), H, W, ph, pw, pc))
        del x, y, dy, dx

    # ---- data-movement members (conv2D.cc:213-463)
    N, C, bs = 4096, 256, 14
    x, o = rnd(N, C * bs), emp(C, N * bs)
    both("tp_block [%dx%d] C=%d bs=%d" % (N, C * bs, C, bs), 8 * N * C * bs, o,
         lambda lib, g, b: lib.cudaF_tp_block(g, b, ptr(x), mdim(x), ptr(o), mdim(o), bs))
    G, bs = 256, 12
    x, o = rnd(N, G * bs), emp(N * bs, G)
    both("tp_inside_block [%dx%d] G=%d bs=%d" % (N, G * bs, G, bs), 8 * N * G * bs, o,
         lambda lib, g, b: lib.cudaF_tp_inside_block(g, b, ptr(x), mdim(x), ptr(o), mdim(o), bs))
    v = rnd(1, G)
    both("add_mat_rep_vec [%dx%d] rep=%d" % (N, G * bs, bs), 8 * N * G * bs, x,
         lambda lib, g, b: lib.cudaF_add_mat_rep_vec(g, b, ptr(v), bs, ptr(x), mdim(x)))
    C, KW, G = 2000, 5, 2000
    k, o = rnd(C * KW, G), emp(C * KW, G)
    both("mod_permute_row [%dx%d] C=%d bs=%d" % (C * KW, G, C, KW), 8 * C * KW * G, o,
         lambda lib, g, b: lib.cudaF_mod_permute_row(g, b, ptr(k), mdim(k), ptr(o), mdim(o), KW, C))
    f = emp(KW * G, C)
    both("flip_mat KH=1 KW=%d C=%d G=%d" % (KW, C, G), 8 * C * KW * G, f,
         lambda lib, g, b: lib.cudaF_flip_mat(g, b, ptr(k), mdim(k), 1, KW, G, ptr(f), mdim(f)))
    del k, o, f
    N, H, W, C, KH, KW = 4096, 1, 14, 256, 1, 3
    x = rnd(N, H * W * C)
    PW = W + 2 * (KW - 1)
    p = emp(N, H * PW * C)
    both("pad_zero [%dx%d] -> [%dx%d]" % (N, H * W * C, N, H * PW * C), 4 * N * C * H * (W + PW), p,
         lambda lib, g, b: lib.cudaF_pad_zero(g, b, ptr(x), mdim(x), H, W, KH, KW, ptr(p), mdim(p)))
    print(json.dumps(rows, indent=1))


if __name__ == "__main__":
    main()
