"""Developer probe for the fully connected GEMMs at the benchmarked sizes (M = 512): error against a
float64 product and time per launch (CUDA events, L2 flushed between launches), for the tile variant the
environment selects (KCNN_TMA_PAIRK=0/1 ...).  Not a test.

    python tools/fc_probe.py            # numerics + timing with the current environment
    python tools/fc_probe.py --ab       # runs itself with KCNN_TMA_PAIRK=0 and =1
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if "--ab" in sys.argv:
    for v in ("0", "1"):
        env = dict(os.environ, KCNN_TMA_PAIRK=v)
        print("== KCNN_TMA_PAIRK=%s" % v, flush=True)
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, timeout=300)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.capi import mdim, ptr, stream  # noqa: E402

L = capi.lib()
torch.cuda.init()
g = torch.Generator(device="cuda")
g.manual_seed(1)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush_buf.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / iters * 1e3


def rel(a, r):
    return float((a.double() - r).abs().max() / r.abs().max())


def pitched(r, c, scale=1.0):
    ld = (c + 3) // 4 * 4
    return (torch.randn(r, ld, device="cuda", generator=g) * scale)[:, :c]


SHAPES = ((512, 4096, 4096), (512, 4096, 3454), (512, 1024, 4096), (256, 4096, 4096), (384, 2048, 1000))
if "--one" in sys.argv:          # ncu target: two launches of each FC2 GEMM, nothing else
    SHAPES = SHAPES[:1]
    timeit = lambda fn, iters=0: (fn(), fn(), torch.cuda.synchronize(), 1.0)[-1]
for (N, din, dout) in SHAPES:
    x, w, dy = pitched(N, din), pitched(dout, din, 0.02), pitched(N, dout)
    b = torch.randn(dout, device="cuda", generator=g)
    y = torch.empty(N, (dout + 3) // 4 * 4, device="cuda")[:, :dout]
    dx = torch.empty(N, din, device="cuda")
    f = lambda: L.cudaF_affine_fprop(stream(), 1, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y))
    d = lambda: L.cudaF_affine_dgrad(stream(), 1, ptr(dy), mdim(dy), ptr(w), mdim(w), ptr(dx), mdim(dx))
    f(); d()
    torch.cuda.synchronize()
    e1 = rel(y, x.double() @ w.double().t() + b.double())
    e2 = rel(dx, dy.double() @ w.double())
    # fused epilogues: ReLU + dropout forward, gates backward
    y2 = torch.empty_like(y)
    mx, my = torch.relu(pitched(N, din)), None
    my = mx * (torch.rand(N, din, device="cuda", generator=g) > 0.5) * 2.0
    t1, t2 = timeit(f), timeit(d)
    fl = 2.0 * N * din * dout
    line = "N=%d %d->%d  fprop err %.1e %.1f us (%.0f TF/s)   dgrad err %.1e %.1f us (%.0f TF/s)" % (
        N, din, dout, e1, t1, fl / t1 / 1e6, e2, t2, fl / t2 / 1e6)
    if hasattr(L, "cudaF_affine_dgrad_fused") and din % 4 == 0:
        dg = lambda: L.cudaF_affine_dgrad_fused(stream(), ptr(dy), mdim(dy), ptr(w), mdim(w), ptr(dx), mdim(dx),
                                                ptr(mx), mdim(mx).stride, ptr(my), mdim(my).stride, 0)
        try:
            ok = dg()
            torch.cuda.synchronize()
            want = torch.where(mx > 0, (dy.double() @ w.double()) * (my.double() / mx.double().clamp_min(1e-30)), 0.0)
            e3 = rel(dx, want)
            t3 = timeit(dg)
            line += "   gated dgrad (ok=%s) err %.1e %.1f us" % (ok, e3, t3)
        except Exception as ex:  # signature drift: this is a probe
            line += "   gated dgrad: %s" % ex
    print(line, flush=True)
