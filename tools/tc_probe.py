"""Developer probe: run a few TF32 tensor-core GEMMs through the C-ABI and print the error
against a float64 numpy product (not a test; used while bringing kernels up on a GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kaldi_cnn_b200 import capi
from kaldi_cnn_b200.capi import mdim, ptr, stream

L = capi.lib()
torch.cuda.init()
rng = np.random.default_rng(0)


def rel(got, ref):
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))


def pitched(a):
    """CuMatrix-like device copy: row pitch rounded up to 4 floats (16 bytes)."""
    r, c = a.shape
    ld = (c + 3) // 4 * 4
    buf = torch.full((r, ld), float("nan"), device="cuda")
    v = buf[:, :c]
    v.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return v


def affine(N, din, dout, math, relu=False):
    x = rng.standard_normal((N, din)).astype(np.float32)
    if relu:
        x = np.maximum(x, 0)
    w = (rng.standard_normal((dout, din)) * 0.05).astype(np.float32)
    b = rng.standard_normal(dout).astype(np.float32)
    dy = rng.standard_normal((N, dout)).astype(np.float32)
    xd, wd, dyd = (pitched(a) for a in (x, w, dy))
    bd = torch.from_numpy(b).cuda()
    y = pitched(np.full((N, dout), np.nan, dtype=np.float32))
    L.cudaF_affine_fprop(stream(), math, ptr(xd), mdim(xd), ptr(wd), mdim(wd), ptr(bd), ptr(y), mdim(y))
    torch.cuda.synchronize()
    e1 = rel(y.cpu().numpy(), x.astype(np.float64) @ w.T.astype(np.float64) + b)
    dx = pitched(np.full((N, din), np.nan, dtype=np.float32))
    L.cudaF_affine_dgrad(stream(), math, ptr(dyd), mdim(dyd), ptr(wd), mdim(wd), ptr(dx), mdim(dx))
    torch.cuda.synchronize()
    e2 = rel(dx.cpu().numpy(), dy.astype(np.float64) @ w.astype(np.float64))
    g = pitched(np.full((dout, din), np.nan, dtype=np.float32))
    bg = torch.full((dout,), float("nan"), device="cuda")
    L.cudaF_affine_wgrad(stream(), math, ptr(xd), mdim(xd), ptr(dyd), mdim(dyd), ptr(g), mdim(g), ptr(bg))
    torch.cuda.synchronize()
    e3 = rel(g.cpu().numpy(), dy.T.astype(np.float64) @ x.astype(np.float64))
    print("affine N=%d %d->%d math=%d  fprop %.2e dgrad %.2e wgrad %.2e" % (N, din, dout, math, e1, e2, e3), flush=True)


if "--ncu" not in sys.argv and "--perf" not in sys.argv:
    for shape in [(128, 32, 128), (128, 64, 128), (256, 256, 1024), (33, 70, 130), (1, 1, 1), (512, 1024, 4096),
                  (512, 4096, 3454), (100, 1056, 1024), (37, 8, 4), (512, 4096, 4096)]:
        affine(*shape, 1)
    affine(512, 4096, 4096, 1, relu=True)
    affine(256, 256, 1024, 0)


def timeit(fn, iters=20):
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


def perf(N=512):
    for math in (1,):
        for (din, dout) in ((1024, 4096), (4096, 4096), (4096, 3454)):
            x = torch.randn(N, din, device="cuda"); w = torch.randn(dout, din, device="cuda") * 0.01
            b = torch.zeros(dout, device="cuda"); y = pitched(np.zeros((N, dout), dtype=np.float32))
            g = torch.empty(dout, din, device="cuda"); bg = torch.empty(dout, device="cuda")
            fl = 2.0 * N * din * dout
            t1 = timeit(lambda: L.cudaF_affine_fprop(stream(), math, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y)))
            t2 = timeit(lambda: L.cudaF_affine_dgrad(stream(), math, ptr(y), mdim(y), ptr(w), mdim(w), ptr(x), mdim(x)))
            t3 = timeit(lambda: L.cudaF_affine_wgrad(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(g), mdim(g), ptr(bg)))
            pv = torch.zeros(dout, din, device="cuda")
            t4 = timeit(lambda: L.cudaF_affine_wgrad_sgd(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(w), mdim(w), ptr(pv), mdim(pv), ptr(b), 0.9, -1e-9, 1e-9))
            print("   fused wgrad+sgd %.1f us" % (t4 * 1e3))
            print("FC %d->%d N=%d math=%d: fprop %.1f us (%.0f TF/s)  dgrad %.1f us (%.0f)  wgrad %.1f us (%.0f)" % (
                din, dout, N, math, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9, t3 * 1e3, fl / t3 / 1e9), flush=True)
        convs = [(40, 21, 1, 40, 4, 128), (1, 18, 64, 1, 3, 128), (1, 16, 128, 1, 3, 256), (1, 14, 256, 1, 3, 256),
                 (1, 6, 256, 1, 3, 512), (1, 4, 512, 1, 3, 512)]
        for (H, W, C, KH, KW, G) in convs:
            OH, OW = H - KH + 1, W - KW + 1
            xi = torch.randn(N, H * W * C, device="cuda"); k = torch.randn(KH * KW * C, G, device="cuda") * 0.01
            bb = torch.zeros(G, device="cuda"); yo = torch.randn(N, OH * OW * G, device="cuda")
            kg = torch.empty(KH * KW * C, G, device="cuda"); bgr = torch.empty(G, device="cuda")
            nb = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, 0, 0, KH, KW, G)
            ws = torch.empty(max(nb, 4) // 4, device="cuda")
            fl = 2.0 * N * OH * OW * G * KH * KW * C
            t1 = timeit(lambda: L.cudaF_conv2d_fprop(stream(), math, ptr(xi), mdim(xi), ptr(k), mdim(k), ptr(bb), ptr(yo), mdim(yo), H, W, C, 0, 0, KH, KW, G, 1))
            t2 = timeit(lambda: L.cudaF_conv2d_dgrad(stream(), math, ptr(yo), mdim(yo), ptr(k), mdim(k), ptr(xi), mdim(xi), H, W, C, 0, 0, KH, KW, G))
            t3 = timeit(lambda: L.cudaF_conv2d_wgrad(stream(), math, ptr(xi), mdim(xi), ptr(yo), mdim(yo), ptr(kg), mdim(kg), ptr(bgr), ptr(ws), H, W, C, 0, 0, KH, KW, G))
            print("conv %dx%dx%d k%dx%d G%d: fprop %.1f us (%.0f TF/s)  dgrad %.1f us (%.0f)  wgrad+bias %.1f us (%.0f)" % (
                H, W, C, KH, KW, G, t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9, t3 * 1e3, fl / t3 / 1e9), flush=True)


if "--perf" in sys.argv:
    perf()


def ncu_target(N=512):
    """A short launch list for `ncu`: FC2 fprop / dgrad / wgrad and conv4 fprop / dgrad / wgrad, TF32."""
    math = 1
    din = dout = 4096
    x = torch.randn(N, din, device="cuda"); w = torch.randn(dout, din, device="cuda") * 0.01
    b = torch.zeros(dout, device="cuda"); y = torch.empty(N, dout, device="cuda")
    g = torch.empty(dout, din, device="cuda"); bg = torch.empty(dout, device="cuda")
    pv = torch.zeros(dout, din, device="cuda")
    for _ in range(2):
        L.cudaF_affine_fprop(stream(), math, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y))
        L.cudaF_affine_dgrad(stream(), math, ptr(y), mdim(y), ptr(w), mdim(w), ptr(x), mdim(x))
        L.cudaF_affine_wgrad(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(g), mdim(g), ptr(bg))
        L.cudaF_affine_wgrad_sgd(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(w), mdim(w), ptr(pv), mdim(pv),
                                 ptr(b), 0.9, -1e-9, 1e-9)
    H, W, C, KH, KW, G = 1, 14, 256, 1, 3, 256
    OW = W - KW + 1
    xi = torch.randn(N, W * C, device="cuda"); k = torch.randn(KW * C, G, device="cuda") * 0.01
    bb = torch.zeros(G, device="cuda"); yo = torch.randn(N, OW * G, device="cuda")
    kg = torch.empty(KW * C, G, device="cuda"); bgr = torch.empty(G, device="cuda")
    nb = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, 0, 0, KH, KW, G)
    ws = torch.empty(max(nb, 4) // 4, device="cuda")
    for _ in range(2):
        L.cudaF_conv2d_fprop(stream(), math, ptr(xi), mdim(xi), ptr(k), mdim(k), ptr(bb), ptr(yo), mdim(yo), H, W, C, 0, 0, KH, KW, G, 1)
        L.cudaF_conv2d_dgrad(stream(), math, ptr(yo), mdim(yo), ptr(k), mdim(k), ptr(xi), mdim(xi), H, W, C, 0, 0, KH, KW, G)
        L.cudaF_conv2d_wgrad(stream(), math, ptr(xi), mdim(xi), ptr(yo), mdim(yo), ptr(kg), mdim(kg), ptr(bgr), ptr(ws), H, W, C, 0, 0, KH, KW, G)
    torch.cuda.synchronize()
    print("ncu target done")


if "--ncu" in sys.argv:
    ncu_target()


def big(N=4096):
    """Main-loop-dominated GEMMs (K = 4096 per tile) to compare tile variants."""
    math = 1
    din = dout = 4096
    x = torch.randn(N, din, device="cuda"); w = torch.randn(dout, din, device="cuda") * 0.01
    b = torch.zeros(dout, device="cuda"); y = torch.empty(N, dout, device="cuda")
    g = torch.empty(dout, din, device="cuda"); bg = torch.empty(dout, device="cuda")
    fl = 2.0 * N * din * dout
    t1 = timeit(lambda: L.cudaF_affine_fprop(stream(), math, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y)))
    t2 = timeit(lambda: L.cudaF_affine_dgrad(stream(), math, ptr(y), mdim(y), ptr(w), mdim(w), ptr(x), mdim(x)))
    t3 = timeit(lambda: L.cudaF_affine_wgrad(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(g), mdim(g), ptr(bg)))
    print("GEMM %d^3 pair=%s: KK %.1f us (%.0f TF/s)  K,MN %.1f us (%.0f)  MN,MN %.1f us (%.0f)" % (
        N, os.environ.get("KCNN_TMA_PAIR", "1"), t1 * 1e3, fl / t1 / 1e9, t2 * 1e3, fl / t2 / 1e9, t3 * 1e3, fl / t3 / 1e9), flush=True)
    a = torch.randn(4096, 4096, device="cuda"); bb = torch.randn(4096, 4096, device="cuda")
    torch.backends.cuda.matmul.allow_tf32 = True
    t4 = timeit(lambda: torch.matmul(a, bb))
    print("cuBLAS TF32 4096^3: %.1f us (%.0f TF/s)" % (t4 * 1e3, 2 * 4096.0 ** 3 / t4 / 1e9), flush=True)


if "--big" in sys.argv:
    big()
