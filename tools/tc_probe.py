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


def affine(N, din, dout, math):
    x = rng.standard_normal((N, din)).astype(np.float32)
    w = (rng.standard_normal((dout, din)) * 0.05).astype(np.float32)
    b = rng.standard_normal(dout).astype(np.float32)
    dy = rng.standard_normal((N, dout)).astype(np.float32)
    xd, wd, bd, dyd = (torch.from_numpy(a).cuda() for a in (x, w, b, dy))
    y = torch.full((N, dout), float("nan"), device="cuda")
    L.cudaF_affine_fprop(stream(), math, ptr(xd), mdim(xd), ptr(wd), mdim(wd), ptr(bd), ptr(y), mdim(y))
    torch.cuda.synchronize()
    e1 = rel(y.cpu().numpy(), x.astype(np.float64) @ w.T.astype(np.float64) + b)
    dx = torch.full((N, din), float("nan"), device="cuda")
    L.cudaF_affine_dgrad(stream(), math, ptr(dyd), mdim(dyd), ptr(wd), mdim(wd), ptr(dx), mdim(dx))
    torch.cuda.synchronize()
    e2 = rel(dx.cpu().numpy(), dy.astype(np.float64) @ w.astype(np.float64))
    g = torch.full((dout, din), float("nan"), device="cuda")
    bg = torch.full((dout,), float("nan"), device="cuda")
    L.cudaF_affine_wgrad(stream(), math, ptr(xd), mdim(xd), ptr(dyd), mdim(dyd), ptr(g), mdim(g), ptr(bg))
    torch.cuda.synchronize()
    e3 = rel(g.cpu().numpy(), dy.T.astype(np.float64) @ x.astype(np.float64))
    print("affine N=%d %d->%d math=%d  fprop %.2e dgrad %.2e wgrad %.2e" % (N, din, dout, math, e1, e2, e3), flush=True)


for shape in [(128, 32, 128), (128, 64, 128), (256, 256, 1024), (33, 70, 130), (1, 1, 1), (512, 1024, 4096)]:
    affine(*shape, 1)
affine(256, 256, 1024, 0)
