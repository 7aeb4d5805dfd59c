"""One launch of every HBM-bound kernel of the path (max-pool forward / backward, FlipMat,
PaddingZero, TpBlock, TpInsideBlock, ModPermuteRow, AddMatRepVec, the channels-last staging pack
and the SGD update) at bandwidth-relevant sizes: the target of

    ncu --set full --clock-control none --import-source on -o gpurun_out/hbm_rNN python tools/hbm_probe.py

(BASELINE.json: "achieved HBM GB/s for pooling and permutes ... evidenced by ncu").  Prints, in
launch order, the ALGORITHMIC bytes of each launch (4 x (elements read + elements
written)) as JSON, so tools/ncu_summarise.py can set them against dram__bytes_{read,write}.sum.
With --time it also reports CUDA-event timings (median of 20, L2 flushed) without ncu.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.capi import mdim, ptr, stream  # noqa: E402

L = capi.lib()
capi.require_gpu()
TIME = "--time" in sys.argv
FLUSH = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda") if TIME else None
rows = []


def run(name, byts, fn):
    ms = None
    if TIME:
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            FLUSH.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
    else:
        fn()
        torch.cuda.synchronize()
    rows.append({"launch": name, "algorithmic_bytes": byts, "ms": ms,
                 "gbs": (byts / (ms * 1e-3) / 1e9) if ms else None})


def rnd(r, c):
    return torch.randn(r, c, device="cuda")


def emp(r, c):
    return torch.empty(r, c, device="cuda")


# ---- max pooling: the C4 sweep shapes at the large-batch end
for (H, W, C, ph, pw, pc, n) in ((1, 16, 2000, 1, 2, 1, 8192), (1, 8, 2000, 1, 2, 10, 8192), (1, 1, 4000, 1, 1, 5, 8192),
                                 (33, 9, 64, 3, 3, 2, 4096), (1, 12, 256, 1, 2, 1, 512)):
    ind, outd = H * W * C, (H // ph) * (W // pw) * (C // pc)
    x, y, dy, dx = rnd(n, ind), emp(n, outd), rnd(n, outd), emp(n, ind)
    tag = "%dx%dx%d pool %dx%dx%d N=%d" % (H, W, C, ph, pw, pc, n)
    run("maxpool_prop " + tag, 4 * n * (ind + outd),
        lambda: L.cudaF_maxpool_prop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), H, W, ph, pw, pc, 0))
    run("maxpool_backprop(exact, zero-fill fused) " + tag, 4 * n * (2 * ind + 2 * outd),
        lambda: L.cudaF_maxpool_backprop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), ptr(dy), mdim(dy), ptr(dx), mdim(dx),
                                           H, W, ph, pw, pc, 0, 1))
    del x, y, dy, dx

# ---- data-movement members (cnslmat/conv2D.cc:213-463) at conv4 / C3(iii) sizes
N, C, bs = 4096, 256, 14
x, o = rnd(N, C * bs), emp(C, N * bs)
run("tp_block [%dx%d] C=%d bs=%d" % (N, C * bs, C, bs), 8 * N * C * bs,
    lambda: L.cudaF_tp_block_s(stream(), ptr(x), mdim(x), ptr(o), mdim(o), bs))
G, bs = 256, 12
x, o = rnd(N, G * bs), emp(N * bs, G)
run("tp_inside_block [%dx%d] G=%d bs=%d" % (N, G * bs, G, bs), 8 * N * G * bs,
    lambda: L.cudaF_tp_inside_block_s(stream(), ptr(x), mdim(x), ptr(o), mdim(o), bs))
v = rnd(1, G)
run("add_mat_rep_vec [%dx%d] rep=%d" % (N, G * bs, bs), 8 * N * G * bs,
    lambda: L.cudaF_add_mat_rep_vec_s(stream(), ptr(v), bs, ptr(x), mdim(x)))
C, KW, G = 2000, 5, 2000
k, o = rnd(C * KW, G), emp(C * KW, G)
run("mod_permute_row [%dx%d] C=%d bs=%d" % (C * KW, G, C, KW), 8 * C * KW * G,
    lambda: L.cudaF_mod_permute_row_s(stream(), ptr(k), mdim(k), ptr(o), mdim(o), KW, C))
f = emp(KW * G, C)
run("flip_mat KH=1 KW=%d C=%d G=%d" % (KW, C, G), 8 * C * KW * G,
    lambda: L.cudaF_flip_mat_s(stream(), ptr(k), mdim(k), 1, KW, G, ptr(f), mdim(f)))
del k, o, f
N, H, W, C, KH, KW = 4096, 1, 14, 256, 1, 3
x = rnd(N, H * W * C)
PW = W + 2 * (KW - 1)
p = emp(N, H * PW * C)
run("pad_zero [%dx%d] -> [%dx%d]" % (N, H * W * C, N, H * PW * C), 4 * N * C * H * (W + PW),
    lambda: L.cudaF_pad_zero_s(stream(), ptr(x), mdim(x), H, W, KH, KW, ptr(p), mdim(p)))
del x, p

# ---- SGD update (nnet0/nnet-component-nnet0.cc:767-775): 3 reads + 2 writes per weight
w, pv, g = rnd(4096, 4096), rnd(4096, 4096), rnd(4096, 4096)
run("sgd_momentum_update [4096x4096]", 20 * 4096 * 4096,
    lambda: L.cudaF_sgd_momentum_update(stream(), ptr(w), mdim(w), ptr(pv), mdim(pv), ptr(g), mdim(g), 0.9, -1e-9, 1e-9))

print(json.dumps(rows, indent=1))
