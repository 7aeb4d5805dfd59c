"""Measurement of the BASELINE.json configurations that bench.py does not put on its JSON line
(SURVEY 8d): C1 (single conv + maxpool + FC, GPU and CPU oracle), C3 (large in_channel x many
filters: fprop / dgrad / wgrad separately), C4 (max-pool bandwidth sweep over batch size).

    python tools/bench_configs.py [--out profiles/r01_configs.json] [--quick]

GPU timings: CUDA events on the launching stream, 5 warm-ups, median of 20, a 160 MB buffer
zeroed between iterations (L2 flush).  CPU timings: the oracle (oracle/, "port") on a bounded
sample.  Rooflines: tensor = measured BF16 dense / 2 (TF32), HBM = measured copy bandwidth
(MEASURED_PEAKS.json).  Algorithmic work as defined in BASELINE.md section 3.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from kaldi_cnn_b200 import capi, components as kc  # noqa: E402
from kaldi_cnn_b200.capi import mdim, ptr, stream  # noqa: E402

L = capi.lib()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"] / 2.0, "measured"
    return 6650.0, 1590.0 / 2.0, "fallback"


HBM, TF32, SRC = peaks()
FLUSH = None


def gpu_ms(fn, iters=20, warm=5):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def pitched(rows, cols, fill=None):
    ld = (cols + 3) // 4 * 4
    buf = torch.randn(rows, ld, device="cuda") if fill is None else torch.full((rows, ld), fill, device="cuda")
    return buf[:, :cols]


def tensor_entry(name, flops, ms):
    tf = flops / (ms * 1e-3) / 1e12
    return {"kernel": name, "ms": ms, "bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "peak": TF32,
            "frac": tf / TF32, "peak_source": SRC + ": bf16 dense / 2"}


def hbm_entry(name, byts, ms):
    gb = byts / (ms * 1e-3) / 1e9
    return {"kernel": name, "ms": ms, "bound": "hbm", "achieved": gb, "unit": "GB/s", "peak": HBM, "frac": gb / HBM,
            "peak_source": SRC}


# ------------------------------------------------------------------------------ C3 --

def conv_three_passes(N, H, W, C, KH, KW, G, math=1):
    OH, OW = H - KH + 1, W - KW + 1
    x = pitched(N, H * W * C); k = pitched(KH * KW * C, G); k.mul_(0.01)
    b = torch.zeros(G, device="cuda"); y = pitched(N, OH * OW * G)
    dx = pitched(N, H * W * C); kg = pitched(KH * KW * C, G); bg = torch.empty(G, device="cuda")
    nb = L.kcnn_conv2d_wgrad_workspace(N, H, W, C, 0, 0, KH, KW, G)
    ws = torch.empty(max(nb, 4) // 4, device="cuda")
    fl = 2.0 * N * OH * OW * G * KH * KW * C
    tag = "%dx%dx%d k%dx%d G%d N%d" % (H, W, C, KH, KW, G, N)
    out = []
    out.append(tensor_entry("conv2d_fprop " + tag, fl, gpu_ms(lambda: L.cudaF_conv2d_fprop(
        stream(), math, ptr(x), mdim(x), ptr(k), mdim(k), ptr(b), ptr(y), mdim(y), H, W, C, 0, 0, KH, KW, G, 1))))
    out.append(tensor_entry("conv2d_dgrad " + tag, fl, gpu_ms(lambda: L.cudaF_conv2d_dgrad(
        stream(), math, ptr(y), mdim(y), ptr(k), mdim(k), ptr(dx), mdim(dx), H, W, C, 0, 0, KH, KW, G))))
    out.append(tensor_entry("conv2d_wgrad+bias " + tag, fl, gpu_ms(lambda: L.cudaF_conv2d_wgrad(
        stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(kg), mdim(kg), ptr(bg), ptr(ws), H, W, C, 0, 0, KH, KW, G))))
    return out


# ------------------------------------------------------------------------------ C4 --

def maxpool_sweep(H, W, C, ph, pw, pc, batches):
    ind, outd = H * W * C, (H // ph) * (W // pw) * (C // pc)
    out = []
    for n in batches:
        x = pitched(n, ind); y = pitched(n, outd); dy = pitched(n, outd); dx = pitched(n, ind)
        tag = "%dx%dx%d pool %dx%dx%d N=%d" % (H, W, C, ph, pw, pc, n)
        ms = gpu_ms(lambda: L.cudaF_maxpool_prop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), H, W, ph, pw, pc, 0))
        out.append(hbm_entry("maxpool_prop " + tag, 4.0 * n * (ind + outd), ms))
        ms = gpu_ms(lambda: L.cudaF_maxpool_backprop_s(stream(), ptr(x), mdim(x), ptr(y), mdim(y), ptr(dy), mdim(dy),
                                                       ptr(dx), mdim(dx), H, W, ph, pw, pc, 0, 1))
        out.append(hbm_entry("maxpool_backprop(exact) " + tag, 4.0 * n * (2 * ind + 2 * outd), ms))
    return out


# ------------------------------------------------------------------------------ C1 --

C1 = {
    "C1a": ["ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=40 kernel-width=4 stride=1 "
            "group=128 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5",
            "MaxpoolComponent in-height=1 in-width=8 in-channel=128 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=2",
            "FullyConnectedComponent input-dim=256 output-dim=1024 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.1"],
    "C1b": ["ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=8 kernel-width=3 stride=1 "
            "group=64 out-height=33 out-width=9 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5",
            "MaxpoolComponent in-height=33 in-width=9 in-channel=64 pool-height-dim=3 pool-width-dim=3 pool-channel-dim=2",
            "FullyConnectedComponent input-dim=1056 output-dim=1024 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.1"],
}


def c1_gpu(lines, N, math):
    kc.set_math_mode(math)
    kc.set_rand_seed(42)
    comps = [kc.Component.from_string(l) for l in lines]
    x = pitched(N, comps[0].input_dim)
    acts = [x] + [pitched(N, c.output_dim) for c in comps]
    dyl = pitched(N, comps[-1].output_dim)
    derivs = [pitched(N, c.input_dim) for c in comps]

    def step():
        for i, c in enumerate(comps):
            c.propagate(acts[i], acts[i + 1])
        d = dyl
        for i in range(len(comps) - 1, -1, -1):
            comps[i].backprop(acts[i], acts[i + 1], d, in_deriv=derivs[i])
            d = derivs[i]
    ms = gpu_ms(step)
    return {"ms_per_step": ms, "frames_per_sec": N / (ms * 1e-3)}


def c1_cpu(lines, N, threads):
    """C1 on the host cores, through bench.py's cpu_baseline leg (the one place outside tests/ that
    runs the CPU port under oracle/)."""
    import bench
    cfg = "\n".join(lines) + "\nSoftmaxComponent dim=1024\n"
    fps, sec, kind, used, how = bench.cpu_train_frames_per_sec(cfg, N, 2, 1, threads)
    return {"ms_per_step": sec * 1e3, "frames_per_sec": fps, "cores": used, "kind": kind,
            "note": "CPU port step incl. a softmax / cross-entropy tail the GPU timing does not have; " + how}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_configs.json"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma list of C1,C3,C4")
    args = ap.parse_args()
    capi.require_gpu()
    res = {"peaks": {"hbm_gbs": HBM, "tf32_tflops": TF32, "source": SRC}}
    with torch.cuda.stream(torch.cuda.Stream()):
        kc.use_current_stream()
        only = set(x for x in args.only.split(",") if x)
        res["C1"] = {}
        for name, lines in (C1.items() if (not only or "C1" in only) else []):
            res["C1"][name] = {"N": 256, "gpu_tf32": c1_gpu(lines, 256, 1), "gpu_fp32": c1_gpu(lines, 256, 0),
                               "cpu": c1_cpu(lines, 64 if args.quick else 256, os.cpu_count() or 1)}
        res["C3"] = []
        shapes = [(256, 1, 4, 512, 1, 3, 512), (256, 1, 14, 256, 1, 3, 256)]
        if not args.quick:
            shapes.append((256, 1, 8, 2000, 1, 5, 2000))
        if only and "C3" not in only:
            shapes = []
        for (N, H, W, C, KH, KW, G) in shapes:
            res["C3"] += conv_three_passes(N, H, W, C, KH, KW, G)
        res["C4"] = []
        batches = [64, 512, 4096] if args.quick else [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384]
        if only and "C4" not in only:
            batches = []
        for (H, W, C, ph, pw, pc) in ((1, 16, 2000, 1, 2, 1), (1, 1, 4000, 1, 1, 5), (1, 8, 2000, 1, 2, 10),
                                      (33, 9, 64, 3, 3, 2)):
            res["C4"] += maxpool_sweep(H, W, C, ph, pw, pc, batches)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    for name, v in res["C1"].items():
        print("%s N=256: GPU tf32 %.0f frames/s (%.3f ms)  fp32 %.0f frames/s  | CPU oracle %.0f frames/s on %d threads" % (
            name, v["gpu_tf32"]["frames_per_sec"], v["gpu_tf32"]["ms_per_step"], v["gpu_fp32"]["frames_per_sec"],
            v["cpu"]["frames_per_sec"], v["cpu"]["cores"]))
    for e in res["C3"] + res["C4"]:
        print("%-62s %8.4f ms  %9.1f %s  %.3f of %s peak" % (e["kernel"], e["ms"], e["achieved"], e["unit"], e["frac"], e["bound"]))


if __name__ == "__main__":
    main()
