#!/usr/bin/env python
"""bench.py -- CNN training throughput (frames/sec) of the kaldi-cnn hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one training minibatch through the hot path: Propagate through every component of the
model, cross-entropy objective + derivative, Backprop in reverse with the parameter update inside
Backprop (momentum / weight-decay SGD), exactly the loop nnet2's NnetUpdater runs.  One frame = one
minibatch row (one labelled frame with its 21-frame context window).

Workload (config.workload): BASELINE.json configs[1] -- the reference's time-axis deep CNN
(egs/exp/nnet/nnet.config: SpliceComponent over 21 frames of 40 mel bins, 6 conv + pool + 3 FC + softmax
3454) with an intermap max-pool after conv1 (pool-channel-dim=2, the egs/local/nnet0/run_conv.sh shape),
per-GPU minibatch 512 (the recipes' GPU minibatch, egs/local/nnet0/run_nnet.sh:19-20), synthetic N(0,1)
filterbank frames (512 x 21 rows of 40 per step), random-init weights.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      the kernel with the largest measured share of the step (per-launch CUDA events over an
                eager replica of the timed step, kcnn_profile_*), against its binding roofline
  step_kernels  every launch group of that step: share, algorithmic FLOPs / bytes, achieved, fraction
  kernels       isolated, L2-flushed rooflines of the kernels BASELINE's metric names (conv tensor pipe,
                max-pool HBM GB/s) at the model's and the C4 sweep's shapes
  configs       the other BASELINE configs (C1a, C1b: GPU + CPU leg; C3; C4) measured in this run
  cpu_baseline  the reference's CPU path (the oracle port) timed on this box's host cores on a bounded sample
  rank_parity   N > 1: P x 32 rows against 1 x 32P rows through the same trainer (parameters, momentum)
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    "c2-intermap": ("nnet_c2_intermap.config",
                    "C2 time-axis deep CNN with intermap pooling: egs/exp/nnet/nnet.config + "
                    "MaxpoolComponent pool-channel-dim=2 after conv1 (Splice 21 x 40, 6 conv, 2 maxpool, 3 FC, "
                    "softmax 3454)"),
    "c2": ("nnet_c2.config",
           "C2: egs/exp/nnet/nnet.config verbatim (Splice 21 x 40, 6 conv, time max-pool, 3 FC, softmax 3454)"),
}


def load_config(name):
    path = os.path.join(ROOT, "kaldi-cnn_b200", "configs", WORKLOADS[name][0])
    return open(path).read()


def parse_config(text):
    """[(component type, {key: value})] of an nnet.config (one component per line)."""
    layers = []
    for line in text.replace("\r", "").split("\n"):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        parts = line.split()
        layers.append((parts[0], dict(p.split("=", 1) for p in parts[1:])))
    return layers


def model_flops_per_frame(cfg_text):
    """Algorithmic training FLOPs per frame: 3 x forward GEMM FLOPs (fprop, dgrad, wgrad) of the
    conv and FC layers."""
    macs = 0
    for kind, kv in parse_config(cfg_text):
        if kind == "ConvolutionComponent":
            g = lambda k, d=0: int(kv.get(k, d))
            oh = g("in-height") + 2 * g("in-pad-height") - g("kernel-height") + 1
            ow = g("in-width") + 2 * g("in-pad-width") - g("kernel-width") + 1
            macs += oh * ow * g("group") * g("kernel-height") * g("kernel-width") * g("in-channel")
        elif kind == "FullyConnectedComponent":
            macs += int(kv["input-dim"]) * int(kv["output-dim"])
    return 2 * macs * 3


def param_count(cfg_text):
    n = 0
    for kind, kv in parse_config(cfg_text):
        if kind == "ConvolutionComponent":
            n += (int(kv["kernel-height"]) * int(kv["kernel-width"]) * int(kv["in-channel"]) + 1) * int(kv["group"])
        elif kind == "FullyConnectedComponent":
            n += (int(kv["input-dim"]) + 1) * int(kv["output-dim"])
    return n


def base_config(args, world):
    """The `config` object both arms print (identical keys and values for the same command line)."""
    n = args.global_batch // world if args.global_batch else args.batch
    return {"workload": WORKLOADS[args.workload][1], "per_gpu_batch": n, "global_batch": n * world,
            "frames_per_example": 21, "feature_dim": 40, "math": args.math}


class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region: NVML polled from a thread (a
    timed region of a few milliseconds still gets its sample: the first is taken as the region starts;
    `nvidia-smi -lms 100` as a subprocess, the recipe's form, needs hundreds of milliseconds to produce
    its first line and is the fallback)."""
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.stop_flag, self.thread, self.nvml, self.handle = gpu_index, [], False, None, None, None
        self.smi = None

    def _handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        self.nvml = pynvml
        try:
            pr = torch.cuda.get_device_properties(self.gpu)
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            return pynvml.nvmlDeviceGetHandleByIndex(idx)

    def _sample(self):
        n, h = self.nvml, self.handle
        try:
            clk = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
            why = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            self.samples.append((clk, self.max_clock, self.power, why, time.perf_counter()))
        except Exception:
            pass

    def _poll(self):
        # NOT a tight loop: at a 1 - 2 ms period the queries themselves slowed a 2-GPU step from 0.77 to 1.58 ms,
        # and even ONE query inside a 38 ms region cost 0.1 ms per step (it stalls the queried GPU for a few
        # ms, and one slow rank stalls every peer barrier).  First sample now (the caller has warm-up steps in
        # flight), then one every 250 ms.
        while not self.stop_flag:
            self._sample()
            for _ in range(250):
                if self.stop_flag:
                    break
                time.sleep(0.001)

    def mark_region(self):
        self.region_t0 = time.perf_counter()

    def start(self):
        try:
            self.handle = self._handle()
            self.max_clock = self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
            self.power = self.nvml.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:                      # fallback: the subprocess form
                q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                     "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.smi = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                             "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                            stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self.smi = None

    def stop(self):
        if self.thread is not None:
            t_end = time.perf_counter()
            self.stop_flag = True
            self.thread.join(timeout=1.0)
            self._sample()                               # right after the timed region
            try:
                self.power = max(self.power, self.nvml.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            t0 = getattr(self, "region_t0", 0.0)
            inside = sum(1 for x in self.samples if t0 <= x[4] <= t_end)
            sm = sorted(x[0] for x in self.samples)
            reasons = set()
            for x in self.samples:
                for bit, name in self.REASONS:
                    if x[3] & bit:
                        reasons.add(name)
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None,
                    "sm_max_mhz": float(max(x[1] for x in self.samples)) if self.samples else None,
                    "power_w_max": self.power,
                    "samples": len(sm), "samples_inside_timed_region": inside, "reasons": sorted(reasons),
                    "source": "NVML: one sample under the load of the last warm-up steps, one every 250 ms inside "
                              "the timed region, one right after it (a query stalls the queried GPU for a few ms, "
                              "so none is forced into a short region)"}
        if self.smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["NVML and nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.smi.terminate()
        try:
            out = self.smi.communicate(timeout=2)[0]
        except Exception:
            self.smi.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons),
                "source": "nvidia-smi -lms 100"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measure_matmul_peaks():
    """the TF32 and FP32 dense peaks are 'to be measured by the builder' with the
    method of MEASURED_PEAKS.json (torch.matmul 8192^3, best of 10): the library GEMM is the yardstick, it is
    not on the product path."""
    import torch
    out = {}
    a = torch.randn(8192, 8192, device="cuda")
    b = torch.randn(8192, 8192, device="cuda")
    was = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("tf32_tflops", True), ("fp32_tflops", False)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.matmul(a, b)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(10 if tf32 else 3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                torch.matmul(a, b)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            out[name] = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = was
    out["how"] = "torch.matmul 8192^3 (cuBLAS), best of 10 (TF32) / 3 (FP32), CUDA events"
    return out


# ------------------------------------------------------------------ CPU arm --

def cpu_backend():
    """('reference', module) when oracle/_ref holds a CPU build of the reference, else ('port', oracle)."""
    try:
        from oracle import ref
        if ref.available():
            return "reference", ref
    except Exception:
        pass
    from oracle import oracle
    oracle.build()
    return "port", oracle


def cpu_train_frames_per_sec(cfg_text, rows, steps, warmup, threads):
    """The CPU arm: the op-for-op port of the reference's CPU path (oracle/), its GEMMs through a
    threaded OpenBLAS as in a Kaldi BLAS build, everything else the reference's own
    scalar loops.  Returns (frames/s, s/step, kind, threads, description of the back end)."""
    import numpy as np
    from oracle.cpu_nnet import CpuNnet
    kind, backend = cpu_backend()
    blas = None
    if hasattr(backend, "use_blas"):
        blas = backend.use_blas(True)
    used = 1
    if hasattr(backend, "set_num_threads"):
        used = backend.set_num_threads(threads) or 1
    net = CpuNnet(cfg_text, seed=42, backend=backend)
    rng = np.random.default_rng(1234)
    x = rng.standard_normal((rows, net.input_dim)).astype(np.float32)
    nout = [L for L in net.layers if L["kind"] == "SoftmaxComponent"][-1]["dim"]
    labels = rng.integers(0, nout, size=rows).astype(np.int64)
    for _ in range(warmup):
        net.train_step(x, labels)
    t0 = time.perf_counter()
    for _ in range(steps):
        net.train_step(x, labels)
    dt = time.perf_counter() - t0
    how = ("GEMMs: OpenBLAS (%s, %d threads); other loops: OpenMP port of the reference's CPU branches"
           % (os.path.basename(blas), used)) if blas else "GEMMs and loops: OpenMP C port (no BLAS found)"
    return rows * steps / dt, dt / steps, kind, used, how


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cfg = load_config(args.workload)
    cores = os.cpu_count() or 1
    rows = args.cpu_rows or (args.batch if args.steps + args.warmup <= 64 else 128)
    fps, sec, kind, used, how = cpu_train_frames_per_sec(cfg, rows, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": "train_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": base_config(args, world),
        "run": {"cpu_sample_rows_per_step": rows},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": used, "kind": kind,
                         "sample": "%d training steps of %d rows of the same model on the host CPU; %s"
                                   % (args.steps, rows, how)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm --

def event_time_ms(fn, iters, flush=None):
    import torch
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        times.append(a.elapsed_time(b))
    times.sort()
    return times[len(times) // 2], sum(times) / len(times)


def ncu_traffic(label_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` summary of
    THIS round's build (profiles/r02_ncu_summary.json, produced by tools/ncu_summarise.py from the capture of
    the same bench command); None when the kernel is not in it.  ncu cannot run inside a timed run."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    if not os.path.exists(p):
        return None
    try:
        rows = [r for r in json.load(open(p)) if label_substr in r.get("kernel", "")]
        if not rows:
            return None
        return sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in rows) / len(rows)
    except Exception:
        return None


def profile_step(L, run_eager_step, pk, mm, math):
    """Per-launch CUDA-event timing of ONE eager step (kcnn_profile_*): every launch with its label,
    algorithmic work, duration and roofline; grouped by (label, kernel)."""
    L.kcnn_profile_start()
    run_eager_step()
    n = L.kcnn_profile_stop()
    tensor_peak = mm["tf32_tflops"] if math == 1 else mm["fp32_tflops"]
    ridge = tensor_peak * 1e12 / (pk["hbm_gbs"] * 1e9)           # FLOP per byte where the two roofs meet
    recs, total = [], 0.0
    kb, lb = ctypes.create_string_buffer(512), ctypes.create_string_buffer(128)
    for i in range(n):
        ms, fl, by = ctypes.c_float(), ctypes.c_double(), ctypes.c_double()
        grid = (ctypes.c_uint * 3)()
        if L.kcnn_profile_get(i, kb, 512, lb, 128, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by), grid) != 0:
            continue
        name = kb.value.decode(errors="replace")
        short = name.split("(")[0].replace("void ", "")[:110]
        recs.append({"label": lb.value.decode(errors="replace"), "kernel": short, "ms": max(ms.value, 0.0),
                     "flops": fl.value, "bytes": by.value, "grid": [grid[0], grid[1], grid[2]]})
        total += max(ms.value, 0.0)
    groups = {}
    for r in recs:
        g = groups.setdefault((r["label"], r["kernel"]), {"label": r["label"], "kernel": r["kernel"], "launches": 0,
                                                          "ms": 0.0, "flops": 0.0, "bytes": 0.0, "grid": r["grid"]})
        g["launches"] += 1; g["ms"] += r["ms"]; g["flops"] += r["flops"]; g["bytes"] += r["bytes"]
    out = []
    for g in groups.values():
        g["share"] = g["ms"] / total if total > 0 else 0.0
        sec = g["ms"] * 1e-3
        if g["flops"] > 0 and (g["bytes"] <= 0 or g["flops"] / g["bytes"] >= ridge):
            ach = g["flops"] / sec / 1e12 if sec > 0 else 0.0
            g.update(bound="tensor", achieved=ach, peak=tensor_peak, unit="TFLOP/s", frac=ach / tensor_peak)
        elif g["bytes"] > 0:
            ach = g["bytes"] / sec / 1e9 if sec > 0 else 0.0
            g.update(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"])
        else:
            g.update(bound=None, achieved=None, peak=None, unit=None, frac=None)
        out.append(g)
    out.sort(key=lambda g: -g["ms"])
    return out, total, n


def kernel_rooflines(args, pk, mm, math):
    """Isolated, L2-flushed CUDA-event timings of the kernels BASELINE's metric names: a convolution of the
    model through the channels-last entry point the step uses (tensor pipe), max-pool forward / backward at
    the C4 sweep shapes and at the model's size (HBM GB/s)."""
    import torch
    from kaldi_cnn_b200 import capi
    from kaldi_cnn_b200.capi import mdim, ptr, stream
    L = capi.lib()
    N = args.batch
    out = []
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    peak = mm["tf32_tflops"] if math == 1 else mm["fp32_tflops"]
    note = "measured in this run: " + mm["how"]

    def gemm_entry(name, fn, flops):
        med, _ = event_time_ms(fn, 20, flush)
        out.append({"kernel": name, "bound": "tensor", "achieved": flops / (med * 1e-3) / 1e12, "peak": peak,
                    "unit": "TFLOP/s", "frac": flops / (med * 1e-3) / 1e12 / peak, "ms": med, "peak_source": note})

    # FC2 (4096 x 4096): the largest GEMMs of the step
    x = torch.randn(N, 4096, device="cuda")
    w = torch.randn(4096, 4096, device="cuda") * 0.01
    b = torch.zeros(4096, device="cuda")
    y = torch.empty(N, 4096, device="cuda")
    fl = 2.0 * N * 4096 * 4096
    gemm_entry("affine_fprop FC2 [%dx4096]x[4096x4096]^T" % N,
               lambda: L.cudaF_affine_fprop(stream(), math, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y)), fl)
    gemm_entry("affine_dgrad FC2", lambda: L.cudaF_affine_dgrad(stream(), math, ptr(y), mdim(y), ptr(w), mdim(w), ptr(x), mdim(x)), fl)
    if math == 1:
        pv = torch.zeros(4096, 4096, device="cuda")
        w2 = w.clone()
        med, _ = event_time_ms(lambda: L.cudaF_affine_wgrad_sgd(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(w2), mdim(w2),
                                                                ptr(pv), mdim(pv), None, 0.9, -1e-9, 1e-9), 20, flush)
        byts = 16.0 * 4096 * 4096 + 4.0 * N * (4096 + 4096)
        out.append({"kernel": "affine_wgrad+sgd FC2 (W, prev_grad read + written in the GEMM epilogue)", "bound": "hbm",
                    "achieved": byts / (med * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": byts / (med * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": med, "peak_source": pk["source"],
                    "algorithmic_bytes": byts, "tensor_tflops": fl / (med * 1e-3) / 1e12})
        # conv4 of nnet.config, channels-last in and out as in the fused step: 1x14x256 -> k1x3 -> 256 maps
        W_, C, KW, G, OW = 14, 256, 3, 256, 12
        xi = torch.randn(N, W_ * C, device="cuda")
        k = torch.randn(KW * C, G, device="cuda") * 0.01
        bb = torch.zeros(G, device="cuda")
        yo = torch.empty(N, OW * G, device="cuda")
        dxo = torch.empty(N, W_ * C, device="cuda")
        fl = 2.0 * N * OW * G * KW * C
        gemm_entry("conv_time_fprop_cl conv4 (M,N,K)=(%d,256,768)" % (N * OW),
                   lambda: L.cudaF_conv_time_fprop_cl(stream(), ptr(xi), N, W_, C, 0, KW, G, ptr(k), mdim(k), ptr(bb), ptr(yo), 1, 0, 1), fl)
        gemm_entry("conv_time_dgrad_cl conv4 (+ReLU gate)",
                   lambda: L.cudaF_conv_time_dgrad_cl(stream(), ptr(yo), N, W_, C, 0, KW, G, ptr(k), mdim(k), ptr(dxo), ptr(xi)), fl)
    for (name, H, W, C, ph, pw, pc, rows) in (("1x8x2000 pool 1x2x10", 1, 8, 2000, 1, 2, 10, 8192),
                                              ("1x16x2000 pool 1x2x1", 1, 16, 2000, 1, 2, 1, 8192),
                                              ("1x12x256 pool 1x2x1 (model)", 1, 12, 256, 1, 2, 1, N)):
        ind, outd = H * W * C, (H // ph) * (W // pw) * (C // pc)
        xi = torch.randn(rows, ind, device="cuda")
        yo = torch.empty(rows, outd, device="cuda")
        dy = torch.randn(rows, outd, device="cuda")
        dx = torch.empty(rows, ind, device="cuda")
        med, _ = event_time_ms(lambda: L.cudaF_maxpool_prop_s(stream(), ptr(xi), mdim(xi), ptr(yo), mdim(yo), H, W, ph, pw, pc, 0),
                               20, flush)
        byts = 4.0 * rows * (ind + outd)
        out.append({"kernel": "maxpool_prop " + name + " N=%d" % rows, "bound": "hbm", "achieved": byts / (med * 1e-3) / 1e9,
                    "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": byts / (med * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": med,
                    "peak_source": pk["source"]})
        med, _ = event_time_ms(lambda: L.cudaF_maxpool_backprop_s(stream(), ptr(xi), mdim(xi), ptr(yo), mdim(yo), ptr(dy), mdim(dy),
                                                                  ptr(dx), mdim(dx), H, W, ph, pw, pc, 0, 1), 20, flush)
        byts = 4.0 * rows * (2 * ind + 2 * outd)
        out.append({"kernel": "maxpool_backprop(exact, zero-fill fused) " + name + " N=%d" % rows, "bound": "hbm",
                    "achieved": byts / (med * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": byts / (med * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": med, "peak_source": pk["source"]})
    return out


def other_configs(math_name):
    """BASELINE.json configs 1, 3 and 4 in the same run (tools/bench_configs.py): C1a / C1b one layer stack
    on the GPU (TF32 and FP32) and on the CPU port, C3 large-channel convolutions pass by pass, C4 max-pool
    sweep (three batch sizes per shape; the full sweep is `python tools/bench_configs.py`)."""
    spec = importlib.util.spec_from_file_location("kcnn_bench_configs", os.path.join(ROOT, "tools", "bench_configs.py"))
    bc = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bc)
    from kaldi_cnn_b200 import components as kc
    res = {"C1": {}, "C3": [], "C4": []}
    cores = os.cpu_count() or 1
    for name, lines in bc.C1.items():
        res["C1"][name] = {"N": 256, "gpu_tf32": bc.c1_gpu(lines, 256, 1), "gpu_fp32": bc.c1_gpu(lines, 256, 0),
                           "cpu": bc.c1_cpu(lines, 256, cores)}
    kc.set_math_mode(1 if math_name == "tf32" else 0)
    for (N, H, W, C, KH, KW, G) in ((256, 1, 4, 512, 1, 3, 512), (256, 1, 14, 256, 1, 3, 256), (256, 1, 8, 2000, 1, 5, 2000)):
        res["C3"] += bc.conv_three_passes(N, H, W, C, KH, KW, G)
    for (H, W, C, ph, pw, pc) in ((1, 16, 2000, 1, 2, 1), (1, 1, 4000, 1, 1, 5), (1, 8, 2000, 1, 2, 10), (33, 9, 64, 3, 3, 2)):
        res["C4"] += bc.maxpool_sweep(H, W, C, ph, pw, pc, [64, 1024, 8192])
    return res


def run_ours(args):
    import numpy as np
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from kaldi_cnn_b200 import components as kc
    from kaldi_cnn_b200 import capi

    math = 1 if args.math == "tf32" else 0
    kc.set_math_mode(math)
    kc.set_rand_seed(42)
    cfg = load_config(args.workload)
    net = kc.Nnet.from_config(cfg, skip_splice=False)          # the Splice front end is part of the step
    if args.global_batch:                       # strong scaling: the global minibatch is fixed
        if args.global_batch % world:
            raise SystemExit("bench.py: --global-batch must be a multiple of the number of GPUs")
        args.batch = args.global_batch // world
    N, dim, nout, fpe = args.batch, net.input_dim, net.output_dim, net.frames_per_example
    averaging = world > 1 and args.dp_mode == "average"
    L = capi.lib()
    pk = peaks()
    warm = max(args.warmup, 3)

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    nbuf = 3
    feats = [torch.randn(N * fpe, dim, device="cuda", generator=gen) for _ in range(nbuf)]
    labels = [torch.randint(0, nout, (N,), device="cuda", generator=gen, dtype=torch.int32) for _ in range(nbuf)]
    stream = torch.cuda.Stream()
    dp = None
    post_step = lambda: None                      # noqa: E731
    dp_reduce = None
    with torch.cuda.stream(stream):
        kc.use_current_stream()
        if averaging:
            # Comparison row of SURVEY 8d C5: the reference's recipe trains independent jobs and averages their
            # models (nnet-am-average, egs/steps/nnet0/train_conv_dropout.sh:323-341).  Here: ordinary local
            # steps on every rank, parameters averaged in memory every --average-every steps.
            from kaldi_cnn_b200.dp import ParameterAveraging
            updatable = [c for c in range(net.num_components) if L.kcnn_component_gradient_floats(net.component(c).h) > 0]
            tensors = [net.component(c).params(k) for c in updatable for k in (0, 1)]
            post_step = ParameterAveraging(tensors, dist, world, args.average_every).after_step
        elif world > 1:
            # the library's own trainer (csrc/nnet2/nnet-dp.cc): pipelined rotation, one fused reduce + SGD +
            # broadcast kernel per layer.  auto: in-switch (NVLS) form from 4 GPUs when a multicast mapping can be
            # had, else the two-shot form over CUDA-IPC peer memory; all ranks agree through one MIN all-reduce.
            from kaldi_cnn_b200.dp import NativeDataParallel
            modes = {"auto": ["ipc"], "nvls": ["nvls"], "p2p": ["ipc"], "ipc": ["ipc"]}[args.dp_reduce]
            for mode in modes:
                try:
                    dp = NativeDataParallel(net, dist, multicast=mode == "nvls")
                except Exception as e:
                    sys.stderr.write("bench.py: data-parallel arena (%s) unavailable (%r)\n" % (mode, e))
                    dp = None
                agree = torch.tensor([1 if dp is not None else 0], device="cuda")
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
                if int(agree.item()) == 1:
                    dp_reduce = mode
                    break
                if dp is not None:
                    dp.close()
                dp = None
            if dp is None:
                raise SystemExit("bench.py: no peer-memory arena could be set up on every rank")

        it = [0]

        def step():
            k = it[0] % nbuf
            it[0] += 1
            if dp is None:
                net.train_step_graph(feats[k], labels[k])      # NnetMinibatchUpdater::TrainStep: eager, record, replay
            else:
                dp.rotate(feats[k], labels[k], N * world)      # NnetDataParallel::Rotate: eager, record, replay

        dp_launches_per_step = None
        if dp is not None:
            dp.prime(feats[nbuf - 1], labels[nbuf - 1])
            # launches of one rotation, counted by the library on the FIRST one: NnetDataParallel::Rotate runs a
            # buffer's first rotation eagerly (the recorded graph later replays exactly those kernels)
            L.kcnn_reset_launch_count()
            step()
            post_step()
            if not dp.last_rotate_replayed:
                dp_launches_per_step = int(L.kcnn_launch_count())
        for _ in range(2 * nbuf + warm):          # every buffer: one eager step, one recorded, then replays
            step()
            post_step()
        stream.synchronize()
        # launches of one step, counted on an eager one (a replayed graph launches the same kernels)
        if dp is None:
            net.set_graphs(False)
            L.kcnn_reset_launch_count()
            step()
            stream.synchronize()
            launches_per_step = int(L.kcnn_launch_count())
            net.set_graphs(True)
            for _ in range(2 * nbuf):
                step()
        else:
            launches_per_step = dp_launches_per_step
        # Clocks: the first sample is taken while a last batch of warm-up steps is running (the GPU is under
        # the load of the timed loop, but an NVML query costs the queried GPU a few ms -- see ClockSampler --
        # so it must not land in a timed region that may be only 12 ms long), further samples every 250 ms
        # inside the timed region, the last one right after it.
        sampler = ClockSampler(local)
        for _ in range(8):
            step()
            post_step()
        if rank == 0:
            sampler.start()
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler.mark_region()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
            post_step()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        replayed = net.last_step_replayed if dp is None else dp.last_rotate_replayed
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        objf = net.objf_and_reset()
        steps_run = it[0]
        if dp is not None and dp.failed(True):
            # a device-side barrier of the reduction gave up waiting for a rank: that step's update was skipped
            raise SystemExit("bench.py: rank %d: peer-memory barrier timed out; result invalid" % rank)
        # a fingerprint of the trained parameters right after the timed steps
        updatable = [c for c in range(net.num_components) if L.kcnn_component_gradient_floats(net.component(c).h) > 0]
        param_checksum = 0.0
        for c in updatable:
            param_checksum += float(net.component(c).params(0).double().abs().sum().item())

        # ---- end to end through the C-ABI with HOST buffers (pinned), copies in the timed region
        e2e = None
        hx = torch.randn(N * fpe, dim).pin_memory()
        hl = torch.randint(0, nout, (N,), dtype=torch.int32).pin_memory()
        hx_np, hl_np = hx.numpy(), hl.numpy()
        h2d = N * fpe * dim * 4 + N * 4
        if args.no_e2e or averaging:
            pass                                 # launch-list captures under ncu / the averaging comparison row
        elif world == 1:
            # (a) synchronous call: copy in, step, objective back, host blocked until the step is done
            for _ in range(3):
                net.train_minibatch_host(hx_np, hl_np)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                net.train_minibatch_host(hx_np, hl_np)
            torch.cuda.synchronize()
            dt_sync = time.perf_counter() - t0
            # (b) pipelined call (the headline): every step still copies ITS inputs host -> device and
            # its objective device -> host inside the timed region, but batch k+1 is staged and copied
            # while batch k computes; the region ends when the last objective has been read.
            for _ in range(4):
                net.train_minibatch_host_async(hx_np, hl_np)
            net.objf_and_reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                net.train_minibatch_host_async(hx_np, hl_np)
            t_calls = time.perf_counter() - t0
            net.objf_and_reset()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e = {"value": N * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": 8, "ms_per_step": dt / args.steps * 1e3,
                   "host_ms_in_calls_per_step": t_calls / args.steps * 1e3,
                   "graph_replayed": bool(net.last_step_replayed),
                   "api": "kcnn_nnet_train_minibatch_host_async (include/kcnn_capi.h): staged through pinned "
                          "buffers, copy stream, library-recorded CUDA graph per slot",
                   "sync_api": {"value": N * args.steps / dt_sync, "ms_per_step": dt_sync / args.steps * 1e3,
                                "api": "kcnn_nnet_train_minibatch_host (blocking, objective returned per call)"}}
        else:
            # the same plugin call on every rank: host buffers in, objective out, nothing but the C ABI in between
            dp.finish(N * world)
            for _ in range(2 * 3 + 2):
                dp.train_minibatch_host_async(hx_np, hl_np, N * world)
            net.objf_and_reset()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                dp.train_minibatch_host_async(hx_np, hl_np, N * world)
            net.objf_and_reset()
            torch.cuda.synchronize()
            dtt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dtt, op=dist.ReduceOp.MAX)
            e2e = {"value": N * world * args.steps / float(dtt.item()), "unit": "frames/s",
                   "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "ms_per_step": float(dtt.item()) / args.steps * 1e3,
                   "api": "kcnn_nnet_dp_train_minibatch_host_async (include/kcnn_capi.h) on every rank: pinned staging, "
                          "copy stream, pipelined rotation recorded as a CUDA graph per slot; bytes are per rank"}
            if dp.failed(True):
                raise SystemExit("bench.py: rank %d: peer-memory barrier timed out; result invalid" % rank)
        if dp is not None:
            dp.finish(N * world)
            dp.close()
            dp = None

        # ---- P x 32 rows against 1 x 32P rows through the same trainer
        rank_parity = None
        if world > 1 and not averaging and not args.no_parity:
            spec = importlib.util.spec_from_file_location("kcnn_dp_check", os.path.join(ROOT, "tools", "dp_native_check.py"))
            chk = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(chk)
            worst, identical, failed, _ = chk.rank_parity(dist, 32, 4, math, dp_reduce == "nvls")
            kc.set_math_mode(math)
            tol = 1e-3 if math == 1 else 1e-5
            rank_parity = {"rows": "%d x 32 vs 1 x %d" % (world, 32 * world), "steps": 4,
                           "max_rel_diff_params_and_momentum": worst, "tolerance": tol * 4,
                           "norms": "weights / biases max-norm relative, momentum Frobenius relative",
                           "ok": bool(worst is not None and worst <= tol * 4 and identical and not failed),
                           "ranks_bit_identical": identical, "model": "C2 + intermap pooling, dropout lines removed "
                           "(a rank's dropout mask is indexed by its local row)"}

        step_kernels, step_ms_eager, kernels, mm, cfgs = [], None, [], None, None
        if rank == 0 and not args.no_kernels:
            mm = measure_matmul_peaks()
            if world == 1:
                net.set_graphs(False)
                step_kernels, step_ms_eager, _ = profile_step(L, lambda: (step(), stream.synchronize()), pk, mm, math)
                net.set_graphs(True)
            kernels = kernel_rooflines(args, pk, mm, math)
            if world == 1 and not args.no_configs:
                try:
                    cfgs = other_configs(args.math)
                except Exception as e:                   # never lose the headline to a side measurement
                    cfgs = {"error": repr(e)}
                kc.set_math_mode(math)

    if rank != 0:
        if dist is not None:
            finish_distributed(dist)
        return

    frames = N * world * args.steps
    value = frames / (ms * 1e-3)
    flops_frame = model_flops_per_frame(cfg)
    step_tflops = flops_frame * N / (ms / args.steps * 1e-3) / 1e12
    cpu = None
    if not args.no_cpu:
        try:
            cpu_rows = args.cpu_rows or args.batch
            fps, sec, kind, used, how = cpu_train_frames_per_sec(cfg, cpu_rows, 2, 1, os.cpu_count() or 1)
            cpu = {"value": fps, "unit": "frames/s", "cores": used, "kind": kind,
                   "sample": "2 training steps of %d rows of the same model (after 1 warm-up) on the host CPU; %s"
                             % (cpu_rows, how)}
        except Exception as e:
            cpu = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    # the roofline kernel is the one with the largest MEASURED share of the step
    dominant = None
    for g in step_kernels:
        if g["bound"] is not None:
            dominant = {"kernel": "%s [%s]" % (g["label"], g["kernel"]), "bound": g["bound"], "achieved": g["achieved"],
                        "peak": g["peak"], "unit": g["unit"], "frac": g["frac"], "traffic": ncu_traffic(g["label"]),
                        "share_of_step": g["share"], "launch_ms": g["ms"] / max(g["launches"], 1),
                        "algorithmic": {"flops": g["flops"], "bytes": g["bytes"]},
                        "how": "largest share among the %d launch groups of one eager step timed per launch with CUDA "
                               "events on the launching streams (kcnn_profile_*); peak: %s" % (
                                   len(step_kernels), "HBM copy bandwidth of MEASURED_PEAKS.json" if g["bound"] == "hbm"
                                   else "TF32 dense matmul measured in this run (%s)" % mm["how"])}
            break
    line = {
        "metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "tf32" if math == 1 else "f32", "data": "synthetic",
        "config": base_config(args, world),
        "run": {"parallelism": "dp%d" % world if world > 1 else "single",
                **({"dp_schedule": "BASELINE (comparison row, not synchronous SGD): independent local steps, "
                                   "parameters averaged every %d steps (nnet-am-average emulation)" % args.average_every,
                    "dp_reduce": "NCCL all-reduce of the parameters"} if averaging else
                   {"dp_schedule": "NnetDataParallel::Rotate: backward(t) + per-layer reduce/SGD/broadcast kernel + "
                                   "forward(t+1), one CUDA graph per input slot",
                    "dp_reduce": ("kcnn_p2p_reduce_sgd_f32, in-switch (multimem) form" if dp_reduce == "nvls" else
                                  "kcnn_p2p_reduce_sgd_f32, two-shot over CUDA-IPC peer memory")} if world > 1 else {}),
                "params": param_count(cfg), "train_mflop_per_frame": flops_frame / 1e6,
                "l2": "working set (weights + momentum = %.0f MB, %d input buffers) exceeds the 126 MB L2"
                      % (param_count(cfg) * 8 / 1e6, nbuf),
                "cuda_graph": bool(replayed), "fused_plan": bool(net.fused_active),
                "splice": "front end included: %d frames x %d features per example, read in place" % (fpe, dim)},
        "step_tflops": step_tflops,
        "objf_per_frame_last": objf / max(N * steps_run, 1),
        "param_checksum": param_checksum,
        "clocks": clocks, "e2e": e2e,
        "gpu_launches": (launches_per_step or 0) * args.steps, "gpu_launches_per_step": launches_per_step,
        "roofline": dominant, "step_kernels": step_kernels, "step_ms_sum_of_launches": step_ms_eager,
        "measured_matmul_peaks": mm, "kernels": kernels, "configs": cfgs, "rank_parity": rank_parity,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        finish_distributed(dist)


def finish_distributed(dist):
    """Leave a multi-rank run without tearing NCCL down: destroying the process group while a
    captured graph still references its communicator can hang at exit (seen with 2 ranks)."""
    import torch
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2-intermap", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=512, help="minibatch rows per GPU")
    ap.add_argument("--math", default=os.environ.get("KCNN_BENCH_MATH", "tf32"), choices=["tf32", "fp32"])
    ap.add_argument("--cpu-rows", type=int, default=0,
                    help="rows per step of the CPU legs; 0 = the workload's own minibatch (--batch) when the run is "
                         "short enough (<= 64 steps incl. warm-up), else 128 (the reference's CPU minibatch, "
                         "egs/local/nnet0/run_nnet.sh:25-27)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-kernels", action="store_true", help="skip the per-launch profile and the isolated kernel rooflines")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 rows")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the P x 32 vs 1 x 32P parity check")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (ncu launch-list captures only)")
    ap.add_argument("--dp-reduce", default=os.environ.get("KCNN_BENCH_DP_REDUCE", "auto"),
                    choices=["auto", "ipc", "p2p", "nvls"],
                    help="arena of the data-parallel trainer: CUDA-IPC peer memory with the two-shot kernel (ipc; p2p is an "
                         "alias), or a multicast mapping with the in-switch kernel (nvls); auto: nvls from 4 GPUs when available")
    ap.add_argument("--dp-mode", default="sync", choices=["sync", "average"],
                    help="sync: fused reduce + SGD every step (the product). average: the comparison row of SURVEY 8d "
                         "C5 -- local steps + parameter averaging every --average-every steps (nnet-am-average style)")
    ap.add_argument("--average-every", type=int, default=8)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fix the GLOBAL minibatch (rows per GPU = global / GPUs); 0 = weak scaling, "
                         "--batch rows per GPU")
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: one process per GPU through torchrun, as the driver does
        import socket
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                  str(args.gpus), "--master-addr", "127.0.0.1", "--master-port", str(port),
                                  os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
