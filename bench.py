#!/usr/bin/env python
"""bench.py -- CNN training throughput (frames/sec) of the kaldi-cnn hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one training minibatch through the hot path: Propagate through every
component of the model, cross-entropy objective + derivative, Backprop in reverse with the
parameter update inside Backprop (momentum / weight-decay SGD), exactly the loop nnet2's
NnetUpdater runs.  One frame = one minibatch row.

Workload (config.workload): BASELINE.json configs[1] -- the reference's time-axis deep CNN
(egs/exp/nnet/nnet.config: 40 mel x 21 frames, 6 conv + pool + 3 FC + softmax 3454) with an
intermap max-pool after conv1 (pool-channel-dim=2, the egs/local/nnet0/run_conv.sh shape),
per-GPU minibatch 512 (the recipes' GPU minibatch, egs/local/nnet0/run_nnet.sh:19-20), synthetic
N(0,1) filterbank windows, random-init weights.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel of the step, timed live with CUDA events
  cpu_baseline  the reference's CPU path (oracle/_ref when built, else the oracle port)
                timed on this box's host cores on a bounded sample
  kernels       per-kernel rooflines named by BASELINE's metric (conv tensor pipe, maxpool HBM GB/s)
"""
import argparse
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
                    "MaxpoolComponent pool-channel-dim=2 after conv1 (40x21x1 input, 6 conv, 2 maxpool, 3 FC, "
                    "softmax 3454)"),
    "c2": ("nnet_c2.config",
           "C2: egs/exp/nnet/nnet.config verbatim (40x21x1 input, 6 conv, time max-pool, 3 FC, softmax 3454)"),
}


def load_config(name):
    path = os.path.join(ROOT, "kaldi-cnn_b200", "configs", WORKLOADS[name][0])
    return open(path).read()


def model_flops_per_frame(cfg_text):
    """Algorithmic training FLOPs per frame: 3 x forward GEMM FLOPs (fprop, dgrad, wgrad) of the
    conv and FC layers."""
    from oracle.cpu_nnet import parse_config
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
    from oracle.cpu_nnet import parse_config
    n = 0
    for kind, kv in parse_config(cfg_text):
        if kind == "ConvolutionComponent":
            n += (int(kv["kernel-height"]) * int(kv["kernel-width"]) * int(kv["in-channel"]) + 1) * int(kv["group"])
        elif kind == "FullyConnectedComponent":
            n += (int(kv["input-dim"]) + 1) * int(kv["output-dim"])
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------ CPU arm --

def cpu_backend():
    """('reference', module) when oracle/_ref is built, else ('port', oracle)."""
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
    if rank != 0:
        return
    cfg = load_config(args.workload)
    cores = os.cpu_count() or 1
    rows = args.cpu_rows or (args.batch if args.steps + args.warmup <= 64 else 128)
    fps, sec, kind, used, how = cpu_train_frames_per_sec(cfg, rows, args.steps, args.warmup, cores)
    line = {
        "impl": "reference", "metric": "train_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][1], "per_gpu_batch": args.batch,
                   "cpu_sample_rows_per_step": rows},
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


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel whose name contains
    `kernel_substr`, from the committed `ncu --set full` summary (profiles/); None when absent."""
    p = os.path.join(ROOT, "profiles", "r01_ncu_summary.json")
    if not os.path.exists(p):
        return None
    try:
        rows = [r for r in json.load(open(p)) if kernel_substr in r.get("kernel", "")]
        if not rows:
            return None
        return sum(r["dram_read_bytes"] + r["dram_write_bytes"] for r in rows) / len(rows)
    except Exception:
        return None


def kernel_rooflines(args, pk, math):
    """Live CUDA-event timings of the kernels BASELINE's metric names, at the model's shapes:
    the dominant GEMM of the step, a conv fprop, and max-pool forward / backward."""
    import torch
    from kaldi_cnn_b200 import capi
    from kaldi_cnn_b200.capi import mdim, ptr, stream
    L = capi.lib()
    N = args.batch
    out = []
    flush = torch.empty(160 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    tf32_peak = pk["bf16_tflops"] / 2.0
    simt_peak = 148 * 128 * 2 * 1.965e9 / 1e12
    peak = tf32_peak if math == 1 else simt_peak
    peak_note = ("measured bf16 dense / 2 (TF32 runs at half the bf16 rate)" if math == 1
                 else "nominal FP32 FMA: 148 SM x 128 lanes x 2 x 1.965 GHz")

    def gemm_entry(name, fn, flops):
        med, _ = event_time_ms(fn, 20, flush)
        out.append({"kernel": name, "bound": "tensor", "achieved": flops / (med * 1e-3) / 1e12, "peak": peak,
                    "unit": "TFLOP/s", "frac": flops / (med * 1e-3) / 1e12 / peak, "ms": med,
                    "peak_source": pk["source"] + ": " + peak_note, "traffic": None})

    # FC2 (4096 x 4096): the largest GEMMs of the step
    x = torch.randn(N, 4096, device="cuda")
    w = torch.randn(4096, 4096, device="cuda") * 0.01
    b = torch.zeros(4096, device="cuda")
    y = torch.empty(N, 4096, device="cuda")
    g = torch.empty(4096, 4096, device="cuda")
    bg = torch.empty(4096, device="cuda")
    fl = 2.0 * N * 4096 * 4096
    gemm_entry("affine_fprop FC2 [%dx4096]x[4096x4096]^T" % N,
               lambda: L.cudaF_affine_fprop(stream(), math, ptr(x), mdim(x), ptr(w), mdim(w), ptr(b), ptr(y), mdim(y)), fl)
    gemm_entry("affine_dgrad FC2", lambda: L.cudaF_affine_dgrad(stream(), math, ptr(y), mdim(y), ptr(w), mdim(w), ptr(x), mdim(x)), fl)
    gemm_entry("affine_wgrad FC2", lambda: L.cudaF_affine_wgrad(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(g), mdim(g), ptr(bg)), fl)
    # The largest single kernel of the step: FC2's weight gradient with the momentum / weight-decay
    # SGD step applied in its epilogue.  It moves 16 B per weight (read + write W and prev_grad) on top
    # of the GEMM operands, so its binding roofline is HBM, not the tensor pipe.
    if math == 1:
        pv = torch.zeros(4096, 4096, device="cuda")
        w2 = w.clone()
        med, _ = event_time_ms(lambda: L.cudaF_affine_wgrad_sgd(stream(), math, ptr(x), mdim(x), ptr(y), mdim(y), ptr(w2), mdim(w2),
                                                                ptr(pv), mdim(pv), ptr(b), 0.9, -1e-9, 1e-9), 20, flush)
        byts = 16.0 * 4096 * 4096 + 4.0 * N * (4096 + 4096)
        out.append({"kernel": "affine_wgrad+sgd FC2 (tma_gemm_persistent_kernel<DenseProb<MN,MN,EPI_SGD>>)", "bound": "hbm",
                    "achieved": byts / (med * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": byts / (med * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": med, "peak_source": pk["source"],
                    "algorithmic_bytes": byts, "tensor_tflops": fl / (med * 1e-3) / 1e12,
                    "traffic": ncu_traffic("DenseProb<1, 1, 2>")})
    # conv4 of nnet.config: 1x14x256 -> k1x3 -> 256 maps
    H, W, C, KH, KW, G = 1, 14, 256, 1, 3, 256
    xi = torch.randn(N, H * W * C, device="cuda")
    k = torch.randn(KH * KW * C, G, device="cuda") * 0.01
    bb = torch.zeros(G, device="cuda")
    yo = torch.empty(N, 12 * G, device="cuda")
    fl = 2.0 * N * 12 * G * KH * KW * C
    gemm_entry("conv2d_fprop conv4 (M,N,K)=(%d,256,768)" % (N * 12),
               lambda: L.cudaF_conv2d_fprop(stream(), math, ptr(xi), mdim(xi), ptr(k), mdim(k), ptr(bb), ptr(yo), mdim(yo),
                                            H, W, C, 0, 0, KH, KW, G, 1), fl)
    # max pooling at the C4 sweep shape 1x8x2000 pool 1x2x10 and at the model's time pool
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
                    "peak_source": pk["source"], "traffic": None})
        med, _ = event_time_ms(lambda: L.cudaF_maxpool_backprop_s(stream(), ptr(xi), mdim(xi), ptr(yo), mdim(yo), ptr(dy), mdim(dy),
                                                                  ptr(dx), mdim(dx), H, W, ph, pw, pc, 0, 1), 20, flush)
        byts = 4.0 * rows * (2 * ind + 2 * outd)
        out.append({"kernel": "maxpool_backprop(exact, zero-fill fused) " + name + " N=%d" % rows, "bound": "hbm",
                    "achieved": byts / (med * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": byts / (med * 1e-3) / 1e9 / pk["hbm_gbs"], "ms": med, "peak_source": pk["source"],
                    "traffic": None})
    return out


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
    net = kc.Nnet.from_config(cfg, skip_splice=True)
    if args.global_batch:                       # strong scaling: the global minibatch is fixed
        if args.global_batch % world:
            raise SystemExit("bench.py: --global-batch must be a multiple of the number of GPUs")
        args.batch = args.global_batch // world
    N, dim, nout = args.batch, net.input_dim, net.output_dim
    averaging = world > 1 and args.dp_mode == "average"
    L = capi.lib()
    pk = peaks()

    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    feats = torch.randn(N, dim, device="cuda", generator=gen)
    labels = torch.randint(0, nout, (N,), device="cuda", generator=gen, dtype=torch.int32)
    stream = torch.cuda.Stream()
    arena = None
    peer = None
    if world > 1 and not averaging:
        with torch.cuda.stream(stream):
            # auto: the in-switch (multimem) kernel from 4 GPUs up -- measured at 8 GPUs: 1.06 ms / step vs
            # 1.15 two-shot vs 1.23 NCCL -- and the two-shot kernel at 2 (1.00 vs 1.06 vs 1.05)
            modes = {"auto": ["nvls", "p2p"] if world > 2 else ["p2p"], "nvls": ["nvls"], "p2p": ["p2p"],
                     "nccl": []}[args.dp_reduce]
            for mode in modes:
                # gradient arena in NVLink peer memory, reduced by the library's own kernel
                try:
                    from kaldi_cnn_b200.dp import PeerMemoryAllReduce
                    peer = PeerMemoryAllReduce(L, dist, net.gradient_floats(), multicast=mode == "nvls")
                except Exception as e:
                    sys.stderr.write("bench.py: peer-memory all-reduce (%s) unavailable (%r)\n" % (mode, e))
                    peer = None
                # all ranks take the same path: one rank without peer memory sends everyone on
                agree = torch.tensor([1 if peer is not None else 0], device="cuda")
                dist.all_reduce(agree, op=dist.ReduceOp.MIN)
                if int(agree.item()) == 1:
                    break
                peer = None
            if peer is not None:
                arena = net.enable_data_parallel(peer.arena)
            if peer is None:
                arena = net.enable_data_parallel()
    ncomp = net.num_components
    updatable = [c for c in range(ncomp) if L.kcnn_component_gradient_floats(net.component(c).h) > 0]

    dp_step = None
    post_step = lambda: None                      # noqa: E731
    if averaging:
        # Comparison row of SURVEY 8d C5: the reference's recipe trains independent jobs and averages their
        # models (nnet-am-average, egs/steps/nnet0/train_conv_dropout.sh:323-341).  Here: ordinary local
        # steps on every rank, parameters averaged in memory every --average-every steps.
        from kaldi_cnn_b200.dp import ParameterAveraging
        tensors = [net.component(c).params(k) for c in updatable for k in (0, 1)]
        post_step = ParameterAveraging(tensors, dist, world, args.average_every).after_step
    elif world > 1:
        # Data parallel: software-pipelined step (dp.py) -- backward of batch t with the per-layer
        # all-reduces, then forward of batch t+1, the FC stack's update sitting between the
        # convolution forward and the FC forward so its all-reduce hides under both.
        from kaldi_cnn_b200.dp import PipelinedDataParallelStep, late_components
        small_group = dist.new_group(ranks=list(range(world)))
        dp_step = PipelinedDataParallelStep(net, arena, updatable, dist, world, late_components(net, updatable),
                                            small_group, skip_reduce=args.dp_skip_reduce, peer=peer)

    def step():
        if dp_step is None:
            net.forward(feats)
            net.objf_and_deriv(labels)
            net.backward()
        else:
            dp_step.rotate(feats, labels, N * world)

    with torch.cuda.stream(stream):
        kc.use_current_stream()
        if dp_step is not None:
            dp_step.prime(feats, labels)
        for _ in range(max(args.warmup, 3)):
            step()
        stream.synchronize()
        L.kcnn_reset_launch_count()
        step()
        stream.synchronize()
        launches_per_step = int(L.kcnn_launch_count())
        graph = None
        if args.graph and (world == 1 or os.environ.get("KCNN_BENCH_DP_GRAPH", "1") != "0"):
            # world > 1: NCCL all-reduces are captured into the same graph (one replay per step
            # on every rank); falls back to eager launches if the capture is refused.
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream):
                    step()
            except Exception as e:           # capture is an optimisation, never a requirement
                sys.stderr.write("bench.py: CUDA graph capture unavailable (%s); running eagerly\n" % e)
                graph = None
                torch.cuda.synchronize()
        run = (lambda: graph.replay()) if graph is not None else step
        for _ in range(3):
            run()
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            run()
            post_step()
        e1.record(stream)
        e1.synchronize()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        objf = net.objf_and_reset()
        # a fingerprint of the trained parameters right after the timed steps: two runs that did the
        # same arithmetic (e.g. --dp-reduce p2p vs nccl at 2 GPUs) print the same number
        param_checksum = 0.0
        for c in updatable:
            param_checksum += float(net.component(c).params(0).double().abs().sum().item())

        if peer is not None and peer.failed():
            # a device-side barrier of the peer-memory reduction gave up waiting for a rank: the
            # gradients of that step were incomplete -- no number is better than a wrong one
            raise SystemExit("bench.py: rank %d: peer-memory all-reduce barrier timed out; result invalid" % rank)

        # ---- end to end through the C-ABI with HOST buffers (pinned), copies in the timed region
        e2e = None
        if args.no_e2e:
            pass                                 # launch-list captures under ncu only
        elif world == 1:
            hx = torch.randn(N, dim).pin_memory()
            hl = torch.randint(0, nout, (N,), dtype=torch.int32).pin_memory()
            hx_np, hl_np = hx.numpy(), hl.numpy()
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
            net.objf_and_reset()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e = {"value": N * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": N * dim * 4 + N * 4,
                   "d2h_bytes_per_step": 8, "ms_per_step": dt / args.steps * 1e3,
                   "api": "kcnn_nnet_train_minibatch_host_async (include/kcnn_capi.h): staged through pinned "
                          "buffers, copy stream, library-recorded CUDA graph per slot",
                   "sync_api": {"value": N * args.steps / dt_sync, "ms_per_step": dt_sync / args.steps * 1e3,
                                "api": "kcnn_nnet_train_minibatch_host (blocking, objective returned per call)"}}
        else:
            # per-rank host->device copy + step + objective read, max over ranks
            hx = torch.randn(N, dim).pin_memory()
            hl = torch.randint(0, nout, (N,), dtype=torch.int32).pin_memory()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                feats.copy_(hx, non_blocking=True)
                labels.copy_(hl, non_blocking=True)
                run()                            # the recorded rotation reads feats / labels in place
                post_step()
                net.objf_and_reset()
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e = {"value": N * world * args.steps / float(dt.item()), "unit": "frames/s",
                   "h2d_bytes_per_step": N * dim * 4 + N * 4, "d2h_bytes_per_step": 8,
                   "api": "kcnn_nnet_forward/backward + %s (one recorded CUDA graph per rotation), pinned host "
                          "buffers per rank, objective read back every step" % ("NCCL all-reduce" if peer is None else
                                               "kcnn_p2p_allreduce_multicast_f32" if peer.multicast_base else
                                               "kcnn_p2p_allreduce_f32")}

        kernels = kernel_rooflines(args, pk, math) if (rank == 0 and not args.no_kernels) else []

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
    dominant = None
    for kr in kernels:
        if kr["kernel"].startswith("affine_wgrad+sgd"):
            dominant = dict(kr)
    if dominant is None:
        for kr in kernels:
            if kr["kernel"].startswith("affine_wgrad"):
                dominant = dict(kr)
    line = {
        "metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "tf32" if math == 1 else "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][1], "per_gpu_batch": N, "global_batch": N * world,
                   "parallelism": "dp%d" % world if world > 1 else "single",
                   **({"dp_schedule": "BASELINE (comparison row, not synchronous SGD): independent local steps, "
                                      "parameters averaged every %d steps (nnet-am-average emulation)"
                                      % args.average_every, "dp_reduce": "NCCL all-reduce of the parameters"}
                      if averaging else
                      {"dp_schedule": "pipelined: backward(t) + all-reduce + update + forward(t+1) per step",
                       "dp_reduce": ("NCCL all-reduce" if peer is None else
                                     "kcnn_p2p_allreduce_multicast_f32 (NVSwitch in-switch reduction, multimem)"
                                     if peer.multicast_base else
                                     "kcnn_p2p_allreduce_f32 (NVLink peer memory, two-shot)")} if world > 1 else {}),
                   "params": param_count(cfg), "train_mflop_per_frame": flops_frame / 1e6,
                   "l2": "working set (weights + momentum + gradients = %.0f MB) exceeds the 126 MB L2"
                         % (param_count(cfg) * 12 / 1e6),
                   "cuda_graph": graph is not None,
                   **({"INVALID": "all-reduces skipped (--dp-skip-reduce diagnosis run)"} if args.dp_skip_reduce else {}),
                   "math": "KCNN_MATH_TF32_TC" if math == 1 else "KCNN_MATH_FP32_SIMT"},
        "step_tflops": step_tflops,
        "objf_per_frame_last": objf / max(N * (args.steps + 3 + 1), 1),
        "param_checksum": param_checksum,
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step,
        "roofline": dominant, "kernels": kernels, "cpu_baseline": cpu,
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
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (ncu launch-list captures only)")
    ap.add_argument("--dp-reduce", default=os.environ.get("KCNN_BENCH_DP_REDUCE", "auto"),
                    choices=["auto", "p2p", "nvls", "nccl"],
                    help="gradient all-reduce of the data-parallel step: the library's NVLink peer-memory kernel "
                         "(kcnn_p2p_allreduce_f32: two-shot), its in-switch variant (nvls: multimem) or NCCL")
    ap.add_argument("--dp-mode", default="sync", choices=["sync", "average"],
                    help="sync: gradient all-reduce every step (the product). average: the comparison row of SURVEY 8d "
                         "C5 -- local steps + parameter averaging every --average-every steps (nnet-am-average style)")
    ap.add_argument("--average-every", type=int, default=8)
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fix the GLOBAL minibatch (rows per GPU = global / GPUs); 0 = weak scaling, "
                         "--batch rows per GPU")
    ap.add_argument("--dp-skip-reduce", action="store_true",
                    help="diagnosis only: run the data-parallel step without its all-reduces (invalid as a result)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
