"""Data-parallel parity of the library's own trainer (csrc/nnet2/nnet-dp.cc, kcnn_nnet_dp_*) on GPUs:
SURVEY 8e, "P ranks x N/P rows reproduce the 1-rank N-row step up to floating-point summation order"
-- at 8 ranks that is 8 x 32 rows against 1 x 256.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/dp_native_check.py

Every rank builds the same network (the C2 model with intermap pooling, dropout lines removed: a rank's
dropout mask is indexed by its LOCAL row, so the masks of a sharded batch and of the whole batch differ
by construction) and trains STEPS global minibatches data-parallel through the host-buffer entry point
kcnn_nnet_dp_train_minibatch_host_async (pipelined rotation, per-layer fused reduce + SGD + broadcast
kernel, CUDA graph from the third call on); rank 0 also trains a second copy on the whole batches the
ordinary single-GPU way.  Parameters AND momentum must agree within the tolerance of the math mode, and
all ranks must hold bit-identical parameters.  Used by tests/test_gpu_dp_native.py and by bench.py
(rank_parity)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kaldi_cnn_b200 import components as kc  # noqa: E402
from kaldi_cnn_b200.dp import NativeDataParallel, shard_rows  # noqa: E402


def config_without_dropout():
    text = open(os.path.join(ROOT, "kaldi-cnn_b200", "configs", "nnet_c2_intermap.config")).read()
    return "\n".join(l for l in text.splitlines() if not l.startswith("DropoutComponent")) + "\n"


def params(net):
    out = []
    for i in range(net.num_components):
        c = net.component(i)
        if c.type in ("ConvolutionComponent", "FullyConnectedComponent"):
            out += [c.params(k).detach().clone() for k in range(3)]
    return out


def rank_parity(dist, rows_per_rank=32, steps=4, math=1, multicast=False, verbose=False):
    """Returns (largest relative difference DP vs single GPU over W / bias (max-norm) and momentum (Frobenius) [rank 0, else None],
    all ranks bit-identical?, barrier timed out?, rotations replayed from a graph?)."""
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = config_without_dropout()
    kc.set_math_mode(math)
    N = rows_per_rank * world
    rng = np.random.default_rng(3)
    kc.set_rand_seed(7)
    net = kc.Nnet.from_config(cfg, skip_splice=False)
    fpe = net.frames_per_example
    ref = None
    if rank == 0:
        kc.set_rand_seed(7)
        ref = kc.Nnet.from_config(cfg, skip_splice=False)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        kc.use_current_stream()
        dp = NativeDataParallel(net, dist, multicast=multicast)
        xs = [rng.standard_normal((N * fpe, net.input_dim)).astype(np.float32) for _ in range(steps)]
        ls = [rng.integers(0, net.output_dim, N).astype(np.int32) for _ in range(steps)]
        b, e = shard_rows(N, rank, world)
        replayed = []
        for k in range(steps):
            dp.train_minibatch_host_async(np.ascontiguousarray(xs[k][b * fpe:e * fpe]), np.ascontiguousarray(ls[k][b:e]), N)
            replayed.append(dp.last_rotate_replayed)
        dp.finish(N)
        failed = dp.failed(True)
        dp.gather_momentum()
        stream.synchronize()
        worst = None
        if ref is not None:
            for k in range(steps):
                ref.train_step(torch.from_numpy(xs[k]).cuda(), torch.from_numpy(ls[k]).cuda())
            stream.synchronize()
            # weights / biases: max-norm relative; momentum matrices (index 2 of every layer's triple) are
            # sums with heavy cancellation whose largest entries can be ~1e-9: Frobenius-relative
            worst, detail = 0.0, []
            for idx, (pa, pb) in enumerate(zip(params(net), params(ref))):
                if idx % 3 == 2:
                    d = float((pa.double() - pb.double()).norm()) / (float(pb.double().norm()) + 1e-30)
                else:
                    d = float((pa - pb).abs().max()) / (float(pb.abs().max()) + 1e-30)
                if not torch.isfinite(pa).all():
                    d = float("inf")
                detail.append("%.1e" % d)
                worst = max(worst, d)
            if verbose:
                print("   per tensor (W, b, prev per layer):", " ".join(detail), flush=True)
        chk = torch.stack([p.double().sum() for p in params(net)]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = lo.item() == hi.item()
        dp.close()
    kc.use_current_stream()
    return worst, identical, failed, replayed


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    multicast = os.environ.get("KCNN_DP_CHECK_REDUCE", "ipc") == "nvls"
    ok = True
    for math, tol in ((1, 1e-3), (0, 1e-5)):
        worst, identical, failed, replayed = rank_parity(dist, 32, 4, math, multicast,
                                                         bool(os.environ.get("KCNN_DP_CHECK_VERBOSE")))
        if rank == 0:
            print("dp_native_check world=%d (%d x 32 rows vs 1 x %d) reduce=%s math=%s: max relative difference of "
                  "parameters / momentum %.3g (tolerance %.0e x 4 steps); ranks identical: %s; barrier timeout: %s; "
                  "graph replays: %s" % (world, world, 32 * world, "nvls" if multicast else "ipc two-shot",
                                         "tf32" if math else "fp32", worst, tol, identical, failed, replayed), flush=True)
            ok = ok and worst <= tol * 4
        ok = ok and identical and not failed
        dist.barrier()
    kc.set_math_mode(0)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
