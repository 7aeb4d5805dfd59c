"""Data-parallel parity on GPUs ("P ranks x N/P rows reproduce the 1-rank N-row step up to
floating-point summation order").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 tools/dp_check.py

Every rank builds the same small network and trains STEPS minibatches data-parallel (its shard of
each global batch, deferred update, gradient all-reduce through the peer-memory kernel or NCCL,
apply with lr / N_global); rank 0 also trains a second copy on the whole batch the ordinary way.
The two sets of parameters must agree within the tolerance of the math mode.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kaldi_cnn_b200 import capi, components as kc  # noqa: E402
from kaldi_cnn_b200.dp import (DataParallelStep, PeerMemoryAllReduce, PipelinedDataParallelStep,  # noqa: E402
                               late_components, shard_rows)

# the C2 model without dropout noise (a rank's mask is indexed by its LOCAL row, so masks of a
# sharded batch and of the whole batch differ by construction)
CFG = open(os.path.join(ROOT, "kaldi-cnn_b200", "configs", "nnet_c2_intermap.config")).read().replace(
    "dropout-proportion=0.5", "dropout-proportion=0.0")
STEPS = 4


def params(net):
    out = []
    for i in range(net.num_components):
        c = net.component(i)
        if c.type in ("ConvolutionComponent", "FullyConnectedComponent"):
            out += [c.params(k).detach().clone() for k in range(3)]
    return out


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = capi.lib()
    mode = os.environ.get("KCNN_DP_CHECK_REDUCE", "p2p")
    ok = True
    for math, tol in ((1, 1e-3), (0, 1e-5)):
        kc.set_math_mode(math)
        N = 64 * world
        rng = np.random.default_rng(3)
        kc.set_rand_seed(7)
        net = kc.Nnet.from_config(CFG, skip_splice=True)
        peer = (PeerMemoryAllReduce(L, dist, net.gradient_floats(), multicast=mode == "nvls")
                if mode in ("p2p", "nvls") else None)
        arena = net.enable_data_parallel(peer.arena if peer else None)
        updatable = [c for c in range(net.num_components) if L.kcnn_component_gradient_floats(net.component(c).h) > 0]
        plain = DataParallelStep(net, arena, updatable, dist, world)
        if peer is not None:                                       # same reductions through the library's kernel
            class _PeerDist:
                def all_reduce(self, t, async_op=True, group=None):
                    off = (t.data_ptr() - arena.data_ptr()) // 4
                    return peer.all_reduce(off, t.numel())
            plain.dist = _PeerDist()
        ref = None
        if rank == 0:
            kc.set_rand_seed(7)
            ref = kc.Nnet.from_config(CFG, skip_splice=True)
        xs = [rng.standard_normal((N, net.input_dim)).astype(np.float32) for _ in range(STEPS)]
        ls = [rng.integers(0, net.output_dim, N).astype(np.int32) for _ in range(STEPS)]
        b, e = shard_rows(N, rank, world)
        for k in range(STEPS):
            kc.set_rand_seed(1000 + k)                             # dropout seeds are drawn lazily
            x = torch.from_numpy(xs[k][b:e]).cuda(); lab = torch.from_numpy(ls[k][b:e]).cuda()
            plain(x, lab, N)
            if ref is not None:
                kc.set_rand_seed(1000 + k)
                ref.train_step(torch.from_numpy(xs[k]).cuda(), torch.from_numpy(ls[k]).cuda())
        torch.cuda.synchronize()
        if ref is not None:
            worst = 0.0
            detail = []
            for pa, pb in zip(params(net), params(ref)):
                scale = float(pb.abs().max()) + 1e-30
                d = float((pa - pb).abs().max()) / scale
                detail.append("%.1e" % d)
                worst = max(worst, d)
                if not torch.isfinite(pa).all():
                    worst = float("inf")
            if os.environ.get("KCNN_DP_CHECK_VERBOSE"):
                print("   per tensor (W, b, prev per layer):", " ".join(detail), flush=True)
            print("dp_check world=%d reduce=%s math=%s: max relative parameter difference %.3g (tolerance %.0e, x4 steps)"
                  % (world, mode, "tf32" if math else "fp32", worst, tol), flush=True)
            ok = ok and worst <= tol * 4
        # every rank holds identical parameters
        chk = torch.stack([p.double().sum() for p in params(net)]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if lo.item() != hi.item():
            ok = False
            if rank == 0:
                print("dp_check: ranks disagree on the parameters", lo.item(), hi.item(), flush=True)
        if peer is not None and peer.failed():
            ok = False
            print("dp_check: a peer barrier timed out on rank", rank, flush=True)
        dist.barrier()
    kc.set_math_mode(0)
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
