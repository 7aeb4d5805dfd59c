"""Check + timing of the peer-memory all-reduce (kcnn_p2p_allreduce_f32) against NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_check.py

Every rank fills a symmetric arena with rank-dependent data, reduces sub-ranges of it with the
library's kernel and with dist.all_reduce on a copy, and compares (bit-exact at N = 2, where
a + b has one order; 1e-6 relative otherwise); then times both on the 147 MB of the C2 model.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from kaldi_cnn_b200 import capi  # noqa: E402
from kaldi_cnn_b200.dp import PeerMemoryAllReduce  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = capi.lib()
    floats = 36_700_160                       # ~ the C2 model's gradient arena (146.8 MB)
    mode = os.environ.get("KCNN_P2P_CHECK_MODE", "p2p")        # p2p (two-shot) | nvls (in-switch, multimem)
    peer = PeerMemoryAllReduce(L, dist, floats, multicast=mode == "nvls")
    g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
    ok = True
    cases = [(0, 64), (64, 4096), (4160, 20608 // 4 * 4), (1 << 20, 4 << 20), (0, peer.floats), (128, 786944 // 4 * 4)]
    for it in range(3):
        for ch, (off, ln) in enumerate(cases):
            peer.arena.copy_(torch.randn(peer.floats, device="cuda", generator=g))
            ref = peer.arena.clone()
            before = peer.arena.clone()
            dist.all_reduce(ref[off:off + ln])
            torch.cuda.synchronize(); dist.barrier()
            peer.all_reduce(off, ln, channel=ch & 1).wait()
            torch.cuda.synchronize(); dist.barrier()
            got = peer.arena
            same_in = torch.equal(got[off:off + ln], ref[off:off + ln]) if world == 2 and mode == "p2p" else \
                torch.allclose(got[off:off + ln], ref[off:off + ln], rtol=1e-6, atol=1e-6)
            same_out = torch.equal(got[:off], before[:off]) and torch.equal(got[off + ln:], before[off + ln:])
            if not (same_in and same_out):
                ok = False
                print("rank %d MISMATCH it %d case %s in %s out %s maxerr %g" % (
                    rank, it, (off, ln), same_in, same_out, float((got[off:off + ln] - ref[off:off + ln]).abs().max())), flush=True)
    # all ranks must hold bit-identical sums after a reduction of the whole arena
    peer.all_reduce(0, peer.floats).wait()
    torch.cuda.synchronize(); dist.barrier()
    chk = peer.arena.double().sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ok = ok and bool(lo.item() == hi.item())
    # timing
    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record(); b.synchronize()
        return a.elapsed_time(b) / iters
    n = peer.floats
    t_p2p = timeit(lambda: peer.all_reduce(0, n).wait())
    scratch = torch.zeros(n, device="cuda")
    t_nccl = timeit(lambda: dist.all_reduce(scratch))
    t_small_p2p = timeit(lambda: peer.all_reduce(0, 49280 // 4 * 4, channel=1).wait())
    small = torch.zeros(49280, device="cuda")
    t_small_nccl = timeit(lambda: dist.all_reduce(small))
    err = peer.failed()
    if rank == 0:
        print("p2p_check world=%d mode=%s ok=%s barrier_error=%s | %.1f MB: kcnn %.3f ms (%.0f GB/s algbw)  nccl %.3f ms (%.0f GB/s) | "
              "197 KB bucket: kcnn %.1f us  nccl %.1f us" % (
                  world, mode, ok, err, n * 4 / 1e6, t_p2p, n * 4 / t_p2p / 1e6, t_nccl, n * 4 / t_nccl / 1e6,
                  t_small_p2p * 1e3, t_small_nccl * 1e3), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok or err:
        sys.exit(1)


if __name__ == "__main__":
    main()
