"""2+ GPU probe: NCCL all_reduce vs torch symmetric-memory all-reduce ops on a gradient-sized
buffer (developer tool, launched with torchrun)."""
import os, sys, time
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
n = 36_700_000
x = torch.ones(n, device="cuda")


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters


t = timeit(lambda: dist.all_reduce(x))
if rank == 0:
    print("nccl all_reduce %d MB: %.3f ms  (%.0f GB/s algbw)" % (n * 4 / 1e6, t, n * 4 / t / 1e6), flush=True)
try:
    group = dist.group.WORLD
    symm.enable_symm_mem_for_group(group.group_name)
    y = symm.empty(n, dtype=torch.float32, device="cuda")
    hdl = symm.rendezvous(y, group.group_name)
    y.fill_(1.0)
    ops = [o for o in dir(torch.ops.symm_mem)]
    if rank == 0:
        print("symm_mem ops:", [o for o in ops if "reduce" in o], "multicast:", getattr(hdl, "multicast_ptr", None) not in (None, 0), flush=True)
    for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
        if not hasattr(torch.ops.symm_mem, name):
            continue
        op = getattr(torch.ops.symm_mem, name)
        try:
            t = timeit(lambda: op(y, "sum", group.group_name))
            if rank == 0:
                print("%s: %.3f ms (%.0f GB/s algbw)" % (name, t, n * 4 / t / 1e6), flush=True)
        except Exception as e:
            if rank == 0:
                print(name, "failed:", repr(e)[:200], flush=True)
except Exception as e:
    if rank == 0:
        print("symmetric memory unavailable:", repr(e)[:300], flush=True)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
