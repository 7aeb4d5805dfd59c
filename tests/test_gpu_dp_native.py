"""The library's data-parallel trainer (NnetDataParallel, kcnn_nnet_dp_*) on two GPUs of one box:
P x N/P rows against 1 x N rows, parameters and (sharded, then gathered) momentum, TF32 and FP32
(tools/dp_native_check.py).  The two-rank test needs >= 2 GPUs and is skipped on a single-GPU box (the
driver's round-end GPU test run), where the world-size-1 test still drives the whole trainer; bench.py
carries the P-rank check into every multi-GPU run as `rank_parity`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(world, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_native_check.py")]
    env = dict(os.environ)
    if world == 1:
        env["CUDA_VISIBLE_DEVICES"] = env.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]    # one rank, one device
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


def test_one_rank_trainer_equals_the_plain_step():
    """World size 1 on any GPU box: the whole trainer -- parameters relocated into the arena, deferred updates,
    grouped reduce + SGD launches (no peers to sum), pipelined rotation, host-buffer entry point, momentum
    gather -- against the plain single-GPU step (FP32: identical up to the order of the bias sums)."""
    _run(1, 29530)


def test_two_rank_step_equals_single_gpu_step():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run(2, 29531)
