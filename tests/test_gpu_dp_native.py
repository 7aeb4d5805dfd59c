"""The library's data-parallel trainer (NnetDataParallel, kcnn_nnet_dp_*) on two GPUs of one box:
P x N/P rows against 1 x N rows, parameters and (sharded, then gathered) momentum, TF32 and FP32
(tools/dp_native_check.py).  Needs >= 2 GPUs: skipped on a single-GPU box (the driver's round-end
GPU test run); bench.py carries the same check into every multi-GPU run as `rank_parity`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_step_equals_single_gpu_step():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "dp_native_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
