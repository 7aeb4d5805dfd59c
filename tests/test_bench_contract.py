"""bench.py's contract, the parts that can be checked without a GPU: the reference arm's JSON line
(keys the driver reads) and the product arm's refusal to run without a CUDA device (no CPU path)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e, cwd=ROOT)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-rows", "16")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_frames_per_sec" and d["unit"] == "frames/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["steps"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu", "--no-kernels")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_step_roofline_matches_the_committed_launch_list():
    """tools/step_roofline.py's launch plan of the C2 step (76 launches) lines up with the committed
    ncu launch list it documents, kernel family by kernel family."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_roofline.py")], capture_output=True,
                       text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    rows = [l for l in r.stdout.splitlines() if l.startswith("| ") and l.split("|")[1].strip().isdigit()]
    assert len(rows) == 76
    for l in rows:
        cells = [c.strip() for c in l.split("|")]
        label, kernel, bound = cells[2], cells[3], cells[6]
        if bound == "tensor":
            assert "tma_gemm" in kernel, l
        if label.startswith("pack"):
            assert "pack_channels_last" in kernel, l
        if label.startswith("maxpool"):
            assert "maxpool" in kernel, l
        if "ReLU bwd" in label:
            assert "relu_bprop" in kernel, l
