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


def test_gpus_n_without_a_launcher_starts_one_process_per_gpu():
    """`python bench.py --gpus 2` outside torchrun re-executes itself through torch.distributed.run (one process per
    GPU, 127.0.0.1 rendezvous); checked here with the CPU arm: exactly ONE line, from rank 0, with n_gpus 2."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--cpu-rows", "16"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2


def test_product_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run("--steps", "1", "--warmup", "0", "--no-cpu", "--no-kernels")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


def test_step_roofline_matches_the_committed_launch_list():
    """tools/step_roofline.py joins the measured per-launch breakdown of the committed bench line
    (profiles/r02_bench_final.json: step_kernels, gpu_launches_per_step) with the committed ncu launch list
    of the same command: every launch of the step appears in the ncu list, and the share of each kernel
    family in the step agrees between CUDA-event timing and ncu (cold, serialised) to a few points."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "step_roofline.py")], capture_output=True,
                       text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1000:]
    bench = json.loads([l for l in open(os.path.join(ROOT, "profiles", "r02_bench_final.json")) if l.startswith("{")][-1])
    rows = [l for l in r.stdout.splitlines() if l.startswith("| ") and l.split("|")[1].strip().isdigit()]
    assert len(rows) == len(bench["step_kernels"]) == bench["gpu_launches_per_step"]
    assert abs(sum(k["share"] for k in bench["step_kernels"]) - 1.0) < 1e-6
    top = max(bench["step_kernels"], key=lambda k: k["share"])
    assert bench["roofline"]["kernel"].startswith(top["label"])          # the roofline is the largest measured share
    assert abs(bench["roofline"]["frac"] - bench["roofline"]["achieved"] / bench["roofline"]["peak"]) < 1e-9
    for l in rows:
        cells = [c.strip() for c in l.split("|")]
        label, kernel, bound = cells[2], cells[3], cells[7]
        if bound == "tensor":
            assert "tma_gemm" in kernel, l
        if "maxpool" in label:
            assert "maxpool" in kernel, l
    cross = [l for l in r.stdout.splitlines() if l.startswith("| `")]
    assert len(cross) >= 10 and not any("not in the list" in l for l in cross), cross
    for l in cross:
        cells = [c.strip() for c in l.split("|")]
        ev, ncu = float(cells[2].rstrip(" %")), float(cells[3].rstrip(" %"))
        assert abs(ev - ncu) <= 4.0, l
