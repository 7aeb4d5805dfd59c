"""The cross-compiled library itself, inspected on the CPU (cuobjdump on the sm_100a code; no GPU): the claims
DESIGN.md makes about the kernels must be visible in the machine code.

  * the contraction kernels issue tcgen05 (SASS UTCHMMA) with TMEM loads (LDTM), fed by TMA tensor loads
    (UTMALDG) and mbarriers (SYNCS) -- not mma.sync (HMMA) and not a library;
  * the staged max-pool moves its slabs with bulk copies (UBLKCP);
  * no kernel uses local memory, only the peer-memory kernels have a stack frame (their peer table is indexed
    by the run-time rank);
  * the reduce + SGD kernel of the data-parallel step fits beside a GEMM CTA: 256 threads x <= 64 registers.
"""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kaldi-cnn_b200", "lib", "libkaldicnn_b200.so")


def _ensure_lib():
    if shutil.which("cuobjdump") is None:
        pytest.skip("needs cuobjdump")
    if not os.path.exists(LIB):                          # a fresh checkout: build first (nvcc cross-compiles)
        sys.path.insert(0, ROOT)
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="module")
def table():
    _ensure_lib()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_resources.py"), LIB], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    rows = {}
    for line in r.stdout.splitlines():
        m = re.match(r"\| `(.+?)` \| (\d+) \| (\d+) \| (\d+) \| (\d+) \| (.*) \|$", line)
        if m:
            ins = dict((k, int(v)) for k, v in (p.split() for p in m.group(6).split(", ") if p.strip()))
            rows.setdefault(m.group(1), []).append(dict(regs=int(m.group(2)), smem=int(m.group(3)), stack=int(m.group(4)),
                                                        local=int(m.group(5)), ins=ins))
    assert len(rows) > 50
    return r.stdout, rows


def test_only_sm_100a_code_is_embedded():
    _ensure_lib()
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_gemm_kernels_are_tcgen05_fed_by_tma(table):
    _, rows = table
    gemms = {k: v for k, v in rows.items() if "tma_gemm" in k}
    assert len(gemms) >= 10
    for name, variants in gemms.items():
        for v in variants:
            assert v["ins"].get("UTCHMMA", 0) > 0, name          # tcgen05.mma
            assert v["ins"].get("LDTM", 0) > 0, name             # tcgen05.ld: the accumulator comes out of TMEM
            assert v["ins"].get("UTMALDG", 0) > 0, name          # operands by TMA tensor loads
            assert v["ins"].get("SYNCS", 0) > 0, name            # mbarrier pipeline
            assert v["ins"].get("HMMA", 0) == 0, name            # no mma.sync path
            assert v["local"] == 0 and v["stack"] == 0, name
    soft = [v for k, vs in rows.items() if "gemm_tc_kernel" in k for v in vs]
    assert soft and all(v["ins"].get("UTCHMMA", 0) > 0 for v in soft)   # the software-producer kernel is tcgen05 too


def test_staged_maxpool_uses_bulk_copies(table):
    _, rows = table
    staged = [v for k, vs in rows.items() if "maxpool_staged" in k for v in vs]
    assert staged and all(v["ins"].get("UBLKCP", 0) > 0 and v["ins"].get("SYNCS", 0) > 0 for v in staged)


def test_no_local_memory_and_stack_only_in_the_peer_kernels(table):
    _, rows = table
    for name, variants in rows.items():
        for v in variants:
            assert v["local"] == 0, name
            if v["stack"] > 0:
                assert "p2p" in name or "softmax_xent" in name or "bump_seeds" in name, (name, v["stack"])


def test_reduce_sgd_kernel_fits_beside_a_gemm_cta(table):
    _, rows = table
    light = [v for k, vs in rows.items() if "p2p_reduce_sgd_light_kernel" in k for v in vs]
    assert len(light) == 2
    assert all(v["regs"] <= 64 and v["smem"] < 4096 for v in light)     # 256 threads x 64 registers = 16 K of 64 K
