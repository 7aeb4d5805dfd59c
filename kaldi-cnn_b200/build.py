"""Builds libkaldicnn_b200.so in-tree with nvcc for sm_100a.

    python kaldi-cnn_b200/build.py [--force] [--verbose]

Every .cu / .cc under csrc/ is compiled with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` and linked into
``kaldi-cnn_b200/lib/libkaldicnn_b200.so``.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libkaldicnn_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "--extended-lambda", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-DHAVE_CUDA=1", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC,
          "-I" + os.path.join(CSRC, "cnslmat")]


def _sources():
    src = []
    for pat in ("**/*.cu", "**/*.cc"):
        src += glob.glob(os.path.join(CSRC, pat), recursive=True)
    return sorted(src)


def _headers_mtime():
    m = 0.0
    for pat in ("**/*.h", "**/*.cuh"):
        for f in glob.glob(os.path.join(CSRC, pat), recursive=True):
            m = max(m, os.path.getmtime(f))
    for f in glob.glob(os.path.join(ROOT, "include", "*.h")):
        m = max(m, os.path.getmtime(f))
    return m


def _compile(src, obj, verbose):
    cmd = [NVCC] + ARCH + COMMON
    if src.endswith(".cc"):
        cmd += ["-x", "cu"]          # host C++ that includes CUDA runtime headers
    cmd += ["-c", src, "-o", obj]
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return r.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr_m = _headers_mtime()
    jobs, objs = [], []
    for s in _sources():
        o = os.path.join(OBJ, os.path.relpath(s, CSRC).replace(os.sep, "__") + ".o")
        objs.append(o)
        if (force or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(s)
                or os.path.getmtime(o) < hdr_m):
            jobs.append((s, o))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for warn in ex.map(lambda so: _compile(so[0], so[1], verbose), jobs):
                if verbose and warn:
                    print(warn)
    if jobs or not os.path.exists(LIB):
        # static cudart (nvcc default): the library needs no CUDA .so to load, and
        # torch streams are plain driver streams, so they are valid here.
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
