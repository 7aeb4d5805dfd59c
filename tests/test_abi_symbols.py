"""CPU: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes
import os

import pytest


def _lib_or_skip():
    import kaldi_cnn_b200 as kc
    if not os.path.exists(kc.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return kc, ctypes.CDLL(kc.LIB_PATH)


def test_every_declared_symbol_is_exported():
    kc, L = _lib_or_skip()
    names = kc.capi.declared_symbols()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, "declared in include/*.h but not exported: %s" % missing


def test_ctypes_prototypes_cover_the_header():
    kc, L = _lib_or_skip()
    from kaldi_cnn_b200 import capi
    declared = set(capi.declared_symbols(("cnsl-cu-kernels.h",)))
    assert declared <= set(capi._PROTOS), sorted(declared - set(capi._PROTOS))


def test_build_info_and_abi_version():
    kc, _ = _lib_or_skip()
    L = kc.load()
    assert b"sm_100a" in L.kcnn_build_info()
    assert L.kcnn_abi_version() >= 1
    assert L.kcnn_launch_count() == 0 or L.kcnn_launch_count() > 0


def test_product_does_not_reference_the_oracle():
    """The product must never import / link anything under oracle/ (no CPU fallback)."""
    import kaldi_cnn_b200 as kc
    bad = []
    for dirpath, _, files in os.walk(kc.PACKAGE_DIR):
        if os.path.basename(dirpath) in ("build", "lib", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if "oracle" in text.replace("kcnn_oracle", "oracle") and f != "capi.py":
                    for line in text.splitlines():
                        s = line.strip()
                        if "oracle" in s and ("import" in s or "#include" in s or "dlopen" in s or "CDLL" in s):
                            bad.append((f, s))
    assert not bad, bad
