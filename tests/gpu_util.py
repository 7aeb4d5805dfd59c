"""Helpers shared by the -m gpu parity tests (they call the product through its C-ABI)."""
import numpy as np
import torch

import kaldi_cnn_b200 as kc
from kaldi_cnn_b200.capi import mdim, ptr, stream  # noqa: F401

L = None


def lib():
    global L
    if L is None:
        kc.capi.require_gpu()
        L = kc.lib()
    return L


def dev(a, pad=0, off=0):
    """numpy [r x c] -> CUDA tensor view with row stride c + pad and column offset off
    (off != 0 gives a CuSubMatrix-like, possibly 16-byte-misaligned view)."""
    a = np.ascontiguousarray(a)
    r, c = a.shape
    buf = torch.full((r, c + pad + off), float("nan"), dtype=torch.float32, device="cuda")
    view = buf[:, off:off + c]
    view.copy_(torch.from_numpy(a))
    return view


def dev_empty(r, c, pad=0, off=0, fill=float("nan")):
    buf = torch.full((r, c + pad + off), fill, dtype=torch.float32, device="cuda")
    return buf[:, off:off + c]


def host(t):
    return t.detach().cpu().numpy().copy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_exact(got, ref, what=""):
    got, ref = np.asarray(got, dtype=np.float32), np.asarray(ref, dtype=np.float32)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    same = bits(got) == bits(ref)
    if not same.all():
        bad = np.argwhere(~same)
        raise AssertionError("%s: %d / %d elements differ bitwise, first at %s: got %r ref %r" % (
            what, bad.shape[0], same.size, tuple(bad[0]), got[tuple(bad[0])], ref[tuple(bad[0])]))


def rel_err(got, ref):
    """max |got - ref| / max |ref| : the norm-wise relative error the tolerances are stated in."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-30))
