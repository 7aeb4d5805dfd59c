"""ctypes front end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.

Every function takes / returns C-contiguous numpy arrays (float32 by default,
float64 when ``dtype=np.float64``) and forwards to the C restatement in
kcnn_oracle_impl.h, which cites the reference file:line it follows.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (gcc only; no reference sources involved)."""
    src = [os.path.join(_HERE, f) for f in ("kcnn_oracle.c", "kcnn_oracle_impl.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    cmd = ["gcc", "-O3", "-mavx2", "-mfma", "-fopenmp", "-fPIC", "-shared", "-std=gnu11",
           "-Wall", "-Wno-maybe-uninitialized", "-o", _LIB_PATH, src[0], "-lm", "-ldl"]
    if subprocess.run(cmd, cwd=_HERE, capture_output=True).returncode != 0:
        cmd.remove("-fopenmp")                      # no libgomp: single-threaded build
        subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB_PATH


def use_blas(on=True):
    """Route the oracle's GEMMs through the OpenBLAS numpy bundles (the reference's CPU build
    does them in Kaldi's BLAS).  CPU-baseline legs only; returns the library path or None."""
    import glob
    L = lib()
    L.ora_use_blas.restype = ctypes.c_int
    L.ora_use_blas.argtypes = [ctypes.c_char_p]
    if not on:
        L.ora_use_blas(None)
        return None
    cands = glob.glob(os.path.join(os.path.dirname(np.__file__), "..", "numpy.libs", "libscipy_openblas64_*.so"))
    for c in sorted(cands):
        if L.ora_use_blas(os.path.abspath(c).encode()) == 1:
            return os.path.abspath(c)
    return None


def set_num_threads(n):
    """Host threads for the oracle's GEMMs; returns the number in effect."""
    L = lib()
    L.ora_set_num_threads.restype = ctypes.c_int
    return int(L.ora_set_num_threads(int(n)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.oraF_xent_objf_and_deriv.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _pre(dtype):
    return ("oraF_", ctypes.c_float) if np.dtype(dtype) == np.float32 else ("oraD_", ctypes.c_double)


def _c(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    assert a.ndim in (1, 2)
    return a


def _fn(name, dtype):
    return getattr(lib(), _pre(dtype)[0] + name)


def conv2d(x, kern, H, W, C, KH, KW, G, concat=True, dtype=np.float32):
    x, kern = _c(x, dtype), _c(kern, dtype)
    N = x.shape[0]
    OH, OW = H - KH + 1, W - KW + 1
    assert x.shape[1] == H * W * C and kern.shape == (KH * KW * C, G)
    out = np.zeros((N, OH * OW * G) if concat else (OH * OW * N, G), dtype=dtype)
    rc = _fn("conv2d", dtype)(_p(x), N, x.shape[1], _p(kern), G, H, W, C, KH, KW, G,
                              _p(out), out.shape[1], int(bool(concat)))
    assert rc == 0
    return out


def add_mat_rep_vec(m, vec, rep, dtype=np.float32):
    m, vec = _c(m, dtype).copy(), _c(vec, dtype)
    assert vec.shape[0] * rep == m.shape[1]
    T = _pre(dtype)[1]
    del T
    _fn("add_mat_rep_vec", dtype)(_p(m), m.shape[0], m.shape[1], m.shape[1], _p(vec), rep)
    return m


def flip_mat(k, KH, KW, C, G, dtype=np.float32):
    k = _c(k, dtype)
    assert k.shape == (KH * KW * C, G)
    out = np.zeros((KH * KW * G, C), dtype=dtype)
    _fn("flip_mat", dtype)(_p(k), G, KH, KW, C, G, _p(out), C)
    return out


def pad_zero(x, H, W, C, KH, KW, dtype=np.float32):
    x = _c(x, dtype)
    assert x.shape[1] == H * W * C
    PH, PW = H + 2 * (KH - 1), W + 2 * (KW - 1)
    out = np.zeros((x.shape[0], PH * PW * C), dtype=dtype)
    _fn("pad_zero", dtype)(_p(x), x.shape[0], x.shape[1], H, W, C, KH, KW, _p(out), out.shape[1])
    return out


def tp_block(x, C, bs, dtype=np.float32):
    x = _c(x, dtype)
    assert x.shape[1] == C * bs
    out = np.zeros((C, x.shape[0] * bs), dtype=dtype)
    _fn("tp_block", dtype)(_p(x), x.shape[0], x.shape[1], C, bs, _p(out), out.shape[1])
    return out


def tp_inside_block(x, G, bs, dtype=np.float32):
    x = _c(x, dtype)
    assert x.shape[1] == G * bs
    out = np.zeros((x.shape[0] * bs, G), dtype=dtype)
    _fn("tp_inside_block", dtype)(_p(x), x.shape[0], x.shape[1], G, bs, _p(out), G)
    return out


def mod_permute_row(x, C, bs, dtype=np.float32):
    x = _c(x, dtype)
    assert x.shape[0] == C * bs
    out = np.zeros_like(x)
    _fn("mod_permute_row", dtype)(_p(x), x.shape[0], x.shape[1], x.shape[1], C, bs, _p(out), x.shape[1])
    return out


def maxpool_out_cols(H, W, C, ph, pw, pc, mode=0):
    if mode == 0:
        return (H // ph) * (W // pw) * (C // pc)
    if mode == 1:
        return H * W * (C - pc + 1)
    o2 = int(np.sqrt(C)) - pc + 1
    return H * W * o2 * o2


def maxpool_prop(x, H, W, ph, pw, pc, mode=0, out_cols=None, dtype=np.float32):
    x = _c(x, dtype)
    C = x.shape[1] // (H * W)
    if out_cols is None:
        out_cols = maxpool_out_cols(H, W, C, ph, pw, pc, mode)
    out = np.zeros((x.shape[0], out_cols), dtype=dtype)
    _fn("maxpool_prop", dtype)(_p(x), x.shape[0], x.shape[1], H, W, ph, pw, pc, mode,
                               _p(out), out_cols, out_cols)
    return out


def maxpool_backprop(x, out_value, out_deriv, H, W, ph, pw, pc, mode=0, dtype=np.float32):
    x, out_value, out_deriv = _c(x, dtype), _c(out_value, dtype), _c(out_deriv, dtype)
    in_deriv = np.zeros_like(x)
    oc = out_value.shape[1]
    _fn("maxpool_backprop", dtype)(_p(x), x.shape[0], x.shape[1], _p(out_value), oc,
                                   _p(out_deriv), oc, oc, _p(in_deriv), x.shape[1],
                                   H, W, ph, pw, pc, mode)
    return in_deriv


def conv_propagate(x, lin, bias, H, W, C, pad_h, pad_w, KH, KW, G, dtype=np.float32):
    x, lin, bias = _c(x, dtype), _c(lin, dtype), _c(bias, dtype)
    OH, OW = H + 2 * pad_h - KH + 1, W + 2 * pad_w - KW + 1
    out = np.zeros((x.shape[0], OH * OW * G), dtype=dtype)
    rc = _fn("conv_propagate", dtype)(_p(x), x.shape[0], x.shape[1], _p(lin), G, _p(bias),
                                      H, W, C, pad_h, pad_w, KH, KW, G, _p(out), out.shape[1])
    assert rc == 0
    return out


def conv_backprop_uses_flip(pad_h, pad_w, KH, KW, OH, OW):
    return bool(lib().oraF_conv_backprop_uses_flip(pad_h, pad_w, KH, KW, OH, OW))


def conv_backprop(out_deriv, lin, H, W, C, pad_h, pad_w, KH, KW, G, branch=-1, dtype=np.float32):
    out_deriv, lin = _c(out_deriv, dtype), _c(lin, dtype)
    N = out_deriv.shape[0]
    in_deriv = np.zeros((N, H * W * C), dtype=dtype)
    rc = _fn("conv_backprop", dtype)(_p(out_deriv), N, out_deriv.shape[1], _p(lin), G,
                                     H, W, C, pad_h, pad_w, KH, KW, G, branch,
                                     _p(in_deriv), in_deriv.shape[1])
    assert rc == 0
    return in_deriv


def conv_update(in_value, out_deriv, lin, bias, prev, H, W, C, pad_h, pad_w, KH, KW, G,
                learning_rate, weight_decay, momentum, apply=True, dtype=np.float32):
    """Returns (lin, bias, prev, weight_grad, bias_grad); inputs are not modified."""
    in_value, out_deriv = _c(in_value, dtype), _c(out_deriv, dtype)
    lin, bias, prev = _c(lin, dtype).copy(), _c(bias, dtype).copy(), _c(prev, dtype).copy()
    grad = np.zeros_like(lin)
    bgrad = np.zeros_like(bias)
    T = _pre(dtype)[1]
    rc = _fn("conv_update", dtype)(_p(in_value), in_value.shape[0], in_value.shape[1],
                                   _p(out_deriv), out_deriv.shape[1], _p(lin), G, _p(bias),
                                   _p(prev), G, H, W, C, pad_h, pad_w, KH, KW, G,
                                   T(learning_rate), T(weight_decay), T(momentum),
                                   int(bool(apply)), _p(grad), _p(bgrad))
    assert rc == 0
    return lin, bias, prev, grad, bgrad


def fc_propagate(x, Wm, bias, dtype=np.float32):
    x, Wm, bias = _c(x, dtype), _c(Wm, dtype), _c(bias, dtype)
    out_dim, in_dim = Wm.shape
    out = np.zeros((x.shape[0], out_dim), dtype=dtype)
    _fn("fc_propagate", dtype)(_p(x), x.shape[0], in_dim, _p(Wm), in_dim, _p(bias),
                               in_dim, out_dim, _p(out), out_dim)
    return out


def fc_backprop(out_deriv, Wm, dtype=np.float32):
    out_deriv, Wm = _c(out_deriv, dtype), _c(Wm, dtype)
    out_dim, in_dim = Wm.shape
    in_deriv = np.zeros((out_deriv.shape[0], in_dim), dtype=dtype)
    _fn("fc_backprop", dtype)(_p(out_deriv), out_deriv.shape[0], out_dim, _p(Wm), in_dim,
                              in_dim, out_dim, _p(in_deriv), in_dim)
    return in_deriv


def fc_update(in_value, out_deriv, Wm, bias, prev, learning_rate, weight_decay, momentum,
              dtype=np.float32):
    in_value, out_deriv = _c(in_value, dtype), _c(out_deriv, dtype)
    Wm, bias, prev = _c(Wm, dtype).copy(), _c(bias, dtype).copy(), _c(prev, dtype).copy()
    out_dim, in_dim = Wm.shape
    T = _pre(dtype)[1]
    _fn("fc_update", dtype)(_p(in_value), in_value.shape[0], in_dim, _p(out_deriv), out_dim,
                            _p(Wm), in_dim, _p(bias), _p(prev), in_dim, in_dim, out_dim,
                            T(learning_rate), T(weight_decay), T(momentum))
    return Wm, bias, prev


def relu_propagate(x):
    x = _c(x, np.float32)
    out = np.zeros_like(x)
    lib().oraF_relu_propagate(_p(x), x.shape[0], x.shape[1], x.shape[1], _p(out), x.shape[1])
    return out


def relu_backprop(out_value, out_deriv):
    out_value, out_deriv = _c(out_value, np.float32), _c(out_deriv, np.float32)
    o = np.zeros_like(out_value)
    n = out_value.shape[1]
    lib().oraF_relu_backprop(_p(out_value), out_value.shape[0], n, n, _p(out_deriv), n, _p(o), n)
    return o


def softmax_propagate(x):
    x = _c(x, np.float32)
    out = np.zeros_like(x)
    lib().oraF_softmax_propagate(_p(x), x.shape[0], x.shape[1], x.shape[1], _p(out), x.shape[1])
    return out


def xent_objf_and_deriv(post, labels):
    post = _c(post, np.float32)
    labels = np.ascontiguousarray(labels, dtype=np.int32)
    d = np.zeros_like(post)
    objf = lib().oraF_xent_objf_and_deriv(_p(post), post.shape[0], post.shape[1], post.shape[1],
                                          _p(labels), _p(d), post.shape[1])
    return float(objf), d


def softmax_backprop(out_value, out_deriv):
    out_value, out_deriv = _c(out_value, np.float32), _c(out_deriv, np.float32)
    o = np.zeros_like(out_value)
    n = out_value.shape[1]
    lib().oraF_softmax_backprop(_p(out_value), out_value.shape[0], n, n, _p(out_deriv), n, _p(o), n)
    return o


def dropout_propagate(x, uniform, dp, low_scale):
    """DropoutComponent::Propagate on a given matrix of uniform draws."""
    x, uniform = _c(x, np.float32), _c(uniform, np.float32)
    out = np.zeros_like(x)
    n = x.shape[1]
    lib().oraF_dropout_propagate(_p(x), x.shape[0], n, n, _p(uniform), n, ctypes.c_float(dp),
                                 ctypes.c_float(low_scale), _p(out), n)
    return out


def dropout_backprop(in_value, out_value, out_deriv):
    in_value, out_value, out_deriv = (_c(a, np.float32) for a in (in_value, out_value, out_deriv))
    o = np.zeros_like(in_value)
    n = in_value.shape[1]
    lib().oraF_dropout_backprop(_p(in_value), in_value.shape[0], n, n, _p(out_value), n, _p(out_deriv), n, _p(o), n)
    return o


def normalize_propagate(x):
    x = _c(x, np.float32)
    out = np.zeros_like(x)
    n = x.shape[1]
    lib().oraF_normalize_propagate(_p(x), x.shape[0], n, n, _p(out), n)
    return out


def normalize_backprop(in_value, out_deriv):
    in_value, out_deriv = _c(in_value, np.float32), _c(out_deriv, np.float32)
    o = np.zeros_like(in_value)
    n = in_value.shape[1]
    lib().oraF_normalize_backprop(_p(in_value), in_value.shape[0], n, n, _p(out_deriv), n, _p(o), n)
    return o


def nonlin_update_stats(out_value, deriv, value_sum, deriv_sum, count):
    """NonlinearComponent::UpdateStats: returns (value_sum, deriv_sum, count) after one call."""
    out_value = _c(out_value, np.float32)
    n = out_value.shape[1]
    vs = np.ascontiguousarray(value_sum, dtype=np.float64).copy()
    ds = np.ascontiguousarray(deriv_sum, dtype=np.float64).copy()
    cnt = ctypes.c_double(count)
    d = None if deriv is None else _c(deriv, np.float32)
    lib().oraF_nonlin_update_stats(_p(out_value), out_value.shape[0], n, n, _p(d) if d is not None else None, n,
                                   _p(vs), _p(ds), ctypes.byref(cnt))
    return vs, ds, cnt.value
