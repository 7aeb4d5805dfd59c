"""ctypes view of libkaldicnn_b200.so (the L0 launcher ABI of include/cnsl-cu-kernels.h
and the component-level C API of include/kcnn_capi.h).

Nothing here computes: it declares prototypes, loads the library and converts
torch tensors into (device pointer, MatrixDim).  Import works without a GPU so
the CPU test-suite can check that every declared symbol is exported.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libkaldicnn_b200.so")

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

MATH_FP32_SIMT = 0
MATH_TF32_TC = 1
POOL_PLAIN, POOL_OVERLAP, POOL_OVERLAP2D = 0, 1, 2


class MatrixDim(ctypes.Structure):
    """include/cu-matrixdim.h"""
    _fields_ = [("rows", c_int), ("cols", c_int), ("stride", c_int)]


class Dim3(ctypes.Structure):
    """CUDA dim3, passed by value to the legacy launchers (ignored there)."""
    _fields_ = [("x", ctypes.c_uint), ("y", ctypes.c_uint), ("z", ctypes.c_uint)]


S, P, M, I, F = c_void_p, c_void_p, MatrixDim, c_int, c_float   # stream, pointer, dim, int, float

# name -> argtypes (restype None unless listed in _RESTYPES)
_PROTOS = {
    "kcnn_set_stream": [S],
    "kcnn_get_stream": [],
    "kcnn_launch_count": [],
    "kcnn_reset_launch_count": [],
    "kcnn_build_info": [],
    "kcnn_abi_version": [],
    "kcnn_profile_start": [],
    "kcnn_profile_stop": [],
    "kcnn_profile_active": [],
    "kcnn_set_pdl": [I],
    "kcnn_profile_label": [ctypes.c_char_p, ctypes.c_double, ctypes.c_double],
    "kcnn_profile_get": [I, ctypes.c_char_p, I, ctypes.c_char_p, I, ctypes.POINTER(c_float),
                         ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                         ctypes.POINTER(ctypes.c_uint)],
    # legacy launchers
    "cudaF_span_row_to_convmat": [Dim3, Dim3, P, M, P, M, I, I, I, I, I, I],
    "cudaF_convmat_to_out": [Dim3, Dim3, P, M, P, M, I, I, I],
    "cudaF_add_mat_rep_vec": [Dim3, Dim3, P, I, P, M],
    "cudaF_flip_mat": [Dim3, Dim3, P, M, I, I, I, P, M],
    "cudaF_pad_zero": [Dim3, Dim3, P, M, I, I, I, I, P, M],
    "cudaF_tp_block": [Dim3, Dim3, P, M, P, M, I],
    "cudaF_tp_inside_block": [Dim3, Dim3, P, M, P, M, I],
    "cudaF_mod_permute_row": [Dim3, Dim3, P, M, P, M, I, I],
    "cudaF_copy_rows_at": [Dim3, Dim3, P, M, P, M, I],
    "cudaF_maxpool_prop": [Dim3, Dim3, P, M, P, M, I, I, I, I, I],
    "cudaF_maxpool_backprop": [Dim3, Dim3, P, M, P, M, P, M, P, M, I, I, I, I, I],
    "cudaF_maxpoolchannel_overlap_prop": [Dim3, Dim3, P, M, P, M, I, I, I, I, I],
    "cudaF_maxpoolchannel_overlap_backprop": [Dim3, Dim3, P, M, P, M, P, M, P, M, I, I, I, I, I],
    "cudaF_maxpoolchannel_overlap2D_prop": [Dim3, Dim3, P, M, P, M, I, I, I, I, I],
    "cudaF_maxpoolchannel_overlap2D_backprop": [Dim3, Dim3, P, M, P, M, P, M, P, M, I, I, I, I, I],
    # stream-ordered launchers
    "cudaF_add_mat_rep_vec_s": [S, P, I, P, M],
    "cudaF_flip_mat_s": [S, P, M, I, I, I, P, M],
    "cudaF_pad_zero_s": [S, P, M, I, I, I, I, P, M],
    "cudaF_tp_block_s": [S, P, M, P, M, I],
    "cudaF_tp_inside_block_s": [S, P, M, P, M, I],
    "cudaF_mod_permute_row_s": [S, P, M, P, M, I, I],
    "cudaF_copy_rows_at_s": [S, P, M, P, M, I],
    "cudaF_maxpool_prop_s": [S, P, M, P, M, I, I, I, I, I, I],
    "cudaF_maxpool_backprop_s": [S, P, M, P, M, P, M, P, M, I, I, I, I, I, I, I],
    # fused entry points
    "cudaF_maxpool_prop_index": [S, P, M, P, M, P, I, I, I, I, I, I],
    "cudaF_maxpool_backprop_index": [S, P, I, P, M, P, M, I, I, I, I, I],
    "cudaF_conv2d_fprop": [S, I, P, M, P, M, P, P, M, I, I, I, I, I, I, I, I, I],
    "cudaF_conv2d_dgrad": [S, I, P, M, P, M, P, M, I, I, I, I, I, I, I, I],
    "cudaF_conv2d_wgrad": [S, I, P, M, P, M, P, M, P, P, I, I, I, I, I, I, I, I],
    "kcnn_conv2d_wgrad_workspace": [I, I, I, I, I, I, I, I, I],
    "cudaF_affine_fprop": [S, I, P, M, P, M, P, P, M],
    "cudaF_affine_dgrad": [S, I, P, M, P, M, P, M],
    "cudaF_affine_wgrad": [S, I, P, M, P, M, P, M, P],
    "cudaF_conv2d_backward": [S, I, P, M, P, M, P, M, P, M, P, M, P, P, M, P, I, F, F, F, P,
                              I, I, I, I, I, I, I, I],
    "cudaF_conv2d_fprop_staged": [S, I, P, M, P, M, P, P, M, I, I, I, I, I, I, I, I, I, P],
    "cudaF_conv2d_fprop_act": [S, I, P, M, P, M, P, P, M, I, I, I, I, I, I, I, I, I, P, I],
    "cudaF_affine_fprop_act": [S, I, P, M, P, M, P, P, M, I],
    "kcnn_conv2d_staging_floats": [I, I, I, I, I, I, I, I, I],
    "cudaF_affine_wgrad_sgd": [S, I, P, M, P, M, P, M, P, M, P, F, F, F],
    "cudaF_sgd_momentum_update": [S, P, M, P, M, P, M, F, F, F],
    "cudaF_vec_axpy": [S, P, P, I, F],
    "cudaF_sum_rows_per_map": [S, P, M, I, P],
    "cudaF_relu_fprop": [S, P, M, P, M],
    "cudaF_relu_bprop": [S, P, M, P, M, P, M],
    "cudaF_softmax_fprop": [S, P, M, P, M],
    "cudaF_softmax_bprop": [S, P, M, P, M, P, M],
    "cudaF_xent_deriv": [S, P, M, P, P, M, P],
    "cudaF_normalize_fprop": [S, P, M, P, M],
    "cudaF_normalize_bprop": [S, P, M, P, M, P, M],
    # fused training step: channels-last activations
    "kcnn_conv_time_shape_ok": [I, I, I, I, I, I],
    "kcnn_conv_full_shape_ok": [I, I, I, I, I, I],
    "cudaF_cl_to_ref": [S, P, I, I, I, P, I],
    "cudaF_conv_time_fprop_cl": [S, P, I, I, I, I, I, I, P, M, P, P, I, I, I],
    "cudaF_conv_time_dgrad_cl": [S, P, I, I, I, I, I, I, P, M, P, P],
    "cudaF_conv_time_wgrad_cl": [S, P, P, I, I, I, I, I, I, P, M, P, M, I, F, F, F],
    "cudaF_conv_full_fprop_cl": [S, P, M, I, I, I, I, I, P, M, P, P, I],
    "cudaF_conv_full_dgrad_cl": [S, P, I, I, I, I, I, I, P, M, P, M],
    "cudaF_conv_full_wgrad_cl": [S, P, M, P, I, I, I, I, I, P, M, P, M, I, F, F, F],
    "cudaF_affine_fprop_fused": [S, P, M, P, M, P, P, M, I, P, M, F, F, F, P],
    "cudaF_affine_dgrad_fused": [S, P, M, P, M, P, M, P, I, P, I, I],
    "cudaF_maxpool_prop_cl": [S, P, I, I, I, I, I, P, P, I],
    "cudaF_maxpool_backprop_cl": [S, P, P, I, P, I, I, I, I, I, P, I],
    "kcnn_colsum_batch_scratch_bytes": [P, I],
    "cudaF_colsum_batch": [S, P, I, P],
    "cudaF_softmax_xent": [S, P, M, P, M, P, P, M, P, P, I],
    "cudaF_bump_seeds": [S, P, I],
}


class ColsumJob(ctypes.Structure):
    """KcnnColsumJob of include/cnsl-cu-kernels.h"""
    _fields_ = [("src", c_void_p), ("rows", c_int), ("cols", c_int), ("ld", c_int), ("op", c_int),
                ("perm_w", c_int), ("perm_c", c_int), ("dst0", c_void_p), ("dst1", c_void_p), ("alpha", c_float)]
_RESTYPES = {
    "kcnn_get_stream": c_void_p,
    "kcnn_launch_count": ctypes.c_ulonglong,
    "kcnn_build_info": ctypes.c_char_p,
    "kcnn_abi_version": c_int,
    "kcnn_profile_stop": c_int,
    "kcnn_profile_active": c_int,
    "kcnn_set_pdl": c_int,
    "kcnn_profile_get": c_int,
    "kcnn_conv2d_wgrad_workspace": c_size_t,
    "cudaF_conv2d_backward": c_int,
    "kcnn_conv2d_staging_floats": c_size_t,
    "cudaF_conv2d_fprop_staged": c_int,
    "cudaF_conv2d_fprop_act": c_int,
    "cudaF_affine_wgrad_sgd": c_int,
    "kcnn_conv_time_shape_ok": c_int,
    "kcnn_conv_full_shape_ok": c_int,
    "cudaF_conv_time_fprop_cl": c_int,
    "cudaF_conv_time_dgrad_cl": c_int,
    "cudaF_conv_time_wgrad_cl": c_int,
    "cudaF_conv_full_fprop_cl": c_int,
    "cudaF_conv_full_dgrad_cl": c_int,
    "cudaF_conv_full_wgrad_cl": c_int,
    "cudaF_affine_fprop_fused": c_int,
    "cudaF_affine_dgrad_fused": c_int,
    "kcnn_colsum_batch_scratch_bytes": c_size_t,
    "cudaF_softmax_xent": c_int,
}

_lib = None


def declared_symbols(header_names=("cnsl-cu-kernels.h", "kcnn_capi.h")):
    """Every function name declared in include/*.h (parsed from the headers)."""
    names = []
    for h in header_names:
        path = os.path.join(ROOT, "include", h)
        if not os.path.exists(path):
            continue
        text = open(path).read()
        text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
        text = re.sub(r"//[^\n]*", " ", text)
        names += re.findall(r"\b((?:cudaF_|kcnn_)\w+)\s*\(", text)
    seen, out = set(), []
    for n in names:
        if n not in seen:
            seen.add(n)
            out.append(n)
    return out


def load(path=None):
    """Load the shared library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "libkaldicnn_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `python kaldi-cnn_b200/build.py`. There is no CPU fallback." % path)
    L = ctypes.CDLL(path)
    for name, args in _PROTOS.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            continue                      # reported by tests/test_abi_symbols.py
        fn.argtypes = args
        fn.restype = _RESTYPES.get(name)
    try:
        from . import capi_components
        capi_components.declare(L)
    except ImportError:
        pass
    _lib = L
    return L


def lib():
    return load()


# ---- torch helpers (tests / bench only) -------------------------------------

def mdim(t):
    """MatrixDim of a 2-D torch tensor whose rows are contiguous."""
    assert t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1), "rows must be contiguous"
    stride = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
    return MatrixDim(t.shape[0], t.shape[1], stride)


def ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_gpu():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("kaldi-cnn_b200 needs a CUDA device (sm_100a); no CPU fallback exists")
    load()
