"""GPU parity, bit-exact: max pooling (forward, reference-exact backward, index mode) and
the data-movement members, through the extern "C" launchers, against the CPU oracle.

Cases follow SURVEY 8(d): C1a / C1b pools, the nnet.config time pool, the C4 sweep shapes,
tie-heavy inputs (post-ReLU zeros, quantised values), NaN / inf / below-sentinel rows,
pitched and 16-byte-misaligned (CuSubMatrix-like) views, empty matrices.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import lib, dev, dev_empty, host, assert_bit_exact, mdim, ptr, stream  # noqa: E402
from kaldi_cnn_b200.capi import Dim3  # noqa: E402

# (N, H, W, C, ph, pw, pc)
POOLS = [
    (256, 1, 8, 128, 1, 2, 2),     # C1a
    (64, 33, 9, 64, 3, 3, 2),      # C1b 3x3x2
    (128, 1, 12, 256, 1, 2, 1),    # nnet.config time pool
    (64, 1, 16, 2000, 1, 2, 1),    # C4 fbank_conv.sh:249
    (64, 1, 1, 4000, 1, 1, 5),     # C4 pure intermap
    (64, 1, 8, 2000, 1, 2, 10),    # C4 run_conv.sh:56-57
    (3, 4, 6, 6, 2, 3, 3),         # small ragged
    (5, 1, 24, 7, 1, 3, 1),        # pw = 3 vector path
    (5, 1, 16, 6, 1, 4, 2),        # pw = 4 vector path
    (7, 1, 6, 10, 1, 2, 5),        # OW = 3: scalar path on a time-axis shape
    (1, 2, 2, 2, 2, 2, 2),         # single output
]


def _inputs(kind, N, cols, rng):
    x = rng.standard_normal((N, cols)).astype(np.float32)
    if kind == "relu":
        x = np.maximum(x, 0)
    elif kind == "quant":
        x = np.round(x * 4) / 4
    elif kind == "special":
        flat = x.reshape(-1)
        idx = rng.integers(0, flat.size, size=max(4, flat.size // 16))
        vals = np.array([np.nan, np.inf, -np.inf, -3e20, -1e20, 0.0, -0.0, 1e-45], dtype=np.float32)
        flat[idx] = vals[rng.integers(0, len(vals), size=idx.size)]
        if N > 1:
            x[1, :] = np.nan
        if N > 2:
            x[2, :] = -2e20
    return x.astype(np.float32)


@pytest.mark.parametrize("shape", POOLS)
@pytest.mark.parametrize("kind", ["randn", "relu", "quant", "special"])
@pytest.mark.parametrize("layout", ["packed", "pitched", "misaligned"])
def test_maxpool_plain_fwd_bwd(ora, shape, kind, layout):
    N, H, W, C, ph, pw, pc = shape
    rng = np.random.default_rng(hash((shape, kind)) % (2 ** 31))
    x = _inputs(kind, N, H * W * C, rng)
    y_ref = ora.maxpool_prop(x, H, W, ph, pw, pc)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = ora.maxpool_backprop(x, y_ref, dy, H, W, ph, pw, pc)

    pad, off = {"packed": (0, 0), "pitched": (12, 0), "misaligned": (5, 3)}[layout]
    L = lib()
    xd = dev(x, pad, off)
    yd = dev_empty(N, y_ref.shape[1], pad, off)
    L.cudaF_maxpool_prop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), H, W, ph, pw, pc, 0)
    assert_bit_exact(host(yd), y_ref, "maxpool_prop")

    dyd = dev(dy, pad + 4, off)        # its own pitch: the reference kernel assumes out_value's
    # (a) L1 semantics: only matching elements written, the rest untouched (here: sentinel 7)
    dxd = dev_empty(N, x.shape[1], pad, off, fill=7.0)
    L.cudaF_maxpool_backprop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(dyd), mdim(dyd),
                               ptr(dxd), mdim(dxd), H, W, ph, pw, pc, 0, 0)
    routed = ora.maxpool_backprop(x, y_ref, np.ones_like(dy), H, W, ph, pw, pc) != 0
    exp = np.where(routed, dx_ref, np.float32(7.0))
    assert_bit_exact(host(dxd), exp, "maxpool_backprop untouched-others")
    # (b) fused zero fill (what MaxpoolComponent::Backprop uses)
    dxd = dev_empty(N, x.shape[1], pad, off)
    L.cudaF_maxpool_backprop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(dyd), mdim(dyd),
                               ptr(dxd), mdim(dxd), H, W, ph, pw, pc, 0, 1)
    assert_bit_exact(host(dxd), dx_ref, "maxpool_backprop zero_others")
    torch.cuda.synchronize()


@pytest.mark.parametrize("shape", POOLS)
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_maxpool_index_mode(ora, shape, layout):
    """Index-routed backward == reference routing when every window has a unique maximum,
    and routes to the FIRST maximum (c -> w -> h order) under ties."""
    N, H, W, C, ph, pw, pc = shape
    rng = np.random.default_rng(11)
    x = rng.permutation(N * H * W * C).reshape(N, -1).astype(np.float32)   # all distinct
    y_ref = ora.maxpool_prop(x, H, W, ph, pw, pc)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = ora.maxpool_backprop(x, y_ref, dy, H, W, ph, pw, pc)
    pad, off = {"packed": (0, 0), "misaligned": (5, 3)}[layout]
    L = lib()
    xd, yd = dev(x, pad, off), dev_empty(N, y_ref.shape[1], pad, off)
    idx = torch.zeros((N, y_ref.shape[1] + (0 if layout == "packed" else 3)), dtype=torch.uint8, device="cuda")
    L.cudaF_maxpool_prop_index(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(idx), idx.stride(0),
                               H, W, ph, pw, pc)
    assert_bit_exact(host(yd), y_ref, "maxpool_prop_index values")
    dyd, dxd = dev(dy, pad, off), dev_empty(N, x.shape[1], pad, off)
    L.cudaF_maxpool_backprop_index(stream(), ptr(idx), idx.stride(0), ptr(dyd), mdim(dyd), ptr(dxd), mdim(dxd),
                                   H, W, ph, pw, pc)
    assert_bit_exact(host(dxd), dx_ref, "maxpool_backprop_index")
    # ties: constant input -> the first window element (position 0) receives err, nothing else
    x0 = np.zeros_like(x)
    xd = dev(x0, pad, off)
    L.cudaF_maxpool_prop_index(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(idx), idx.stride(0),
                               H, W, ph, pw, pc)
    assert int(idx[:, :y_ref.shape[1]].max()) == 0
    L.cudaF_maxpool_backprop_index(stream(), ptr(idx), idx.stride(0), ptr(dyd), mdim(dyd), ptr(dxd), mdim(dxd),
                                   H, W, ph, pw, pc)
    got = host(dxd)
    assert np.count_nonzero(got) <= dy.size and np.isclose(got.sum(), dy.sum(), rtol=1e-4, atol=1e-3)


OVERLAPS = [
    # (N, H, W, C, pc, mode)
    (16, 1, 8, 12, 3, 1),
    (8, 2, 3, 9, 4, 1),
    (16, 1, 8, 16, 2, 2),      # 4x4 map, 2x2 window -> 3x3
    (8, 2, 2, 25, 3, 2),       # 5x5 map, 3x3 window -> 3x3
]


@pytest.mark.parametrize("shape", OVERLAPS)
@pytest.mark.parametrize("kind", ["randn", "relu"])
def test_maxpool_overlap_modes(ora, shape, kind):
    N, H, W, C, pc, mode = shape
    rng = np.random.default_rng(13)
    x = _inputs(kind, N, H * W * C, rng)
    y_ref = ora.maxpool_prop(x, H, W, 1, 1, pc, mode=mode)
    dy = rng.standard_normal(y_ref.shape).astype(np.float32)
    dx_ref = ora.maxpool_backprop(x, y_ref, dy, H, W, 1, 1, pc, mode=mode)
    L = lib()
    xd, yd = dev(x, 3, 1), dev_empty(N, y_ref.shape[1], 3, 1)
    L.cudaF_maxpool_prop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), H, W, 1, 1, pc, mode)
    assert_bit_exact(host(yd), y_ref, "overlap prop")
    dyd, dxd = dev(dy, 2, 0), dev_empty(N, x.shape[1], 3, 1, fill=0.0)
    L.cudaF_maxpool_backprop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), ptr(dyd), mdim(dyd),
                               ptr(dxd), mdim(dxd), H, W, 1, 1, pc, mode, 0)
    assert_bit_exact(host(dxd), dx_ref, "overlap backprop")


def test_maxpool_empty_and_legacy_abi(ora):
    L = lib()
    xd, yd = dev_empty(0, 16), dev_empty(0, 8)
    L.cudaF_maxpool_prop_s(stream(), ptr(xd), mdim(xd), ptr(yd), mdim(yd), 1, 8, 1, 2, 1, 0)
    # legacy launcher: dim3 Gr, Bl by value, ignored; runs on kcnn_set_stream()'s stream
    rng = np.random.default_rng(3)
    x = rng.standard_normal((9, 1 * 8 * 6)).astype(np.float32)
    y_ref = ora.maxpool_prop(x, 1, 8, 1, 2, 3)
    xd, yd = dev(x), dev_empty(9, y_ref.shape[1])
    L.kcnn_set_stream(stream())
    L.cudaF_maxpool_prop(Dim3(1, 1, 1), Dim3(16, 16, 1), ptr(xd), mdim(xd), ptr(yd), mdim(yd), 1, 8, 1, 2, 3)
    assert_bit_exact(host(yd), y_ref, "legacy cudaF_maxpool_prop")
    L.kcnn_set_stream(None)


# ---------------------------------------------------------------- permutes --

@pytest.mark.parametrize("N,C,bs", [(256, 3, 440), (7, 5, 1), (33, 128, 8), (64, 12, 31), (2, 1, 3), (40, 70000, 1)])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_tp_block(ora, N, C, bs, layout):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((N, C * bs)).astype(np.float32)
    ref = ora.tp_block(x, C, bs)
    pad, off = (0, 0) if layout == "packed" else (7, 1)
    xd, od = dev(x, pad, off), dev_empty(C, N * bs, pad, off)
    lib().cudaF_tp_block_s(stream(), ptr(xd), mdim(xd), ptr(od), mdim(od), bs)
    assert_bit_exact(host(od), ref, "tp_block")


@pytest.mark.parametrize("N,G,bs", [(256, 128, 8), (5, 64, 297), (9, 3, 1), (17, 33, 5), (3, 512, 2)])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_tp_inside_block(ora, N, G, bs, layout):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((N, G * bs)).astype(np.float32)
    ref = ora.tp_inside_block(x, G, bs)
    pad, off = (0, 0) if layout == "packed" else (7, 1)
    xd, od = dev(x, pad, off), dev_empty(N * bs, G, pad, off)
    lib().cudaF_tp_inside_block_s(stream(), ptr(xd), mdim(xd), ptr(od), mdim(od), bs)
    assert_bit_exact(host(od), ref, "tp_inside_block")


@pytest.mark.parametrize("C,bs,G", [(3, 160, 128), (128, 3, 256), (1, 1, 5), (7, 5, 3)])
def test_mod_permute_row(ora, C, bs, G):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((C * bs, G)).astype(np.float32)
    ref = ora.mod_permute_row(x, C, bs)
    xd, od = dev(x, 3, 1), dev_empty(C * bs, G, 5, 2)
    lib().cudaF_mod_permute_row_s(stream(), ptr(xd), mdim(xd), ptr(od), mdim(od), bs, C)
    assert_bit_exact(host(od), ref, "mod_permute_row")


@pytest.mark.parametrize("KH,KW,C,G", [(40, 4, 3, 128), (1, 3, 128, 256), (8, 3, 3, 64), (1, 1, 1, 1), (2, 3, 5, 7)])
def test_flip_mat(ora, KH, KW, C, G):
    rng = np.random.default_rng(4)
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    ref = ora.flip_mat(k, KH, KW, C, G)
    kd, fd = dev(k, 3, 1), dev_empty(KH * KW * G, C, 2, 1)
    lib().cudaF_flip_mat_s(stream(), ptr(kd), mdim(kd), KH, KW, G, ptr(fd), mdim(fd))
    assert_bit_exact(host(fd), ref, "flip_mat")


@pytest.mark.parametrize("N,H,W,C,KH,KW", [(16, 40, 11, 3, 2, 2), (4, 1, 8, 128, 1, 3), (3, 33, 9, 4, 8, 3), (2, 3, 3, 1, 1, 1)])
def test_pad_zero(ora, N, H, W, C, KH, KW):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    ref = ora.pad_zero(x, H, W, C, KH, KW)
    xd, pd = dev(x, 3, 1), dev_empty(N, ref.shape[1], 1, 2)
    lib().cudaF_pad_zero_s(stream(), ptr(xd), mdim(xd), H, W, KH, KW, ptr(pd), mdim(pd))
    assert_bit_exact(host(pd), ref, "pad_zero")


@pytest.mark.parametrize("N,G,rep", [(256, 128, 8), (7, 3, 5), (64, 64, 297), (1, 1, 1)])
@pytest.mark.parametrize("layout", ["packed", "misaligned"])
def test_add_mat_rep_vec(ora, N, G, rep, layout):
    rng = np.random.default_rng(6)
    m = rng.standard_normal((N, G * rep)).astype(np.float32)
    v = rng.standard_normal(G).astype(np.float32)
    ref = ora.add_mat_rep_vec(m, v, rep)
    pad, off = (0, 0) if layout == "packed" else (6, 1)
    md = dev(m, pad, off)
    vd = torch.from_numpy(v).cuda()
    lib().cudaF_add_mat_rep_vec_s(stream(), ptr(vd), rep, ptr(md), mdim(md))
    assert_bit_exact(host(md), ref, "add_mat_rep_vec")


def test_legacy_im2col_col2im_copy_rows(ora):
    """The reference's three-kernel Conv2D pipeline still works through the legacy launchers
    (cnslmat/conv2D.cc:105, 150, 181), with the GEMM done on the host here."""
    rng = np.random.default_rng(7)
    N, H, W, C, KH, KW, G = 6, 5, 7, 3, 2, 3, 4
    OH, OW = H - KH + 1, W - KW + 1
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    L = lib()
    L.kcnn_set_stream(stream())
    g, b = Dim3(1, 1, 1), Dim3(16, 16, 1)
    xd = dev(x, 2, 1)
    half = (OH * OW * N) // 2
    spans = []
    for r0, nr in ((0, half), (half, OH * OW * N - half)):
        sd = dev_empty(nr, KH * KW * C, 3, 0)
        L.cudaF_span_row_to_convmat(g, b, ptr(xd), mdim(xd), ptr(sd), mdim(sd), H, W, C, KH, KW, r0)
        spans.append(host(sd))
    span = np.concatenate(spans, 0)
    conv = (span.astype(np.float64) @ k.astype(np.float64)).astype(np.float32)
    cd_full = dev_empty(OH * OW * N, G, 1, 0)
    for r0, nr in ((0, half), (half, OH * OW * N - half)):
        part = dev(conv[r0:r0 + nr], 2, 0)
        L.cudaF_copy_rows_at(g, b, ptr(part), mdim(part), ptr(cd_full), mdim(cd_full), r0)
    assert_bit_exact(host(cd_full), conv, "copy_rows_at")
    od = dev_empty(N, OH * OW * G, 1, 1)
    L.cudaF_convmat_to_out(g, b, ptr(cd_full), mdim(cd_full), ptr(od), mdim(od), OH, OW, N)
    L.kcnn_set_stream(None)
    ref = ora.conv2d(x, k, H, W, C, KH, KW, G, dtype=np.float64)
    assert np.abs(host(od) - ref).max() < 1e-5 * np.abs(ref).max()
