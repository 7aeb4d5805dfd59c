"""L0 entry points of the fused step (include/cnsl-cu-kernels.h, group 4), one by one:

  * channels-last max-pool forward / backward: BIT-EXACT against the oracle's restatement of
    _maxpool_prop / _maxpool_backprop (cnsl-cu-kernels.cu:231-308) on the permuted data, including the
    tie-heavy (post-ReLU), signed-zero, NaN / inf and below-sentinel inputs of SURVEY 8d;
  * channels-last convolution forward / input-gradient / weight-gradient (+ SGD): against the FP64 oracle
    (conv2D.cc / nnet0/nnet-component-nnet0.cc:423-446, 461-544, 738-777) at TF32 tolerance 1e-3, and
    against the library's own reference-layout entry points (same GEMM: 1e-6);
  * affine forward with ReLU + dropout epilogue, affine input gradient with the ReLU / dropout gate and
    the channels-last store: against NumPy restatements of upstream nnet2/nnet-component.cc:799-827,
    1216-1258, 3592-3637;
  * batched column sums against FP64 NumPy; softmax + cross-entropy kernel BIT-EXACT against the three
    separate kernels it replaces."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import assert_bit_exact, dev, host, lib, mdim, ptr, rel_err, stream  # noqa: E402
from kaldi_cnn_b200 import capi  # noqa: E402
from oracle import oracle as ora  # noqa: E402

TF32 = 1e-3


def to_cl(a, C, W):
    """[N x C*W] reference layout ([c][w]) -> channels-last [N x W*C]."""
    n = a.shape[0]
    return np.ascontiguousarray(a.reshape(n, C, W).transpose(0, 2, 1).reshape(n, W * C))


def to_ref(a, C, W):
    n = a.shape[0]
    return np.ascontiguousarray(a.reshape(n, W, C).transpose(0, 2, 1).reshape(n, C * W))


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def pool_inputs(N, dim, rng):
    x = rng.standard_normal((N, dim)).astype(np.float32)
    sets = {"randn": x, "relu-ties": np.maximum(x, 0), "quantised": np.round(x * 4) / 4}
    s = x.copy()
    s[0, ::5] = np.nan; s[1, ::3] = np.inf; s[2, ::4] = -np.inf; s[3, :] = -3e20
    s[4, ::2] = 0.0; s[4, 1::2] = -0.0
    sets["special"] = s
    return sets


@pytest.mark.parametrize("W,C,pw,pc", [(18, 128, 1, 2), (12, 256, 2, 1), (8, 200, 2, 10), (6, 64, 3, 2)])
@pytest.mark.parametrize("ref_out", [False, True], ids=["cl-out", "ref-out"])
def test_maxpool_cl_bit_exact(W, C, pw, pc, ref_out):
    L = lib()
    rng = np.random.default_rng(W * 131 + C)
    N = 37
    WO, CO = W // pw, C // pc
    for name, x in pool_inputs(N, W * C, rng).items():
        y_ref = ora.maxpool_prop(x, 1, W, 1, pw, pc)
        xcl = cuda(to_cl(x, C, W))
        out = torch.full((N, WO * CO), float("nan"), device="cuda")
        out_relu = torch.full((N, WO * CO), float("nan"), device="cuda")
        L.cudaF_maxpool_prop_cl(stream(), ptr(xcl), N, W, C, pw, pc, ptr(out), ptr(out_relu), WO * CO if ref_out else 0)
        y = host(out) if ref_out else to_ref(host(out), CO, WO)
        yr = host(out_relu) if ref_out else to_ref(host(out_relu), CO, WO)
        assert_bit_exact(y, y_ref, "pool fwd " + name)
        with np.errstate(invalid="ignore"):
            relu_ref = np.where(y_ref > 0, y_ref, np.float32(0))
        assert_bit_exact(yr, relu_ref, "pool fwd + relu " + name)
        dy = rng.standard_normal((N, WO * CO)).astype(np.float32)
        dx_ref = ora.maxpool_backprop(x, y_ref, dy, 1, W, 1, pw, pc)
        dycl = cuda(to_cl(dy, CO, WO))
        for gate in (0, 1):
            dx = torch.full((N, W * C), float("nan"), device="cuda")
            L.cudaF_maxpool_backprop_cl(stream(), ptr(xcl), ptr(out), WO * CO if ref_out else 0, ptr(dycl), N, W, C, pw, pc,
                                        ptr(dx), gate)
            want = dx_ref
            if gate:
                with np.errstate(invalid="ignore"):
                    want = np.where(x > 0, dx_ref, np.float32(0))
            assert_bit_exact(to_ref(host(dx), C, W), want, "pool bwd gate=%d %s" % (gate, name))


CONV_SHAPES = [  # N, W, C, pw, KW, G
    (64, 18, 64, 0, 3, 128),       # conv2 of the benchmarked model
    (96, 6, 256, 0, 3, 512),       # conv5: K-split forward and backward
    (50, 4, 512, 0, 3, 512),       # conv6
    (33, 9, 40, 1, 3, 72),         # padding, ragged tiles, C and G not multiples of 32
    (16, 8, 200, 0, 5, 200),       # C3(iii)-like, scaled
    (32, 8, 2000, 0, 5, 2000),     # C3(iii) itself (fbank_conv.sh:251: C = G = 2000, 1x5), reduced batch
]


@pytest.mark.parametrize("N,W,C,pw,KW,G", CONV_SHAPES)
def test_conv_time_channels_last_vs_oracle(N, W, C, pw, KW, G):
    L = lib()
    assert L.kcnn_conv_time_shape_ok(N, W, C, pw, KW, G) == 1
    rng = np.random.default_rng(N + W + C)
    OW = W + 2 * pw - KW + 1
    x = rng.standard_normal((N, C * W)).astype(np.float32)
    k = (rng.standard_normal((KW * C, G)) * 0.05).astype(np.float32)
    b = rng.standard_normal(G).astype(np.float32)
    dy = rng.standard_normal((N, G * OW)).astype(np.float32)
    mask = rng.standard_normal((N, C * W)).astype(np.float32)
    xcl, kd, bd = cuda(to_cl(x, C, W)), dev(k), cuda(b)
    # forward: channels-last out (+ ReLU), reference-layout out
    y_ref = ora.conv_propagate(x, k, b, 1, W, C, 0, pw, 1, KW, G, dtype=np.float64)
    ycl = torch.empty(N, OW * G, device="cuda")
    assert L.cudaF_conv_time_fprop_cl(stream(), ptr(xcl), N, W, C, pw, KW, G, ptr(kd), mdim(kd), ptr(bd), ptr(ycl), 1, 0, 0)
    assert rel_err(to_ref(host(ycl), G, OW), y_ref) <= TF32
    assert L.cudaF_conv_time_fprop_cl(stream(), ptr(xcl), N, W, C, pw, KW, G, ptr(kd), mdim(kd), ptr(bd), ptr(ycl), 1, 0, 1)
    assert rel_err(to_ref(host(ycl), G, OW), np.maximum(y_ref, 0)) <= TF32
    yrf = dev(np.zeros((N, G * OW), np.float32), pad=4)
    assert L.cudaF_conv_time_fprop_cl(stream(), ptr(xcl), N, W, C, pw, KW, G, ptr(kd), mdim(kd), ptr(bd), ptr(yrf), 0,
                                      mdim(yrf).stride, 0)
    assert rel_err(host(yrf), y_ref) <= TF32
    # input gradient (+ gate)
    dx_ref = ora.conv_backprop(dy, k, 1, W, C, 0, pw, 1, KW, G, dtype=np.float64)
    dycl, mcl = cuda(to_cl(dy, G, OW)), cuda(to_cl(mask, C, W))
    dxcl = torch.empty(N, W * C, device="cuda")
    assert L.cudaF_conv_time_dgrad_cl(stream(), ptr(dycl), N, W, C, pw, KW, G, ptr(kd), mdim(kd), ptr(dxcl), None)
    assert rel_err(to_ref(host(dxcl), C, W), dx_ref) <= TF32
    assert L.cudaF_conv_time_dgrad_cl(stream(), ptr(dycl), N, W, C, pw, KW, G, ptr(kd), mdim(kd), ptr(dxcl), ptr(mcl))
    assert rel_err(to_ref(host(dxcl), C, W), np.where(mask > 0, dx_ref, 0)) <= TF32
    # weight gradient, then the momentum step in the epilogue (reference :767-773)
    lr, wd, mom = 0.02, 0.0002, 0.9
    prev = (rng.standard_normal(k.shape) * 0.01).astype(np.float32)
    upd = ora.conv_update(x, dy, k, b, prev, 1, W, C, 0, pw, 1, KW, G, lr, wd, mom, dtype=np.float64)
    dk_ref = upd[3] if len(upd) > 3 and upd[3] is not None else None
    gk = dev(np.zeros_like(k))
    zero = dev(np.zeros_like(k))
    assert L.cudaF_conv_time_wgrad_cl(stream(), ptr(xcl), ptr(dycl), N, W, C, pw, KW, G, ptr(gk), mdim(gk), ptr(zero),
                                      mdim(zero), 0, 0.0, 0.0, 0.0)
    if dk_ref is not None:
        assert rel_err(host(gk), dk_ref) <= TF32
    kk, pp = dev(k), dev(prev)
    a_grad = np.float32(np.float64(np.float32(lr) / np.float32(N)))
    a_decay = np.float32(-1.0 * np.float64(np.float32(lr) / np.float32(N)) * np.float32(wd))
    assert L.cudaF_conv_time_wgrad_cl(stream(), ptr(xcl), ptr(dycl), N, W, C, pw, KW, G, ptr(kk), mdim(kk), ptr(pp),
                                      mdim(pp), 1, mom, float(a_decay), float(a_grad))
    assert rel_err(host(pp), upd[2]) <= TF32
    step = np.abs(upd[0] - k).max()
    assert np.abs(host(kk) - upd[0]).max() <= TF32 * step * 4
    # the SGD epilogue applies exactly sgd(gradient): compare with the stored gradient, FP32 arithmetic
    gkh = host(gk)
    p_ref = np.float32(mom) * prev
    p_ref = np.float32(a_decay) * k + p_ref
    p_ref = np.float32(a_grad) * gkh + p_ref
    assert rel_err(host(pp), p_ref) <= 2e-6


def test_conv_full_channels_last_vs_oracle():
    L = lib()
    N, H, W, C, KW, G = 70, 40, 21, 1, 4, 128
    assert L.kcnn_conv_full_shape_ok(N, H, W, C, KW, G) == 1
    OW = W - KW + 1
    rng = np.random.default_rng(9)
    x = rng.standard_normal((N, C * W * H)).astype(np.float32)
    k = (rng.standard_normal((KW * H * C, G)) * 0.05).astype(np.float32)
    b = rng.standard_normal(G).astype(np.float32)
    dy = rng.standard_normal((N, G * OW)).astype(np.float32)
    xd, kd, bd = dev(x), dev(k), cuda(b)
    y_ref = ora.conv_propagate(x, k, b, H, W, C, 0, 0, H, KW, G, dtype=np.float64)
    ycl = torch.empty(N, OW * G, device="cuda")
    assert L.cudaF_conv_full_fprop_cl(stream(), ptr(xd), mdim(xd), H, W, C, KW, G, ptr(kd), mdim(kd), ptr(bd), ptr(ycl), 1)
    assert rel_err(to_ref(host(ycl), G, OW), np.maximum(y_ref, 0)) <= TF32
    dycl = cuda(to_cl(dy, G, OW))
    dx = dev(np.zeros_like(x))
    assert L.cudaF_conv_full_dgrad_cl(stream(), ptr(dycl), N, H, W, C, KW, G, ptr(kd), mdim(kd), ptr(dx), mdim(dx))
    assert rel_err(host(dx), ora.conv_backprop(dy, k, H, W, C, 0, 0, H, KW, G, dtype=np.float64)) <= TF32
    prev = np.zeros_like(k)
    upd = ora.conv_update(x, dy, k, b, prev, H, W, C, 0, 0, H, KW, G, 0.02, 0.0002, 0.9, dtype=np.float64)
    kk, pp = dev(k), dev(prev)
    a_grad = np.float32(np.float64(np.float32(0.02) / np.float32(N)))
    a_decay = np.float32(-1.0 * np.float64(np.float32(0.02) / np.float32(N)) * np.float32(0.0002))
    assert L.cudaF_conv_full_wgrad_cl(stream(), ptr(xd), mdim(xd), ptr(dycl), H, W, C, KW, G, ptr(kk), mdim(kk), ptr(pp),
                                      mdim(pp), 1, 0.9, float(a_decay), float(a_grad))
    assert rel_err(host(pp), upd[2]) <= TF32


def np_dropout_scale(seed, rows, cols, dp, low, high):
    """DropoutComponent's mask as the kernels draw it (kcnn_common.cuh: mix32 / dropout_scale_at)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) * np.uint64(0x100000001B3) + idx) & M
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        x = x ^ (x >> np.uint64(31))
    r = ((x >> np.uint64(32)) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return np.where(r - np.float32(dp) > 0, np.float32(high), np.float32(low)).astype(np.float32)


@pytest.mark.parametrize("M,K,N", [(96, 192, 256), (512, 1024, 4096), (70, 256, 3454)])
def test_affine_fused_epilogues(M, K, N):
    L = lib()
    rng = np.random.default_rng(M + K + N)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) * 0.05).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    padn = (-N) % 4                                  # Kaldi pitches rows to 16 bytes (CuDevice::PitchInElements)
    xd, wd_, bd = dev(x), dev(w), cuda(b)
    # forward: ReLU + dropout in the epilogue
    seed_val = 123456789
    seed = torch.tensor([seed_val], dtype=torch.int64, device="cuda")
    dp, low = 0.5, 0.0
    high = (1.0 - dp * low) / (1.0 - dp)
    y, yd = dev(np.zeros((M, N), np.float32), pad=padn), dev(np.zeros((M, N), np.float32), pad=padn)
    assert L.cudaF_affine_fprop_fused(stream(), ptr(xd), mdim(xd), ptr(wd_), mdim(wd_), ptr(bd), ptr(y), mdim(y), 1,
                                      ptr(yd), mdim(yd), dp, low, high, ptr(seed))
    y_ref = np.maximum(x.astype(np.float64) @ w.T.astype(np.float64) + b, 0)
    assert rel_err(host(y), y_ref) <= TF32
    scale = np_dropout_scale(seed_val, M, N, dp, low, high)
    assert_bit_exact(host(yd), host(y) * scale, "dropout epilogue")          # same mask as DropoutComponent draws
    assert abs(float((scale == 0).mean()) - dp) < 0.02
    # input gradient with the gates, plain and channels-last store
    dy = rng.standard_normal((M, N)).astype(np.float32)
    dyd = dev(dy, pad=padn)
    rx = np.maximum(rng.standard_normal((M, K)), 0).astype(np.float32)        # the producer's ReLU output
    ry = (rx * np_dropout_scale(7, M, K, 0.5, 0.0, 2.0)).astype(np.float32)   # ... and its dropout output
    rxd, ryd = dev(rx), dev(ry)
    dx_ref = dy.astype(np.float64) @ w.astype(np.float64)
    dx = dev(np.zeros((M, K), np.float32))
    assert L.cudaF_affine_dgrad_fused(stream(), ptr(dyd), mdim(dyd), ptr(wd_), mdim(wd_), ptr(dx), mdim(dx), None, 0, None, 0, 0)
    assert rel_err(host(dx), dx_ref) <= TF32
    assert L.cudaF_affine_dgrad_fused(stream(), ptr(dyd), mdim(dyd), ptr(wd_), mdim(wd_), ptr(dx), mdim(dx), ptr(rxd),
                                      mdim(rxd).stride, None, 0, 0)
    assert rel_err(host(dx), np.where(rx > 0, dx_ref, 0)) <= TF32
    assert L.cudaF_affine_dgrad_fused(stream(), ptr(dyd), mdim(dyd), ptr(wd_), mdim(wd_), ptr(dx), mdim(dx), ptr(rxd),
                                      mdim(rxd).stride, ptr(ryd), mdim(ryd).stride, 0)
    with np.errstate(divide="ignore", invalid="ignore"):
        gated = np.where(rx > 0, dx_ref * ry / np.where(rx > 0, rx, 1), 0)
    assert rel_err(host(dx), gated) <= TF32
    R = 4                                                                     # K = G * R, channels-last store
    dxcl = torch.zeros(M, K, device="cuda")
    assert L.cudaF_affine_dgrad_fused(stream(), ptr(dyd), mdim(dyd), ptr(wd_), mdim(wd_), ptr(dxcl), mdim(dxcl), ptr(rxd),
                                      mdim(rxd).stride, None, 0, R)
    assert rel_err(to_ref(host(dxcl), K // R, R), np.where(rx > 0, dx_ref, 0)) <= TF32


def test_colsum_batch_vs_numpy():
    L = lib()
    rng = np.random.default_rng(4)
    shapes = [(512, 4096, 0, 0), (9216, 128, 0, 0), (300, 70, 0, 0), (512, 2304, 18, 128), (64, 40, 0, 0), (1, 33, 0, 0)]
    ops = [1, 0, 2, 2, 3, 0]
    mats, dst0, dst1, jobs = [], [], [], (capi.ColsumJob * len(shapes))()
    for i, ((r, c, pw_, pc_), op) in enumerate(zip(shapes, ops)):
        m = rng.standard_normal((r, c)).astype(np.float32)
        md = dev(m, pad=4 if i % 2 else 0)
        mats.append((m, md))
        dt = torch.float64 if op >= 2 else torch.float32
        d0 = torch.full((c,), 0.5, dtype=dt, device="cuda")
        d1 = torch.full((c,), 0.25, dtype=dt, device="cuda")
        dst0.append(d0); dst1.append(d1)
        jobs[i] = capi.ColsumJob(md.data_ptr(), r, c, mdim(md).stride, op, pw_, pc_, d0.data_ptr(), d1.data_ptr(), 0.125)
    nbytes = L.kcnn_colsum_batch_scratch_bytes(jobs, len(shapes))
    scratch = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    for rep in range(2):                              # twice: the arrival counters must re-arm themselves
        for d in dst0: d.fill_(0.5)
        for d in dst1: d.fill_(0.25)
        L.cudaF_colsum_batch(stream(), jobs, len(shapes), ptr(scratch))
        torch.cuda.synchronize()
        for i, ((r, c, pw_, pc_), op) in enumerate(zip(shapes, ops)):
            m = mats[i][0].astype(np.float64)
            s, cnt = m.sum(axis=0), (m > 0).sum(axis=0).astype(np.float64)
            if pw_:
                s = s.reshape(pw_, pc_).T.reshape(-1)
                cnt = cnt.reshape(pw_, pc_).T.reshape(-1)
            got0, got1 = dst0[i].cpu().numpy(), dst1[i].cpu().numpy()
            tol = 1e-5 * np.abs(mats[i][0]).sum(axis=0).max()
            if op == 0:
                assert np.abs(got0 - s).max() <= tol
            elif op == 1:
                assert np.abs(got0 - (0.5 + 0.125 * s)).max() <= tol
            elif op == 2:
                assert np.abs(got0 - (0.5 + s)).max() <= tol and np.array_equal(got1, 0.25 + cnt)
            else:
                assert np.abs(got0 - (0.5 + s)).max() <= tol and np.all(got1 == 0.25)


@pytest.mark.parametrize("rows,cols", [(512, 3454), (37, 40), (5, 4096)])
def test_softmax_xent_fused_is_bit_exact(rows, cols):
    L = lib()
    rng = np.random.default_rng(rows + cols)
    x = (rng.standard_normal((rows, cols)) * 3).astype(np.float32)
    lab = rng.integers(0, cols, rows).astype(np.int32)
    xd, ld = dev(x, pad=4), cuda(lab)
    # the three kernels it replaces (kernels_elementwise.cu)
    post, d, din = dev(np.zeros_like(x)), dev(np.zeros_like(x)), dev(np.zeros_like(x))
    objf = torch.zeros(1, dtype=torch.float64, device="cuda")
    L.cudaF_softmax_fprop(stream(), ptr(xd), mdim(xd), ptr(post), mdim(post))
    L.cudaF_xent_deriv(stream(), ptr(post), mdim(post), ptr(ld), ptr(d), mdim(d), ptr(objf))
    L.cudaF_softmax_bprop(stream(), ptr(post), mdim(post), ptr(d), mdim(d), ptr(din), mdim(din))
    seeds_t = torch.tensor([10, 20], dtype=torch.int64, device="cuda")
    seed_ptrs = (ctypes.c_void_p * 2)(seeds_t.data_ptr(), seeds_t.data_ptr() + 8)
    for from_logits in (True, False):
        post2 = dev(np.zeros_like(x)) if from_logits else dev(host(post))
        din2 = dev(np.zeros_like(x), pad=8)
        objf2 = torch.zeros(1, dtype=torch.float64, device="cuda")
        assert L.cudaF_softmax_xent(stream(), ptr(xd) if from_logits else None, mdim(xd), ptr(post2), mdim(post2), ptr(ld),
                                    ptr(din2), mdim(din2), ptr(objf2), seed_ptrs, 2)
        assert_bit_exact(host(post2), host(post), "posteriors")
        assert_bit_exact(host(din2), host(din), "d objf / d logits")
        assert abs(float(objf2.item()) - float(objf.item())) <= 1e-9 * abs(float(objf.item()))
    assert seeds_t.tolist() == [12, 22]                         # each call advanced both dropout seeds once
    # and against the FP64 definition (nnet2/nnet-component.cc:930-1000 + hard-label cross-entropy)
    e = np.exp(x.astype(np.float64) - x.max(axis=1, keepdims=True))
    p = np.maximum(e / e.sum(axis=1, keepdims=True), 1e-20)
    want = -p
    want[np.arange(rows), lab] += 1.0
    assert np.abs(host(din) - want).max() <= 2e-6
    assert abs(float(objf.item()) - np.log(p[np.arange(rows), lab]).sum()) <= 1e-5 * rows
