"""The element-wise components between the hot-path layers and the Splice front end
(8f-3), each against the oracle's restatement of upstream nnet2/nnet-component.cc, through the
component C ABI (kcnn_component_propagate / kcnn_component_backprop):

  RectifiedLinearComponent  :799-827    Propagate / Backprop bit-exact
  SoftmaxComponent          :930-1000   Propagate (floor 1e-20) / Backprop, FP32 tolerance 2e-6
  DropoutComponent          :3592-3637  Propagate bit-exact on the SAME uniform draws, Backprop bit-exact
  NormalizeComponent        :576-639    Propagate / Backprop 1e-5, floor row included
  NonlinearComponent::UpdateStats :337-363  <ValueSum> <DerivSum> <Count> after Backprop with an update
  cross-entropy objective / derivative      (nnet-update.cc contract, oracle/kcnn_oracle.c)
  SpliceComponent           :2524-2866  Propagate / Backprop bit-exact (pure data movement), contiguous
                                        and gapped contexts, const-component-dim, several chunks

The oracle has no golden vectors for these (the reference ships none): it restates the
reference's statements in order, which is what these tests pin the kernels to."""
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from tests.gpu_util import assert_bit_exact, host, lib, mdim, ptr, rel_err, stream  # noqa: E402
from kaldi_cnn_b200 import components as kc  # noqa: E402
from oracle import oracle as ora  # noqa: E402


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def special_rows(x):
    x = x.copy()
    x[0, ::7] = 0.0
    if x.shape[0] > 1:
        x[1, ::5] = -0.0
    if x.shape[0] > 3:
        x[2, :] = 0.0
        x[3, ::3] = 1e-30
    return x


@pytest.mark.parametrize("rows,dim", [(512, 2304), (37, 130), (1, 7)])
def test_relu_component(rows, dim):
    kc.use_current_stream()
    rng = np.random.default_rng(dim)
    x = special_rows(rng.standard_normal((rows, dim)).astype(np.float32))
    dy = rng.standard_normal((rows, dim)).astype(np.float32)
    comp = kc.Component.from_string("RectifiedLinearComponent dim=%d" % dim)
    y = comp.propagate(cuda(x))
    assert_bit_exact(host(y), ora.relu_propagate(x), "ReLU forward")
    dx = comp.backprop(None, y, cuda(dy), update=True)
    assert_bit_exact(host(dx), ora.relu_backprop(host(y), dy), "ReLU backward")
    # statistics: value sums, derivative sums (the 0/1 Heaviside matrix), frame count
    text = comp.write(binary=False).decode()
    vs = np.array([float(v) for v in re.search(r"<ValueSum>\s+\[([^\]]*)\]", text).group(1).split()])
    ds = np.array([float(v) for v in re.search(r"<DerivSum>\s+\[([^\]]*)\]", text).group(1).split()])
    cnt = float(re.search(r"<Count>\s+(\S+)", text).group(1))
    yh = host(y)
    vs_ref, ds_ref, cnt_ref = ora.nonlin_update_stats(yh, (yh > 0).astype(np.float32), np.zeros(dim), np.zeros(dim), 0.0)
    assert cnt == cnt_ref == rows
    assert np.array_equal(ds, ds_ref)                                   # integer counts: exact
    assert np.allclose(vs, vs_ref, rtol=2e-6, atol=2e-6 * np.abs(yh).sum(axis=0).max())    # float row sums, other order


@pytest.mark.parametrize("rows,dim", [(512, 3454), (33, 40), (3, 5000)])
def test_softmax_component_and_xent(rows, dim):
    kc.use_current_stream()
    L = lib()
    rng = np.random.default_rng(rows + dim)
    x = (rng.standard_normal((rows, dim)) * 4).astype(np.float32)
    x[0, :] = -200.0
    x[0, 3] = 60.0                                                      # everything else hits the 1e-20 floor
    comp = kc.Component.from_string("SoftmaxComponent dim=%d" % dim)
    y = comp.propagate(cuda(x))
    y_ref = ora.softmax_propagate(x)
    assert rel_err(host(y), y_ref) <= 4e-6                             # FP32 exp-sum tree vs the oracle's double sum
    assert host(y).min() >= np.float32(1e-20) and (host(y)[0, :3] == np.float32(1e-20)).all()
    lab = rng.integers(0, dim, rows).astype(np.int32)
    # objective and derivative on the device posteriors
    d = torch.zeros(rows, dim, device="cuda")
    objf = torch.zeros(1, dtype=torch.float64, device="cuda")
    L.cudaF_xent_deriv(stream(), ptr(y), mdim(y), ptr(cuda(lab)), ptr(d), mdim(d), ptr(objf))
    objf_ref, d_ref = ora.xent_objf_and_deriv(host(y), lab)
    assert_bit_exact(host(d), d_ref, "cross-entropy derivative")
    assert abs(float(objf.item()) - objf_ref) <= 1e-6 * abs(objf_ref)      # logf in FP32 summed in double vs log in double
    dx = comp.backprop(None, y, d, update=True)
    dx_ref = ora.softmax_backprop(host(y), host(d))
    assert np.abs(host(dx) - dx_ref).max() <= 4e-6
    text = comp.write(binary=False).decode()
    vs = np.array([float(v) for v in re.search(r"<ValueSum>\s+\[([^\]]*)\]", text).group(1).split()])
    vs_ref, _, cnt_ref = ora.nonlin_update_stats(host(y), None, np.zeros(dim), np.zeros(dim), 0.0)
    assert float(re.search(r"<Count>\s+(\S+)", text).group(1)) == cnt_ref == rows
    assert np.allclose(vs, vs_ref, rtol=1e-5, atol=1e-6)


def np_uniform(seed, rows, cols):
    """The uniform draws of DropoutComponent's mask (kcnn_common.cuh: mix32 on seed * 0x100000001B3 + index)."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    idx = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(cols) + np.arange(cols, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) * np.uint64(0x100000001B3) + idx) & M
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M
        x = x ^ (x >> np.uint64(31))
    return ((x >> np.uint64(32)) >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


@pytest.mark.parametrize("dp,scale", [(0.5, 0.0), (0.3, 0.2), (0.0, 0.0)])
def test_dropout_component(dp, scale):
    kc.use_current_stream()
    rows, dim = 300, 516
    rng = np.random.default_rng(int(dp * 100))
    x = special_rows(np.maximum(rng.standard_normal((rows, dim)), 0).astype(np.float32))     # post-ReLU input
    comp = kc.Component.from_string("DropoutComponent dim=%d dropout-proportion=%g dropout-scale=%g" % (dim, dp, scale))
    seed = 987654
    kc.set_rand_seed(seed)                                  # the component takes its seed at the first Propagate
    xd = cuda(x)
    for step in range(2):                                   # the device seed advances by one per Propagate
        y = comp.propagate(xd)
        u = np_uniform(seed + step, rows, dim)
        y_ref = ora.dropout_propagate(x, u, dp, scale)
        assert_bit_exact(host(y), y_ref, "dropout forward, step %d" % step)
        if dp > 0:
            assert abs(float((u - np.float32(dp) > 0).mean()) - (1 - dp)) < 0.01
    dy = rng.standard_normal((rows, dim)).astype(np.float32)
    dx = comp.backprop(xd, y, cuda(dy), update=False)
    assert_bit_exact(host(dx), ora.dropout_backprop(x, host(y), dy), "dropout backward")


@pytest.mark.parametrize("rows,dim", [(200, 300), (5, 4096)])
def test_normalize_component(rows, dim):
    kc.use_current_stream()
    rng = np.random.default_rng(dim)
    x = rng.standard_normal((rows, dim)).astype(np.float32)
    x[1, :] = 0.0                                           # hits the 2^-66 floor: f = 2^33, no second term
    x[2, :] = 1e-12
    dy = rng.standard_normal((rows, dim)).astype(np.float32)
    comp = kc.Component.from_string("NormalizeComponent dim=%d" % dim)
    y = comp.propagate(cuda(x))
    y_ref = ora.normalize_propagate(x)
    assert rel_err(host(y), y_ref) <= 1e-5
    rms = np.sqrt((host(y)[3:].astype(np.float64) ** 2).mean(axis=1))
    assert np.abs(rms - 1).max() <= 1e-5
    dx = comp.backprop(cuda(x), y, cuda(dy), update=False)
    dx_ref = ora.normalize_backprop(x, dy)
    live = np.ones(rows, bool)
    live[1] = False
    assert rel_err(host(dx)[live], dx_ref[live]) <= 1e-5
    assert rel_err(host(dx)[1], dx_ref[1]) <= 1e-6          # floor row: in_deriv = 2^33 * out_deriv exactly
    assert comp.type == "NormalizeComponent" and kc.Component.read(comp.write()).info() == comp.info()


def splice_reference(x, num_chunks, in_chunk, out_chunk, first_out_offset, context, const_dim):
    """Upstream SpliceComponent::Propagate (nnet2/nnet-component.cc:2640-2722) for contiguous ChunkInfos:
    input offsets 0 .. in_chunk-1, output offsets first_out_offset .. first_out_offset + out_chunk - 1."""
    dim = x.shape[1] - const_dim
    out = np.zeros((num_chunks * out_chunk, dim * len(context) + const_dim), np.float32)
    for ch in range(num_chunks):
        for oi in range(out_chunk):
            for c, off in enumerate(context):
                out[ch * out_chunk + oi, c * dim:(c + 1) * dim] = x[ch * in_chunk + first_out_offset + oi + off, :dim]
            if const_dim:
                out[ch * out_chunk + oi, dim * len(context):] = x[ch * in_chunk + oi, dim:]
    return out


@pytest.mark.parametrize("args,context,const_dim", [
    ("input-dim=40 left-context=10 right-context=10 const-component-dim=0", list(range(-10, 11)), 0),
    ("input-dim=13 context=-4:-1:0:3", [-4, -1, 0, 3], 0),
    ("input-dim=24 left-context=2 right-context=1 const-component-dim=4", [-2, -1, 0, 1], 4),
])
@pytest.mark.parametrize("out_chunk", [1, 5])
def test_splice_component(args, context, const_dim, out_chunk):
    kc.use_current_stream()
    L = lib()
    comp = kc.Component.from_string("SpliceComponent " + args)
    dim_in = comp.input_dim
    left, right = -context[0], context[-1]
    num_chunks, in_chunk = 6, out_chunk + left + right
    rng = np.random.default_rng(len(context) + out_chunk)
    x = rng.standard_normal((num_chunks * in_chunk, dim_in)).astype(np.float32)
    want = splice_reference(x, num_chunks, in_chunk, out_chunk, left, context, const_dim)
    assert comp.output_dim == want.shape[1]
    xd = cuda(x)
    y = torch.full((num_chunks * out_chunk, comp.output_dim), float("nan"), device="cuda")
    rc = L.kcnn_component_propagate_chunks(comp.h, num_chunks, 0, in_chunk - 1, left, left + out_chunk - 1,
                                           ptr(xd), xd.shape[0], dim_in, mdim(xd).stride,
                                           ptr(y), y.shape[0], y.shape[1], mdim(y).stride)
    assert rc == 0, L.kcnn_last_error()
    assert_bit_exact(host(y), want, "splice forward")
    if const_dim == 0 and context == list(range(context[0], context[-1] + 1)) and out_chunk == 1:
        # the training layout: one output frame per chunk, consecutive offsets -> a plain reshape
        assert np.array_equal(want, x.reshape(num_chunks, in_chunk * dim_in))
    # backward: the transpose of the gather (every input frame collects the blocks that copied it)
    dy = rng.standard_normal(want.shape).astype(np.float32)
    dx = torch.full((num_chunks * in_chunk, dim_in), float("nan"), device="cuda")
    dyd = cuda(dy)
    rc = L.kcnn_component_backprop_chunks(comp.h, num_chunks, 0, in_chunk - 1, left, left + out_chunk - 1,
                                          ptr(dyd), dyd.shape[0], mdim(dyd).stride, ptr(dx), mdim(dx).stride)
    assert rc == 0, L.kcnn_last_error()
    dim = dim_in - const_dim
    dx_ref = np.zeros_like(x, dtype=np.float64)
    for ch in range(num_chunks):
        for oi in range(out_chunk):
            for c, off in enumerate(context):
                dx_ref[ch * in_chunk + left + oi + off, :dim] += dy[ch * out_chunk + oi, c * dim:(c + 1) * dim]
            if const_dim:
                dx_ref[ch * in_chunk + oi, dim:] = dy[ch * out_chunk + oi, dim * len(context):]
    assert np.abs(host(dx) - dx_ref).max() <= 1e-5
    # serialisation round trip (token order of reference :2857-2890)
    for binary in (False, True):
        again = kc.Component.read(comp.write(binary=binary), binary=binary)
        assert again.info() == comp.info() and again.output_dim == comp.output_dim
