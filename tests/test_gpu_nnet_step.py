"""NnetMinibatchUpdater::TrainStep (kcnn_nnet_train_minibatch_host): the library records the
whole minibatch step into a CUDA graph on its second call and replays it afterwards.  The
replayed steps must be THE SAME steps as the eager ones: same objective, same parameters, same
nonlinearity statistics -- also across a learning-rate change, which invalidates the graph.
Reference behaviour being reproduced: nnet2's NnetUpdater loop = Propagate
through all components, objective + derivative, Backprop in reverse with the update inside."""
import re

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from kaldi_cnn_b200 import components as kc  # noqa: E402

CFG = """
ConvolutionComponent in-height=8 in-width=12 in-channel=1 kernel-height=8 kernel-width=3 stride=1 group=32 out-height=1 out-width=10 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5
RectifiedLinearComponent dim=320
MaxpoolComponent in-height=1 in-width=10 in-channel=32 pool-height-dim=1 pool-width-dim=1 pool-channel-dim=2
ConvolutionComponent in-height=1 in-width=10 in-channel=16 kernel-height=1 kernel-width=3 stride=1 group=64 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5
RectifiedLinearComponent dim=512
ConvolutionComponent in-height=1 in-width=8 in-channel=64 kernel-height=1 kernel-width=3 stride=1 group=64 out-height=1 out-width=6 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5
MaxpoolComponent in-height=1 in-width=6 in-channel=64 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=1
RectifiedLinearComponent dim=192
FullyConnectedComponent input-dim=192 output-dim=256 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.1 weight-decay=0.0005 momentum=0.9
RectifiedLinearComponent dim=256
DropoutComponent dim=256 dropout-proportion=0.5 dropout-scale=0.0
FullyConnectedComponent input-dim=256 output-dim=40 learning-rate=0.02 param-stddev=0.05 bias-stddev=0 weight-decay=0.0005 momentum=0.9
SoftmaxComponent dim=40
"""


def _counts(net):
    return [float(x) for x in re.findall(r"<Count>\s+(\S+)", net.write(binary=False).decode())]


def _params(net):
    out = []
    for i in range(net.num_components):
        c = net.component(i)
        if c.type in ("ConvolutionComponent", "FullyConnectedComponent"):
            out += [c.params(k).detach().clone() for k in range(3)]
    return out


@pytest.mark.parametrize("math", [0, 1], ids=["fp32", "tf32"])
def test_graph_replayed_steps_equal_eager_steps(math):
    kc.set_math_mode(math)
    N = 96
    nets = []
    for _ in range(2):
        kc.set_rand_seed(11)
        nets.append(kc.Nnet.from_config(CFG))
    a, b = nets
    rng = np.random.default_rng(5)
    stream = torch.cuda.Stream()
    hx = torch.empty(N, a.input_dim).pin_memory()
    hl = torch.empty(N, dtype=torch.int32).pin_memory()
    replayed = []
    with torch.cuda.stream(stream):
        kc.use_current_stream()
        for step in range(8):
            if step == 4:                         # invalidates the recorded graph
                for net in nets:
                    for i in range(net.num_components):
                        if net.component(i).type == "FullyConnectedComponent":
                            net.component(i).set_learning_rate(0.005)
            x = rng.standard_normal((N, a.input_dim)).astype(np.float32)
            lab = rng.integers(0, a.output_dim, N).astype(np.int32)
            hx.copy_(torch.from_numpy(x)); hl.copy_(torch.from_numpy(lab))
            if step == 0:
                kc.set_rand_seed(77)              # DropoutComponent draws its seed at the first Propagate
            objf_a = a.train_minibatch_host(hx.numpy(), hl.numpy())       # eager, record, replay x2, eager, record, ...
            replayed.append(a.last_step_replayed)
            xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda()
            if step == 0:
                kc.set_rand_seed(77)
            b.train_step(xd, ld)                                          # always eager
            objf_b = b.objf_and_reset()
            assert np.isfinite(objf_a) and abs(objf_a - objf_b) <= 1e-6 * abs(objf_b), (step, objf_a, objf_b)
        stream.synchronize()
    for pa, pb in zip(_params(a), _params(b)):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-8), float((pa - pb).abs().max())
    # eager, record + launch, replay, replay; learning-rate change: eager, record + launch, replay x2
    assert replayed == [False, True, True, True, False, True, True, True], replayed
    ca, cb = _counts(a), _counts(b)
    assert ca == cb and len(ca) >= 5 and all(c == 8 * N for c in ca[:4]), (ca, cb)
    kc.set_math_mode(0)
    kc.use_current_stream()


def test_legacy_default_stream_runs_eagerly():
    """The legacy default stream cannot be captured: TrainStep must simply run every step."""
    kc.set_math_mode(0)
    kc.set_rand_seed(3)
    net = kc.Nnet.from_config(CFG)
    kc.use_current_stream()
    rng = np.random.default_rng(1)
    x = rng.standard_normal((32, net.input_dim)).astype(np.float32)
    lab = rng.integers(0, net.output_dim, 32).astype(np.int32)
    vals = [net.train_minibatch_host(x, lab) for _ in range(4)]
    assert all(np.isfinite(v) for v in vals) and vals[3] > vals[0]        # the objective improves on a fixed batch


@pytest.mark.parametrize("math", [0, 1], ids=["fp32", "tf32"])
def test_fused_relu_epilogue_equals_separate_relu_component(math):
    """[Convolution | FullyConnected] + RectifiedLinear as ONE launch (Component::PropagateRelu:
    the ReLU of upstream nnet2/nnet-component.cc:799-811 applied in the GEMM epilogue) gives
    bit-identical activations, parameters and statistics to the two components run separately."""
    kc.set_math_mode(math)
    N = 64
    nets = []
    for fuse in (True, False):
        kc.set_rand_seed(21)
        net = kc.Nnet.from_config(CFG)
        net.set_fusion(fuse)
        nets.append(net)
    a, b = nets
    rng = np.random.default_rng(8)
    kc.use_current_stream()
    for step in range(3):
        x = torch.from_numpy(rng.standard_normal((N, a.input_dim)).astype(np.float32)).cuda()
        lab = torch.from_numpy(rng.integers(0, a.output_dim, N).astype(np.int32)).cuda()
        for net in nets:
            if step == 0:
                kc.set_rand_seed(5)
            net.train_step(x, lab)
        assert a.objf_and_reset() == b.objf_and_reset()
        for i in (2, 5, 10):          # ReLU outputs after conv1, conv2, fc1
            assert torch.equal(a.activation(i), b.activation(i)), i
    for pa, pb in zip(_params(a), _params(b)):
        assert torch.equal(pa, pb)
    assert _counts(a) == _counts(b)
    kc.set_math_mode(0)


def test_pipelined_host_steps_equal_synchronous_steps():
    """kcnn_nnet_train_minibatch_host_async (double-buffered staging, copy stream, one recorded
    graph per device slot) trains exactly like the synchronous call, and the caller may overwrite
    its buffers right after each call."""
    kc.set_math_mode(1)
    N, steps = 64, 9
    nets = []
    for _ in range(2):
        kc.set_rand_seed(31)
        nets.append(kc.Nnet.from_config(CFG))
    a, b = nets
    rng = np.random.default_rng(12)
    xs = [rng.standard_normal((N, a.input_dim)).astype(np.float32) for _ in range(steps)]
    ls = [rng.integers(0, a.output_dim, N).astype(np.int32) for _ in range(steps)]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        kc.use_current_stream()
        buf_x, buf_l = np.empty_like(xs[0]), np.empty_like(ls[0])
        kc.set_rand_seed(9)
        for k in range(steps):
            buf_x[...] = xs[k]; buf_l[...] = ls[k]
            a.train_minibatch_host_async(buf_x, buf_l)
            buf_x.fill(np.nan)                       # the staged copy must already be independent of it
        total_a = a.objf_and_reset()
        assert a.last_step_replayed
        kc.set_rand_seed(9)
        total_b = sum(b.train_minibatch_host(xs[k], ls[k]) for k in range(steps))
        stream.synchronize()
    assert np.isfinite(total_a) and abs(total_a - total_b) <= 1e-6 * abs(total_b), (total_a, total_b)
    for pa, pb in zip(_params(a), _params(b)):
        assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-8), float((pa - pb).abs().max())
    assert _counts(a) == _counts(b)
    assert np.isfinite(a.running_objf)
    kc.set_math_mode(0)
    kc.use_current_stream()
