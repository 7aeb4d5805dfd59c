"""The oracle's WHOLE training step (oracle/cpu_nnet.py: CpuNnet -- the checker the GPU step is held to in
tests/test_gpu_fused_step.py) pinned against an independent formulation: the same network written with
torch.nn.functional on the CPU in FP64, gradients by autograd, the update written out from the reference's
formulas.  Nothing of the oracle's index algebra (PaddingZero / FlipMat / TpBlock / ModPermuteRow chains,
both dgrad branches) is shared with conv2d / max_pool3d / autograd, so agreement to 1e-12 pins

  * Propagate of every layer kind and the [C][W][H] layout (cnsl-cu-kernels.cu:28-32),
  * Backprop (nnet0/nnet-component-nnet0.cc:461-544, 881-892; nnet2/nnet-component.cc:813-827, 985-1000,
    1246-1247, 3634-3636) -- the input derivative of the network and, through the updated parameters,
    every intermediate derivative,
  * Update (nnet0/nnet-component-nnet0.cc:738-777, 1133-1143): lr = learning_rate / N, prev = m prev - lr wd W
    + lr dW, W += prev, b += lr db; the convolution ignoring the config line's weight-decay / momentum
    (SURVEY App. C.2), over two steps so that the momentum term is exercised.

CPU only; torch here is the independent reference, never the product."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.cpu_nnet import CpuNnet

CONFIG = """
ConvolutionComponent in-height=6 in-width=9 in-channel=2 kernel-height=3 kernel-width=4 stride=1 group=4 out-height=4 out-width=6 learning-rate=0.03 param-stddev=0.3 bias-stddev=0.5 weight-decay=0.5 momentum=0.1
MaxpoolComponent in-height=4 in-width=6 in-channel=4 pool-height-dim=2 pool-width-dim=3 pool-channel-dim=2
RectifiedLinearComponent dim=8
ConvolutionComponent in-height=2 in-width=2 in-channel=2 in-pad-height=0 in-pad-width=1 kernel-height=1 kernel-width=2 stride=1 group=3 out-height=2 out-width=3 learning-rate=0.05 param-stddev=0.4 bias-stddev=0.5
RectifiedLinearComponent dim=18
DropoutComponent dim=18 dropout-proportion=0.3 dropout-scale=0.25
FullyConnectedComponent input-dim=18 output-dim=7 learning-rate=0.02 param-stddev=0.3 bias-stddev=0.1 weight-decay=0.001 momentum=0.8
RectifiedLinearComponent dim=7
FullyConnectedComponent input-dim=7 output-dim=5 learning-rate=0.04 param-stddev=0.4 bias-stddev=0.2
SoftmaxComponent dim=5
"""
# the same stack where the convolution's input derivative takes the OTHER branch of Backprop (:499-528 vs
# :529-540): a wide kernel on a short output
CONFIG_B = """
ConvolutionComponent in-height=5 in-width=8 in-channel=3 kernel-height=5 kernel-width=6 stride=1 group=4 out-height=1 out-width=3 learning-rate=0.03 param-stddev=0.3 bias-stddev=0.5
RectifiedLinearComponent dim=12
ConvolutionComponent in-height=1 in-width=3 in-channel=4 in-pad-height=0 in-pad-width=1 kernel-height=1 kernel-width=3 stride=1 group=5 out-height=1 out-width=3 learning-rate=0.05 param-stddev=0.4 bias-stddev=0.5
MaxpoolComponent in-height=1 in-width=3 in-channel=5 pool-height-dim=1 pool-width-dim=3 pool-channel-dim=1
FullyConnectedComponent input-dim=5 output-dim=4 learning-rate=0.02 param-stddev=0.3 bias-stddev=0.1
SoftmaxComponent dim=4
"""


class TorchNet:
    """The same layers on torch tensors that share NOTHING with the oracle but the parameter values."""

    def __init__(self, oracle_net):
        self.layers = []
        for L in oracle_net.layers:
            T = dict(L)
            for k in ("lin", "bias", "prev", "W"):
                if isinstance(L.get(k), np.ndarray):        # ("W" is the FC matrix OR a layer's width)
                    T[k] = torch.tensor(np.array(L[k], dtype=np.float64), requires_grad=k in ("lin", "bias", "W"))
            self.layers.append(T)

    def forward(self, x, masks):
        a, mi = x, 0
        for L in self.layers:
            k = L["kind"]
            n = a.shape[0]
            if k == "ConvolutionComponent":
                # activations [C][W][H], H fastest; kernel row (c KW + kw) KH + kh, column g
                w = L["lin"].t().reshape(L["G"], L["C"], L["KW"], L["KH"])
                y = F.conv2d(a.reshape(n, L["C"], L["W"], L["H"]), w, L["bias"], padding=(L["pw"], L["ph"]))
                a = y.reshape(n, -1)
            elif k == "MaxpoolComponent":
                y = F.max_pool3d(a.reshape(n, 1, L["C"], L["W"], L["H"]), (L["pc"], L["pw"], L["ph"]))
                a = y.reshape(n, -1)
            elif k == "FullyConnectedComponent":
                a = a @ L["W"].t() + L["bias"]
            elif k == "RectifiedLinearComponent":
                a = torch.relu(a)
            elif k == "DropoutComponent":
                a = a * masks[mi]
                mi += 1
            elif k == "SoftmaxComponent":
                a = torch.log_softmax(a, dim=1)
        return a                                   # log posteriors

    def step(self, x, labels, masks):
        n = x.shape[0]
        x = x.clone().requires_grad_(True)
        objf = self.forward(x, masks)[torch.arange(n), labels].sum()          # sum_i log p[i, label_i]
        params = [L[k] for L in self.layers for k in ("lin", "W", "bias") if torch.is_tensor(L.get(k))]
        grads = torch.autograd.grad(objf, [x] + params)
        g = dict(zip([id(p) for p in params], grads[1:]))
        with torch.no_grad():
            for L in self.layers:
                wk = "lin" if "lin" in L else "W" if torch.is_tensor(L.get("W")) else None
                if wk is None:
                    continue
                lr = L["lr"] / n                                               # :767, :1136
                L["prev"] = L["mom"] * L["prev"] - lr * L["wd"] * L[wk] + lr * g[id(L[wk])]
                L[wk] += L["prev"]
                L["bias"] += lr * g[id(L["bias"])]
        return float(objf.detach()), grads[0].numpy()


def _masks(net, n, rng):
    ms = []
    for L in net.layers:
        if L["kind"] == "DropoutComponent":
            hi = (1.0 - L["dp"] * L["scale"]) / (1.0 - L["dp"])               # nnet2/nnet-component.cc:3605-3615
            ms.append(np.where(rng.random((n, 18)) > L["dp"], hi, L["scale"]))
    return ms


@pytest.mark.parametrize("config", [CONFIG, CONFIG_B], ids=["all-layer-kinds", "other-dgrad-branch"])
def test_oracle_training_step_against_autograd(config):
    rng = np.random.default_rng(77)
    n = 6
    net = CpuNnet(config, seed=5, dtype=np.float64)
    ref = TorchNet(net)
    conv0 = net.layers[0]
    assert (conv0["wd"], conv0["mom"]) == (0.0002, 0.9)      # App. C.2: the config line's values are not applied
    nout = net.layers[-1]["dim"]
    for step in range(2):
        x = rng.standard_normal((n, net.input_dim))
        labels = rng.integers(0, nout, n)
        masks = _masks(net, n, rng)
        post = net.forward(x, dropout_masks=masks)
        objf = net.backward(labels, update=True)
        t_objf, t_dx = ref.step(torch.tensor(x), torch.tensor(labels), [torch.tensor(m) for m in masks])
        assert np.allclose(post.sum(axis=1), 1.0, atol=1e-12)
        assert abs(objf - t_objf) <= 1e-12 * max(1.0, abs(t_objf)), (step, objf, t_objf)
        assert np.abs(net.input_deriv - t_dx).max() <= 1e-12 * max(1.0, np.abs(t_dx).max()), step
        assert np.abs(t_dx).max() > 1e-4                      # the comparison is not of zeros
        for i, (L, T) in enumerate(zip(net.layers, ref.layers)):
            for k in ("lin", "W", "bias", "prev"):
                if isinstance(L.get(k), np.ndarray):
                    want = T[k].detach().numpy()
                    err = np.abs(np.asarray(L[k]) - want).max() / max(np.abs(want).max(), 1e-30)
                    assert err <= 1e-11, (step, i, L["kind"], k, err)


def test_both_dgrad_branches_are_covered():
    """Between them the convolutions of the two configurations take both branches of Backprop
    (nnet0/nnet-component-nnet0.cc:499-528 no-flip, :529-540 flip; the rule of SURVEY 8a11)."""
    from oracle import oracle as ora
    taken = set()
    for config in (CONFIG, CONFIG_B):
        for L in CpuNnet(config, seed=1, dtype=np.float64).layers:
            if L["kind"] == "ConvolutionComponent":
                taken.add(bool(ora.conv_backprop_uses_flip(L["ph"], L["pw"], L["KH"], L["KW"], L["OH"], L["OW"])))
    assert taken == {True, False}


def test_glue_oracles_against_autograd():
    """The FP32 restatements of the element-wise glue (oracle.py: normalize / softmax / dropout / relu, after
    nnet2/nnet-component.cc:576-639, 930-1000, 3592-3637, 799-827) against autograd of their defining formulas."""
    from oracle import oracle as ora
    rng = np.random.default_rng(3)
    x = rng.standard_normal((7, 33)).astype(np.float32)
    x[2] *= 1e-12                                                # a row whose mean square is under the floor 2^-66
    dy = rng.standard_normal((7, 33)).astype(np.float32)
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    # NormalizeComponent: y = x * max(mean(x^2), 2^-66)^-1/2; a floored row passes no derivative through the norm
    yt = xt * torch.clamp((xt * xt).mean(dim=1, keepdim=True), min=2.0 ** -66) ** -0.5
    (gt,) = torch.autograd.grad(yt, xt, torch.tensor(dy.astype(np.float64)))
    y, g = ora.normalize_propagate(x), ora.normalize_backprop(x, dy)
    assert np.abs(y - yt.detach().numpy()).max() <= 1e-5 * np.abs(yt.detach().numpy()).max()
    for r in range(7):                                           # row by row: row 2 is 1e12 times larger
        want = gt[r].numpy()
        assert np.abs(g[r] - want).max() <= 2e-5 * np.abs(want).max(), r
    assert np.abs(y[2] - x[2] * 2.0 ** 33).max() <= 1e-5 * np.abs(x[2] * 2.0 ** 33).max()
    # SoftmaxComponent: Backprop is y * (d - <y, d>)
    x = rng.standard_normal((5, 11)).astype(np.float32) * 3
    xt = torch.tensor(x.astype(np.float64), requires_grad=True)
    yt = torch.softmax(xt, dim=1)
    d = rng.standard_normal((5, 11)).astype(np.float32)
    (gt,) = torch.autograd.grad(yt, xt, torch.tensor(d.astype(np.float64)))
    y = ora.softmax_propagate(x)
    assert np.abs(y - yt.detach().numpy()).max() <= 1e-6
    assert np.abs(ora.softmax_backprop(y, d) - gt.numpy()).max() <= 1e-5 * max(1.0, np.abs(gt.numpy()).max())
    # RectifiedLinearComponent: the gate is the OUTPUT (> 0)
    y = ora.relu_propagate(x)
    assert np.array_equal(y, np.maximum(x, 0))
    assert np.array_equal(ora.relu_backprop(y, d), np.where(x > 0, d, 0).astype(np.float32))
    # DropoutComponent: out = in * mask with mask in {scale, (1 - p scale) / (1 - p)}; Backprop = d * out / in
    u = rng.random((5, 11)).astype(np.float32)
    p, lo = 0.3, 0.25
    out = ora.dropout_propagate(x, u, p, lo)
    ratio = out / x
    hi = (1 - p * lo) / (1 - p)
    assert np.all((np.abs(ratio - lo) < 1e-6) | (np.abs(ratio - hi) < 1e-6))
    assert np.abs(ora.dropout_backprop(x, out, d) - d * ratio).max() <= 1e-6 * np.abs(d).max() * hi
    # cross-entropy: objf = sum log p[label], derivative 1 / p at the label
    lab = rng.integers(0, 11, 5).astype(np.int32)
    y = ora.softmax_propagate(x)
    objf, dd = ora.xent_objf_and_deriv(y, lab)
    want = np.zeros_like(y)
    want[np.arange(5), lab] = 1.0 / y[np.arange(5), lab]
    assert np.allclose(dd, want, rtol=1e-6)
    assert abs(objf - np.log(y[np.arange(5), lab].astype(np.float64)).sum()) <= 1e-5
