"""__graft_entry__.smoke(): one small forward + backward + update of the hot path on
cuda:0 (ConvolutionComponent -> MaxpoolComponent -> FullyConnectedComponent, the C1a shapes at
N = 32), checked against the CPU oracle."""
import numpy as np


def run():
    import torch
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    torch.cuda.set_device(0)
    from kaldi_cnn_b200 import components as kc
    from oracle import oracle as ora
    ora.build()
    N, H, W, C, KH, KW, G = 32, 40, 11, 3, 40, 4, 128
    kc.set_rand_seed(1)
    for math, tol in ((0, 1e-5), (1, 1e-3)):
        kc.set_math_mode(math)
        conv = kc.Component.from_string(
            "ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=40 kernel-width=4 stride=1 "
            "group=128 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5")
        pool = kc.Component.from_string(
            "MaxpoolComponent in-height=1 in-width=8 in-channel=128 pool-height-dim=1 pool-width-dim=2 "
            "pool-channel-dim=2")
        fc = kc.Component.from_string(
            "FullyConnectedComponent input-dim=256 output-dim=1024 learning-rate=0.02 param-stddev=0.01 "
            "bias-stddev=1 weight-decay=0.0005 momentum=0.9")
        rng = np.random.default_rng(1234)
        x = rng.standard_normal((N, H * W * C)).astype(np.float32)
        dz = rng.standard_normal((N, 1024)).astype(np.float32)
        g = lambda t: t.detach().cpu().numpy().copy()
        lin, bias = g(conv.params(0)), g(conv.params(1))[0]
        Wm, b = g(fc.params(0)), g(fc.params(1))[0]
        xd, dzd = torch.from_numpy(x).cuda(), torch.from_numpy(dz).cuda()
        y = conv.propagate(xd)
        p = pool.propagate(y)
        z = fc.propagate(p)
        dp = fc.backprop(p, None, dzd, update=True)
        dy = pool.backprop(y, p, dp, update=False)
        dx = conv.backprop(xd, None, dy, update=True)
        torch.cuda.synchronize()
        # oracle chain
        y_r = ora.conv_propagate(x, lin, bias, H, W, C, 0, 0, KH, KW, G)
        p_r = ora.maxpool_prop(g(y), 1, 8, 1, 2, 2)            # pooling is bit-exact given the same input
        assert np.array_equal(g(p).view(np.uint32), p_r.view(np.uint32)), "maxpool forward not bit-exact"
        z_r = ora.fc_propagate(g(p), Wm, b)
        dp_r = ora.fc_backprop(dz, Wm)
        dy_r = ora.maxpool_backprop(g(y), g(p), g(dp), 1, 8, 1, 2, 2)
        assert np.array_equal(g(dy).view(np.uint32), dy_r.view(np.uint32)), "maxpool backward not bit-exact"
        dx_r = ora.conv_backprop(g(dy), lin, H, W, C, 0, 0, KH, KW, G)
        lin_r = ora.conv_update(x, g(dy), lin, bias, np.zeros_like(lin), H, W, C, 0, 0, KH, KW, G,
                                0.02, 0.0002, 0.9)[0]
        err = lambda a, r: float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30))
        checks = {"conv fprop": err(g(y), y_r), "fc fprop": err(g(z), z_r), "fc dgrad": err(g(dp), dp_r),
                  "conv dgrad": err(g(dx), dx_r),
                  "conv update": float(np.abs(g(conv.params(0)) - lin_r).max() / np.abs(lin_r - lin).max())}
        for k, v in checks.items():
            assert v <= tol * 4, "smoke parity failed (math=%d): %s rel err %.3g" % (math, k, v)
        print("smoke ok math=%d: %s" % (math, ", ".join("%s %.2e" % kv for kv in checks.items())))
    kc.set_math_mode(0)
    fused_step()


FUSED_CFG = """
ConvolutionComponent in-height=8 in-width=13 in-channel=1 kernel-height=8 kernel-width=4 stride=1 group=32 out-height=1 out-width=10 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5
RectifiedLinearComponent dim=320
MaxpoolComponent in-height=1 in-width=10 in-channel=32 pool-height-dim=1 pool-width-dim=1 pool-channel-dim=2
ConvolutionComponent in-height=1 in-width=10 in-channel=16 kernel-height=1 kernel-width=3 stride=1 group=64 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.5
MaxpoolComponent in-height=1 in-width=8 in-channel=64 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=1
RectifiedLinearComponent dim=256
FullyConnectedComponent input-dim=256 output-dim=128 learning-rate=0.02 param-stddev=0.05 bias-stddev=0.1 weight-decay=0.0005 momentum=0.9
RectifiedLinearComponent dim=128
FullyConnectedComponent input-dim=128 output-dim=24 learning-rate=0.02 param-stddev=0.05 bias-stddev=0 weight-decay=0.0005 momentum=0.9
SoftmaxComponent dim=24
"""


def fused_step():
    """One whole training step of a small conv / pool / FC network through NnetMinibatchUpdater's fused
    plan (the path bench.py times) against the CPU oracle's step: objective and updated weights."""
    import torch
    from kaldi_cnn_b200 import components as kc
    from oracle.cpu_nnet import CpuNnet
    kc.set_math_mode(1)
    kc.set_rand_seed(7)
    net = kc.Nnet.from_config(FUSED_CFG)
    cpu = CpuNnet(FUSED_CFG, seed=7)
    g = lambda t: t.detach().cpu().numpy().copy()
    upd = [i for i in range(net.num_components) if net.component(i).type in ("ConvolutionComponent", "FullyConnectedComponent")]
    for i in upd:
        c, L = net.component(i), cpu.layers[i]
        L["lin" if c.type == "ConvolutionComponent" else "W"] = g(c.params(0))
        L["bias"], L["prev"] = g(c.params(1))[0], g(c.params(2))
        L["wd"], L["mom"] = c.weight_decay_momentum()
    rng = np.random.default_rng(3)
    x = rng.standard_normal((64, net.input_dim)).astype(np.float32)
    lab = rng.integers(0, net.output_dim, 64).astype(np.int32)
    kc.use_current_stream()
    net.train_step(torch.from_numpy(x).cuda(), torch.from_numpy(lab).cuda())
    assert net.fused_active, "smoke: the fused plan did not engage"
    objf = net.objf_and_reset()
    cpu.forward(x)
    objf_ref = cpu.backward(lab.astype(np.int64), update=True)
    assert abs(objf - objf_ref) <= 1e-3 * abs(objf_ref), (objf, objf_ref)
    worst = 0.0
    for i in upd:
        L = cpu.layers[i]
        ref = L["lin"] if "lin" in L else L["W"]
        got = g(net.component(i).params(0))
        worst = max(worst, float(np.abs(got - ref).max() / np.abs(ref).max()))
    assert worst <= 1e-3, worst
    kc.set_math_mode(0)
    print("smoke ok fused step: objf %.6f (oracle %.6f), updated weights max rel err %.2e" % (objf, objf_ref, worst))
