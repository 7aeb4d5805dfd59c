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
