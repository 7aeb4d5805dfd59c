"""CPU check of the HOST ORDER in tests/ref_conv_check.py: the functions that chain the reference's
kernels into ConvolutionComponent::Backprop / Update (nnet0/nnet-component-nnet0.cc:461-540, 738-777)
are run here with the oracle's primitives standing in for the kernels (the GPU tests show the two
agree bit for bit) and must reproduce the oracle's own Backprop / Update: the chaining itself is
then right before it ever meets a GPU."""
import numpy as np
import pytest
import torch

from tests.ref_conv_check import out_shape, reference_conv_backprop, reference_conv_gradient


class CpuOps:
    """RefOps of tests/ref_conv_check.py with every kernel replaced by the oracle primitive."""

    def __init__(self, ora):
        self.o = ora

    @staticmethod
    def _n(t):
        return np.ascontiguousarray(t.numpy())

    def _chk(self, op, x, got, *a):
        """The shape RefOps would allocate on the GPU for this call == what the oracle produced, and the
        input is as wide as the member's assert demands (conv2D.cc:55-57, 247, 292, 353, 393)."""
        assert tuple(got.shape) == tuple(out_shape(op, x.shape[0], *a)), (op, got.shape, a)
        return torch.from_numpy(got)

    def tp_block(self, x, C, bs):
        assert x.shape[1] == C * bs
        return self._chk("tp_block", x, self.o.tp_block(self._n(x), C, bs), C, bs)

    def tp_inside_block(self, x, G, bs):
        assert x.shape[1] == G * bs
        return self._chk("tp_inside_block", x, self.o.tp_inside_block(self._n(x), G, bs), G, bs)

    def flip_mat(self, k, KH, KW, C, G):
        assert tuple(k.shape) == (KH * KW * C, G)
        return self._chk("flip_mat", k, self.o.flip_mat(self._n(k), KH, KW, C, G), KH, KW, C, G)

    def pad_zero(self, x, H, W, C, KH, KW):
        assert x.shape[1] == H * W * C
        return self._chk("pad_zero", x, self.o.pad_zero(self._n(x), H, W, C, KH, KW), H, W, C, KH, KW)

    def mod_permute_row(self, x, C, bs):
        assert x.shape[0] == C * bs
        return torch.from_numpy(self.o.mod_permute_row(self._n(x), C, bs))

    def conv2d(self, x, kern, H, W, C, KH, KW, G, concat):
        assert x.shape[1] == H * W * C and tuple(kern.shape) == (KH * KW * C, G)
        return self._chk("conv2d", x, self.o.conv2d(self._n(x), self._n(kern), H, W, C, KH, KW, G, concat=concat),
                         H, W, C, KH, KW, G, concat)


# (N, H, W, C, ph, pw, KH, KW, G)
CASES = [(5, 6, 7, 3, 0, 0, 3, 4, 4), (4, 1, 10, 8, 0, 0, 1, 3, 6), (3, 6, 7, 2, 1, 2, 3, 4, 5), (2, 8, 5, 3, 0, 0, 8, 2, 4)]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("branch", [-1, 0, 1])
def test_backprop_chain_order(ora, case, branch):
    N, H, W, C, ph, pw, KH, KW, G = case
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    rng = np.random.default_rng(5)
    dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    want = ora.conv_backprop(dy, k, H, W, C, ph, pw, KH, KW, G, branch=branch)
    got = reference_conv_backprop(CpuOps(ora), torch.from_numpy(dy), torch.from_numpy(k), H, W, C, ph, pw, KH, KW, G,
                                  branch=branch).numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5 * max(np.abs(want).max(), 1e-30)


@pytest.mark.parametrize("case", CASES)
def test_gradient_chain_order(ora, case):
    N, H, W, C, ph, pw, KH, KW, G = case
    OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
    rng = np.random.default_rng(6)
    x = rng.standard_normal((N, H * W * C)).astype(np.float32)
    dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
    k = rng.standard_normal((KH * KW * C, G)).astype(np.float32)
    gw, gb = ora.conv_update(x, dy, k, np.zeros(G, np.float32), np.zeros_like(k), H, W, C, ph, pw, KH, KW, G,
                             0.02, 0.0, 0.0, apply=False)[3:5]
    got_w, got_b = reference_conv_gradient(CpuOps(ora), torch.from_numpy(x), torch.from_numpy(dy), H, W, C, ph, pw,
                                           KH, KW, G)
    assert np.abs(got_w.numpy() - gw).max() <= 1e-5 * max(np.abs(gw).max(), 1e-30)
    assert np.abs(got_b.numpy() - gb).max() <= 1e-5 * max(np.abs(gb).max(), 1e-30)
