"""Data-parallel host logic (kaldi-cnn_b200/dp.py) on CPU: two gloo ranks, each with half of
the minibatch, must reproduce the single-process step.  The compute behind the Nnet
interface is the CPU oracle here (the CUDA model implements the same interface on a GPU box;
tests/test_gpu_components.py covers that side)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import kaldi_cnn_b200  # noqa: E402,F401
from kaldi_cnn_b200.dp import (DataParallelStep, PipelinedDataParallelStep, late_components,  # noqa: E402
                               shard_rows)

H, W, C, KH, KW, G = 1, 10, 8, 1, 3, 12
OW = W - KW + 1
DIN, DOUT = G * OW, 9
LR, WD, MOM = 0.02, 0.0002, 0.9


class OracleNet:
    """conv -> relu -> fc -> softmax with the Nnet interface DataParallelStep drives."""
    num_components = 4

    def __init__(self, seed=3):
        from oracle import oracle
        self.o = oracle
        rng = np.random.default_rng(seed)
        self.k = (rng.standard_normal((KH * KW * C, G)) * 0.1).astype(np.float32)
        self.kb = rng.standard_normal(G).astype(np.float32)
        self.kp = np.zeros_like(self.k)
        self.w = (rng.standard_normal((DOUT, DIN)) * 0.1).astype(np.float32)
        self.wb = np.full(DOUT, 0.1, np.float32)
        self.wp = np.zeros_like(self.w)
        sizes = [self.k.size + G, 0, self.w.size + DOUT, 0]
        self.off = np.concatenate([[0], np.cumsum(sizes)])
        self.arena = torch.zeros(int(self.off[-1]), dtype=torch.float32)
        self.objf = 0.0

    def forward(self, x):
        self.forward_range(x, 0, 3)

    def forward_range(self, x, first, last):
        o = self.o
        for c in range(first, last + 1):
            if c == 0:
                self.a0 = x
                self.a1 = o.conv_propagate(x, self.k, self.kb, H, W, C, 0, 0, KH, KW, G)
            elif c == 1:
                self.a2 = np.maximum(self.a1, 0)
            elif c == 2:
                self.a3 = o.fc_propagate(self.a2, self.w, self.wb)
            else:
                self.a4 = o.softmax_propagate(self.a3)

    def objf_and_deriv(self, labels):
        n = self.a4.shape[0]
        self.objf += float(np.log(self.a4[np.arange(n), labels]).sum())
        d = np.zeros_like(self.a4)
        d[np.arange(n), labels] = 1.0 / self.a4[np.arange(n), labels]
        self.d = d

    def gradient_bucket(self, c):
        return int(self.off[c]), int(self.off[c + 1] - self.off[c])

    def _put(self, c, wgrad, bgrad):
        off, _ = self.gradient_bucket(c)
        flat = np.concatenate([wgrad.ravel(), bgrad.ravel()]).astype(np.float32)
        self.arena[off:off + flat.size] = torch.from_numpy(flat)

    def backward(self, last, first):
        o = self.o
        for c in range(last, first - 1, -1):
            d = self.d
            if c == 3:
                self.d = o.softmax_backprop(self.a4, d)
            elif c == 2:
                self._put(2, d.T.astype(np.float64) @ self.a2.astype(np.float64), d.sum(0))
                self.d = o.fc_backprop(d, self.w)
            elif c == 1:
                self.d = np.where(self.a2 > 0, d, 0).astype(np.float32)
            else:
                _, _, _, g, bg = o.conv_update(self.a0, d, self.k, self.kb, self.kp, H, W, C, 0, 0, KH, KW, G,
                                               LR, WD, MOM, apply=False)
                self._put(0, g, bg)
                self.d = o.conv_backprop(d, self.k, H, W, C, 0, 0, KH, KW, G)

    def apply_gradients(self, rows):
        for c in (0, 2):
            self.apply_component_gradient(c, rows)

    def apply_component_gradient(self, comp, rows):
        lr = np.float32(LR) / np.float32(rows)
        a = self.arena.numpy()
        for c, (w, b, p) in ((0, (self.k, self.kb, self.kp)), (2, (self.w, self.wb, self.wp))):
            if c != comp:
                continue
            off, _ = self.gradient_bucket(c)
            g = a[off:off + w.size].reshape(w.shape)
            bg = a[off + w.size:off + w.size + b.size]
            p *= np.float32(MOM)
            p += np.float32(-lr * WD) * w
            p += lr * g
            w += p
            b += lr * bg


def _data(n, seed=11):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, H * W * C)).astype(np.float32), rng.integers(0, DOUT, n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _pipelined_worker(rank, world, port, n_global, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        small = dist.new_group(ranks=list(range(world)))
        b, e = shard_rows(n_global, rank, world)
        net = OracleNet()
        step = PipelinedDataParallelStep(net, net.arena, [0, 2], dist, world, late_from=2, small_group=small)
        batches = [_data(n_global, seed=11 + i) for i in range(3)]
        step.prime(batches[0][0][b:e], batches[0][1][b:e])
        for x, y in batches[1:]:
            step.rotate(x[b:e], y[b:e], n_global)
        step.finish(n_global)
        np.savez(os.path.join(out, "prank%d.npz" % rank), k=net.k, kb=net.kb, w=net.w, wb=net.wb, kp=net.kp)
    finally:
        dist.destroy_process_group()


def _worker(rank, world, port, n_global, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, y = _data(n_global)
        b, e = shard_rows(n_global, rank, world)
        net = OracleNet()
        step = DataParallelStep(net, net.arena, [0, 2], dist, world)
        for _ in range(2):
            step(x[b:e], y[b:e], n_global)
        np.savez(os.path.join(out, "rank%d.npz" % rank), k=net.k, kb=net.kb, w=net.w, wb=net.wb, kp=net.kp)
    finally:
        dist.destroy_process_group()


def test_shard_rows_partitions_exactly():
    for n in (1, 7, 256, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(4, 2, 2)


@pytest.mark.parametrize("world, n", [(2, 12), (4, 14)], ids=["world2", "world4-ragged"])
def test_world_size_2_gloo_matches_single_process(tmp_path, world, n):
    """P ranks x their row shards (14 rows over 4 ranks: 4 + 4 + 3 + 3) == one process x all rows, two steps;
    the replicas stay bit-identical."""
    x, y = _data(n)
    ref = OracleNet()
    single = DataParallelStep(ref, ref.arena, [0, 2], None, 1)
    for _ in range(2):
        single(x, y, n)
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    want = dict(k=ref.k, kb=ref.kb, w=ref.w, wb=ref.wb, kp=ref.kp)
    for name, w in want.items():
        for r in range(world):
            assert np.abs(got[r][name] - w).max() <= 1e-5 * max(np.abs(w).max(), 1e-30), (name, r)
            assert np.array_equal(got[0][name], got[r][name])      # replicas stay bit-identical


def test_world_gt_1_needs_a_process_group():
    net = OracleNet()
    with pytest.raises(ValueError):
        DataParallelStep(net, net.arena, [0, 2], None, 2)


def test_late_components_picks_the_big_buckets():
    net = OracleNet()
    assert late_components(net, [0, 2], min_floats=DIN * DOUT) == 2
    assert late_components(net, [0, 2], min_floats=1) == 0
    assert late_components(net, [0, 2], min_floats=10 ** 9) is None


def test_pipelined_step_matches_plain_step_world_2(tmp_path):
    """Three batches through the software-pipelined step on two gloo ranks == three plain
    single-process steps (every weight is updated before the forward pass that reads it)."""
    n = 12
    ref = OracleNet()
    single = DataParallelStep(ref, ref.arena, [0, 2], None, 1)
    for i in range(3):
        x, y = _data(n, seed=11 + i)
        single(x, y, n)
    mp.spawn(_pipelined_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(str(tmp_path), "prank%d.npz" % r)) for r in (0, 1)]
    want = dict(k=ref.k, kb=ref.kb, w=ref.w, wb=ref.wb, kp=ref.kp)
    for name, w in want.items():
        for r in (0, 1):
            assert np.abs(got[r][name] - w).max() <= 1e-5 * max(np.abs(w).max(), 1e-30), (name, r)
        assert np.array_equal(got[0][name], got[1][name])


def _averaging_worker(rank, world, port, out):
    from kaldi_cnn_b200.dp import ParameterAveraging
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = torch.arange(12, dtype=torch.float32).reshape(3, 4) * (rank + 1)
        pitched = torch.zeros(3, 8)
        b = pitched[:, :5]                                   # a pitched matrix view: not contiguous
        b.fill_(float(rank))
        avg = ParameterAveraging([w, b], dist, world, every=2)
        flags = []
        for _ in range(4):
            w += 1.0                                          # the "local step"
            flags.append(avg.after_step())
        np.savez(os.path.join(out, "arank%d.npz" % rank), w=w.numpy(), b=pitched.numpy(), flags=np.array(flags))
    finally:
        dist.destroy_process_group()


def test_parameter_averaging_baseline_world_2(tmp_path):
    """The comparison row of SURVEY 8d C5 (nnet-am-average emulation): local steps, mean of the
    parameters every `every` steps, pitched views handled, padding untouched."""
    mp.spawn(_averaging_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(str(tmp_path), "arank%d.npz" % r)) for r in (0, 1)]
    base = np.arange(12, dtype=np.float32).reshape(3, 4)
    # step 1, 2: +2 on each rank, then mean of (base + 2, 2 base + 2) = 1.5 base + 2; steps 3, 4 likewise
    want = 1.5 * base + 4.0
    for r in (0, 1):
        assert list(got[r]["flags"]) == [False, True, False, True]
        assert np.allclose(got[r]["w"], want)
        assert np.allclose(got[r]["b"][:, :5], 0.5) and np.all(got[r]["b"][:, 5:] == 0.0)
    assert np.array_equal(got[0]["w"], got[1]["w"])


def test_parameter_averaging_argument_checks():
    from kaldi_cnn_b200.dp import ParameterAveraging
    with pytest.raises(ValueError):
        ParameterAveraging([], None, 2, 1)
    with pytest.raises(ValueError):
        ParameterAveraging([], None, 1, 0)
    one = ParameterAveraging([torch.ones(3)], None, 1, 1)
    assert one.after_step() is True                          # world 1: nothing to exchange


class _FakePeer:
    """Stands in for dp.PeerMemoryAllReduce on CPU: same interface (arena, all_reduce(offset, length,
    channel) -> handle with wait()), the reduction done by gloo on the arena slice."""

    def __init__(self, arena):
        self.arena = arena
        self.calls = []

    def all_reduce(self, offset, length, channel=0):
        self.calls.append((int(offset), int(length), int(channel)))
        return dist.all_reduce(self.arena[offset:offset + length], async_op=True)

    def failed(self):
        self.polled = getattr(self, "polled", 0) + 1      # the step must ask at its synchronisation points
        return False


def _peer_worker(rank, world, port, n_global, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_rows(n_global, rank, world)
        net = OracleNet()
        peer = _FakePeer(net.arena)
        step = PipelinedDataParallelStep(net, net.arena, [0, 2], dist, world, late_from=2, peer=peer)
        batches = [_data(n_global, seed=11 + i) for i in range(3)]
        step.prime(batches[0][0][b:e], batches[0][1][b:e])
        for x, y in batches[1:]:
            step.rotate(x[b:e], y[b:e], n_global)
        step.finish(n_global)
        assert peer.polled >= 1
        np.savez(os.path.join(out, "qrank%d.npz" % rank), k=net.k, kb=net.kb, w=net.w, wb=net.wb, kp=net.kp,
                 calls=np.array(peer.calls))
    finally:
        dist.destroy_process_group()


def test_pipelined_step_through_the_peer_memory_interface_world_2(tmp_path):
    """The host glue of the NVLink path (PipelinedDataParallelStep(peer=...)): buckets are reduced
    through peer.all_reduce(offset, length, channel) -- FC buckets on channel 0, convolution buckets
    on channel 1 -- and the result is the plain step's."""
    n = 12
    ref = OracleNet()
    single = DataParallelStep(ref, ref.arena, [0, 2], None, 1)
    for i in range(3):
        x, y = _data(n, seed=11 + i)
        single(x, y, n)
    mp.spawn(_peer_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(str(tmp_path), "qrank%d.npz" % r)) for r in (0, 1)]
    for name, w in dict(k=ref.k, kb=ref.kb, w=ref.w, wb=ref.wb, kp=ref.kp).items():
        for r in (0, 1):
            assert np.abs(got[r][name] - w).max() <= 1e-5 * max(np.abs(w).max(), 1e-30), (name, r)
        assert np.array_equal(got[0][name], got[1][name])
    fc_off, fc_len = ref.gradient_bucket(2)
    cv_off, cv_len = ref.gradient_bucket(0)
    calls = [tuple(c) for c in got[0]["calls"]]
    assert calls == [(fc_off, fc_len, 0), (cv_off, cv_len, 1)] * 3          # issue order: top layer first, every batch
