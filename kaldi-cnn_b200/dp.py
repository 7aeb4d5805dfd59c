"""Data-parallel training step over minibatch rows.

Replaces the reference's file-based multi-job averaging (nnet-am-average once per iteration,
egs/steps/nnet0/train_conv_dropout.sh:323-341) by a per-step gradient all-reduce:

  * rank r owns rows [r*N/P, (r+1)*N/P) of the global minibatch (`shard_rows`)
  * every updatable component runs Backprop with the update DEFERRED and leaves its
    un-normalised dW, db in one bucket of a gradient arena
  * buckets are all-reduced (sum) top layer first, each as soon as its layer's backward has
    been issued, so the transfer overlaps the rest of the backward
  * every rank then applies the identical momentum / weight-decay step with
    lr = learning_rate / N_global (the reference normalises by the minibatch rows,
    nnet0/nnet-component-nnet0.cc:767, 1136) -- P ranks x N/P rows reproduce the 1-rank N-row
    step up to floating-point summation order.

Host logic only: `net` is anything with the Nnet interface of components.py (the CUDA model on
a GPU box, an oracle-backed model in the world-size-2 gloo test).
"""


def shard_rows(num_rows, rank, world):
    """[begin, end) of the global minibatch rows rank `rank` of `world` processes."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(num_rows, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def bucket_order(updatable):
    """Components in the order their gradients become available (top layer first)."""
    return sorted(updatable, reverse=True)


def late_components(net, updatable, min_floats=1 << 20):
    """Index of the first updatable component (in network order) from which on every updatable
    component has a gradient bucket of at least `min_floats` -- the fully connected stack, 96 % of
    the gradient bytes of the reference model -- or None."""
    late = None
    for c in sorted(updatable, reverse=True):
        if net.gradient_bucket(c)[1] >= min_floats:
            late = c
        else:
            break
    return late


class DataParallelStep:
    def __init__(self, net, arena, updatable, dist=None, world=1, skip_reduce=False):
        """skip_reduce: diagnosis only (bench.py --dp-skip-reduce): the deferred-update step without
        the all-reduces, to separate their cost from the cost of splitting Update."""
        self.net, self.arena, self.dist, self.world = net, arena, dist, world
        self.skip_reduce = skip_reduce
        self.updatable = bucket_order(updatable)
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs a torch.distributed module / process group")

    def __call__(self, feats, labels, rows_global):
        net = self.net
        net.forward(feats)
        net.objf_and_deriv(labels)
        works, hi = [], net.num_components - 1
        for c in self.updatable:
            net.backward(hi, c)                       # layers hi .. c, update deferred
            off, ln = net.gradient_bucket(c)
            if self.world > 1 and not self.skip_reduce:
                works.append(self.dist.all_reduce(self.arena[off:off + ln], async_op=True))
            hi = c - 1
        if hi >= 0:
            net.backward(hi, 0)
        if not works:
            net.apply_gradients(rows_global)
            return
        # Pipelined apply: the reductions complete in issue order, so layer c's SGD step runs as
        # soon as ITS bucket has arrived, under the reductions of the layers below it.
        for c, w in zip(self.updatable, works):
            w.wait()
            net.apply_component_gradient(c, rows_global)


class PipelinedDataParallelStep:
    """The same step, software-pipelined across the batch boundary so that the big all-reduces
    get the whole convolution backward AND the next batch's convolution forward to hide under:

        prime(x0, y0):      forward(x0), objective / derivative
        rotate(x1, y1, N):  backward(batch 0) with the all-reduces issued as each layer finishes
                            small buckets (convolutions): wait, apply         [second communicator]
                            forward(x1) through the layers BELOW the first late component
                            late buckets (FC stack): wait, apply
                            forward(x1) through the rest, objective / derivative of batch 1

    Every weight is updated before the first forward pass that reads it, so the numbers are
    those of the plain step; one rotate() is exactly one backward + update + forward.  The
    small buckets go through their own communicator (`dist_small_group`): NCCL runs the
    collectives of one communicator in issue order, and the convolution gradients -- produced
    last -- must not queue behind 140 MB of FC gradients."""

    def __init__(self, net, arena, updatable, dist, world, late_from, small_group=None, skip_reduce=False):
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs a torch.distributed module / process group")
        self.net, self.arena, self.dist, self.world = net, arena, dist, world
        self.updatable = bucket_order(updatable)
        self.late_from = late_from if late_from is not None else net.num_components
        self.small_group = small_group
        self.skip_reduce = skip_reduce
        self.primed = False

    def prime(self, feats, labels):
        self.net.forward(feats)
        self.net.objf_and_deriv(labels)
        self.primed = True

    def _reduce(self, c):
        off, ln = self.net.gradient_bucket(c)
        if self.world <= 1 or self.skip_reduce:
            return None
        group = self.small_group if c < self.late_from else None
        return self.dist.all_reduce(self.arena[off:off + ln], group=group, async_op=True)

    def rotate(self, feats_next, labels_next, rows_global):
        if not self.primed:
            raise RuntimeError("prime() the pipeline with the first batch")
        net = self.net
        works, hi = [], net.num_components - 1
        for c in self.updatable:
            net.backward(hi, c)
            works.append((c, self._reduce(c)))
            hi = c - 1
        if hi >= 0:
            net.backward(hi, 0)
        for c, w in works:                            # convolutions: small, own communicator
            if c < self.late_from:
                if w is not None:
                    w.wait()
                net.apply_component_gradient(c, rows_global)
        split = min(self.late_from, net.num_components)
        if split > 0:
            net.forward_range(feats_next, 0, split - 1)
        for c, w in works:                            # FC stack, in issue order
            if c >= self.late_from:
                if w is not None:
                    w.wait()
                net.apply_component_gradient(c, rows_global)
        if split < net.num_components:
            net.forward_range(feats_next, split, net.num_components - 1)
        net.objf_and_deriv(labels_next)

    def finish(self, rows_global):
        """Backward + update of the batch still in the pipeline (no further forward)."""
        net = self.net
        works, hi = [], net.num_components - 1
        for c in self.updatable:
            net.backward(hi, c)
            works.append((c, self._reduce(c)))
            hi = c - 1
        if hi >= 0:
            net.backward(hi, 0)
        for c, w in works:
            if w is not None:
                w.wait()
            net.apply_component_gradient(c, rows_global)
        self.primed = False
