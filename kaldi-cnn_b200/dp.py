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
import os


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


class NativeDataParallel:
    """The data-parallel trainer of the library itself (csrc/nnet2/nnet-dp.cc, C ABI kcnn_nnet_dp_*):
    schedule, streams, graph capture and the per-layer fused reduce + SGD + broadcast kernel are C++ /
    CUDA; this class only performs the rendezvous a host must do -- allocate the symmetric arena, send
    its handle to the peers, map theirs -- and forwards calls.

    Arena exchange: CUDA IPC handles (kcnn_ipc_alloc / kcnn_ipc_open) carried by
    dist.all_gather_object, i.e. no framework-private API.  multicast=True asks for the in-switch (NVLS)
    form of the kernel, which needs a multicast mapping of the arena; the only provider in this image is
    torch's symmetric memory, so that variant allocates through it (and raises when it is unavailable)."""

    def __init__(self, net, dist, multicast=False):
        import ctypes
        import torch
        from . import capi
        self.lib = L = capi.lib()
        self.net, self.dist = net, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        floats = int(L.kcnn_nnet_dp_arena_floats(net.h))
        if floats == 0:
            raise RuntimeError(L.kcnn_last_error().decode())
        self._ipc_local, self._ipc_peers, self._symm = None, [], None
        mc = 0
        if multicast:
            import torch.distributed._symmetric_memory as symm
            group = dist.group.WORLD
            try:
                symm.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
            self._symm = symm.empty(floats, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
            hdl = symm.rendezvous(self._symm, group)
            self._symm.zero_()
            mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
            if mc == 0:
                raise RuntimeError("symmetric memory has no multicast mapping on this platform")
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            local = self._symm.data_ptr()
            self._hdl = hdl
        else:
            ptr = ctypes.c_void_p()
            handle = (ctypes.c_ubyte * 64)()
            if L.kcnn_ipc_alloc(ctypes.c_size_t(floats * 4), ctypes.byref(ptr), handle) != 0:
                raise RuntimeError("kcnn_ipc_alloc failed")
            self._ipc_local = ptr.value
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle))
            ptrs = []
            for r, hb in enumerate(handles):
                if r == self.rank:
                    ptrs.append(ptr.value)
                    continue
                q = ctypes.c_void_p()
                buf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
                if L.kcnn_ipc_open(buf, ctypes.byref(q)) != 0:
                    raise RuntimeError("kcnn_ipc_open failed for rank %d (no peer access?)" % r)
                self._ipc_peers.append(q.value)
                ptrs.append(q.value)
            local = ptr.value
        torch.cuda.synchronize()
        dist.barrier()
        self.multicast_base = mc
        self.bases = (ctypes.c_ulonglong * self.world)(*ptrs)
        self.h = ctypes.c_void_p(L.kcnn_nnet_dp_create(net.h, self.rank, self.world, ctypes.c_void_p(local), self.bases,
                                                       ctypes.c_ulonglong(mc)))
        if not self.h:
            raise RuntimeError(L.kcnn_last_error().decode())
        torch.cuda.synchronize()
        dist.barrier()

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.kcnn_last_error().decode())

    def prime(self, feats, labels):
        import ctypes
        from . import capi
        d = capi.mdim(feats)
        self._check(self.lib.kcnn_nnet_dp_prime(self.h, ctypes.c_void_p(feats.data_ptr()), d.rows, d.stride,
                                                ctypes.c_void_p(labels.data_ptr())))

    def _raise_if_failed(self, synchronise):
        """A peer barrier timed out on this rank (a peer is missing or stalled for longer than
        KCNN_P2P_TIMEOUT_MS): the update of that step was skipped, the replicas may have diverged.  The
        asynchronous form reads the error words the trainer copies to pinned memory with every rotation --
        no synchronisation, so it reports a failure one or two steps late but every step."""
        if self.failed(synchronise):
            raise RuntimeError("kaldi-cnn_b200 data parallel: a peer-memory barrier timed out on rank %d; "
                               "the step's update was skipped and the replicas may have diverged" % self.dist.get_rank())

    def rotate(self, feats_next, labels_next, rows_global):
        import ctypes
        from . import capi
        d = capi.mdim(feats_next)
        self._check(self.lib.kcnn_nnet_dp_rotate(self.h, ctypes.c_void_p(feats_next.data_ptr()), d.rows, d.stride,
                                                 ctypes.c_void_p(labels_next.data_ptr()), int(rows_global)))
        self._raise_if_failed(False)

    def finish(self, rows_global):
        self._check(self.lib.kcnn_nnet_dp_finish(self.h, int(rows_global)))
        self._raise_if_failed(True)

    def train_minibatch_host_async(self, feats_np, labels_np, rows_global):
        import ctypes
        self._check(self.lib.kcnn_nnet_dp_train_minibatch_host_async(
            self.h, feats_np.ctypes.data_as(ctypes.c_void_p), labels_np.ctypes.data_as(ctypes.c_void_p),
            labels_np.shape[0], int(rows_global)))
        self._raise_if_failed(False)

    def failed(self, synchronise=True):
        return bool(self.lib.kcnn_nnet_dp_failed(self.h, int(synchronise)))

    def gather_momentum(self):
        self._check(self.lib.kcnn_nnet_dp_gather_momentum(self.h))

    @property
    def last_rotate_replayed(self):
        return bool(self.lib.kcnn_nnet_dp_last_rotate_replayed(self.h))

    def close(self):
        """Parameters move back into the components; the arena is released."""
        import torch
        torch.cuda.synchronize()
        self.dist.barrier()
        if self.h:
            self.lib.kcnn_nnet_dp_delete(self.h)
            self.h = None
        torch.cuda.synchronize()
        self.dist.barrier()
        import ctypes
        for q in self._ipc_peers:
            self.lib.kcnn_ipc_close(ctypes.c_void_p(q))
        self._ipc_peers = []
        self.dist.barrier()
        if self._ipc_local:
            self.lib.kcnn_ipc_free(ctypes.c_void_p(self._ipc_local))
            self._ipc_local = None


class _Done:
    """What dist.all_reduce(async_op=True) returns, for the peer-memory path: wait() makes the
    CURRENT stream wait for the reduction (an event wait, capturable into a CUDA graph)."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        import torch
        torch.cuda.current_stream().wait_event(self.event)


class PeerMemoryAllReduce:
    """Gradient arena in NVLink peer memory + the library's own all-reduce kernel
    (csrc/cnslmat/kernels_p2p.cu): kcnn_p2p_allreduce_f32 -- two-shot, 128-bit peer loads and
    stores -- or, with multicast=True, kcnn_p2p_allreduce_multicast_f32 -- the sum formed inside the
    NVSwitch (multimem.ld_reduce / multimem.st); device-side flags, no NCCL on the data path.  The arena is one symmetric
    allocation (torch.distributed._symmetric_memory: cuMem handles exchanged through the
    process group) of `floats` gradient floats followed by the flag words.

    all_reduce(offset, length, channel) enqueues the reduction of arena[offset:offset+length]
    on the channel's own stream behind everything the current stream has enqueued so far and
    returns a handle whose wait() orders the current stream behind it -- the contract of
    dist.all_reduce(async_op=True).  Two channels = two independent streams / flag sets (small
    convolution buckets must not queue behind 140 MB of FC gradients)."""

    def __init__(self, lib, dist, floats, multicast=False):
        """multicast: reduce inside the NVSwitch (kcnn_p2p_allreduce_multicast_f32, multimem.ld_reduce /
        multimem.st on the allocation's multicast mapping); raises RuntimeError when the
        platform gives the allocation no multicast address."""
        import ctypes
        import torch
        import torch.distributed._symmetric_memory as symm
        self.lib, self.dist = lib, dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        flag = int(lib.kcnn_p2p_flag_floats())
        self.floats = (int(floats) + 63) // 64 * 64
        self.flag_off = self.floats
        group = dist.group.WORLD
        try:
            symm.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass                                   # newer torch: not needed
        self.buf = symm.empty(self.floats + flag, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(ptrs) == self.world and ptrs[self.rank] == self.buf.data_ptr()
        self.bases = (ctypes.c_ulonglong * self.world)(*ptrs)
        self.arena = self.buf[:self.floats]
        self.multicast_base = 0
        if multicast:
            self.multicast_base = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
            if self.multicast_base == 0:
                raise RuntimeError("symmetric memory has no multicast mapping on this platform")
        prio = int(os.environ.get("KCNN_P2P_STREAM_PRIORITY", "-1"))      # reductions first: they are short CTAs
        self.streams = [torch.cuda.Stream(priority=prio), torch.cuda.Stream(priority=prio)]

    def all_reduce(self, offset, length, channel=0):
        import ctypes
        import torch
        if length % 4 or offset % 4:
            raise ValueError("offset / length must be multiples of 4 floats")
        cur = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(cur)
        st = self.streams[channel]
        st.wait_event(ready)
        if self.multicast_base:
            rc = self.lib.kcnn_p2p_allreduce_multicast_f32(
                ctypes.c_void_p(st.cuda_stream), self.bases, ctypes.c_ulonglong(self.multicast_base), self.rank,
                self.world, int(offset), int(length), int(self.flag_off), int(channel))
        else:
            rc = self.lib.kcnn_p2p_allreduce_f32(ctypes.c_void_p(st.cuda_stream), self.bases, self.rank, self.world,
                                                 int(offset), int(length), int(self.flag_off), int(channel))
        if rc != 0:
            raise RuntimeError("kcnn_p2p_allreduce_f32 rejected its arguments")
        done = torch.cuda.Event()
        done.record(st)
        return _Done(done)

    def failed(self):
        """True when a barrier gave up waiting for a peer (synchronises)."""
        import ctypes
        return bool(self.lib.kcnn_p2p_error(ctypes.c_void_p(self.buf.data_ptr()), int(self.flag_off)))


class ParameterAveraging:
    """The baseline the gradient all-reduce replaces, for the comparison row of SURVEY 8d C5: the
    reference's recipe runs independent training jobs and averages their models through the
    filesystem (nnet-am-average, egs/steps/nnet0/train_conv_dropout.sh:323-341).  Emulated in
    memory: every rank takes ordinary local steps; after every `every`-th step the given parameter
    tensors (linear_params_ and bias_params_, what Nnet::Scale / AddNnet touch) are replaced by
    their mean over the ranks.  Momentum (prev_grad_) stays local, as in the recipe."""

    def __init__(self, tensors, dist, world, every):
        if every < 1:
            raise ValueError("every must be >= 1")
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs a torch.distributed module / process group")
        self.tensors, self.dist, self.world, self.every = list(tensors), dist, world, every
        self.steps = 0

    def after_step(self):
        """Call once per local training step; True when this call averaged."""
        self.steps += 1
        if self.steps % self.every:
            return False
        self.average()
        return True

    def average(self):
        if self.world <= 1:
            return
        for t in self.tensors:
            flat = t if t.is_contiguous() else t.contiguous()      # pitched matrices: reduce a packed copy
            self.dist.all_reduce(flat)
            flat.mul_(1.0 / self.world)
            if flat is not t:
                t.copy_(flat)


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

    def __init__(self, net, arena, updatable, dist, world, late_from, small_group=None, skip_reduce=False,
                 peer=None):
        """peer: a PeerMemoryAllReduce whose .arena IS `arena` -- the reductions then run on the
        library's own NVLink kernel instead of NCCL."""
        if world > 1 and dist is None:
            raise ValueError("world > 1 needs a torch.distributed module / process group")
        self.peer = peer
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
        return self._reduce_impl(c, off, ln)

    def _reduce_impl(self, c, off, ln):
        if self.peer is not None:
            return self.peer.all_reduce(off, ln, channel=1 if c < self.late_from else 0)
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
        # a barrier of the peer-memory all-reduce that gave up leaves an error mark: check it at
        # every synchronisation point of the pipeline
        if self.peer is not None and self.peer.failed():
            raise RuntimeError("kaldi-cnn_b200 data parallel: a peer-memory barrier timed out; the reduced "
                               "gradients of that step are not valid")
