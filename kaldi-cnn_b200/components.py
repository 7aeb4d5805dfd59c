"""Python views of the component-level C ABI (include/kcnn_capi.h), used by the parity
tests and bench.py.  Each method forwards to one C call; nothing is computed here.

    comp = Component.from_string("ConvolutionComponent in-height=40 ...")
    y = comp.propagate(x)                      # x, y: 2-D CUDA tensors (rows contiguous)
    dx = comp.backprop(x, y, dy, update=True)  # ordinary SGD: to_update == the component
"""
import ctypes

from . import capi

_inited = False


class KcnnError(RuntimeError):
    pass


def _lib():
    global _inited
    L = capi.load()
    if not _inited:
        import torch
        if not torch.cuda.is_available():
            raise KcnnError("kaldi-cnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        torch.cuda.current_device()        # make sure the primary context exists
        if L.kcnn_select_gpu(b"yes") != 0:
            raise KcnnError(L.kcnn_last_error().decode())
        _inited = True
    return L


def _check(rc):
    if rc != 0:
        raise KcnnError(capi.load().kcnn_last_error().decode())


def use_current_stream():
    """Point the library at torch's current CUDA stream."""
    import torch
    _lib().kcnn_set_compute_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))


def bytes_pinned_by_graphs():
    """Bytes of the device cache kept out of circulation because a recorded CUDA graph may address them."""
    return int(_lib().kcnn_device_bytes_pinned_by_graphs())


def set_math_mode(mode):
    _lib().kcnn_set_math_mode(int(mode))


def set_rand_seed(seed):
    _lib().kcnn_set_rand_seed(int(seed))


def _mat(t):
    if t is None:
        return (ctypes.c_void_p(0), 0, 0, 0)
    d = capi.mdim(t)
    return (ctypes.c_void_p(t.data_ptr()), d.rows, d.cols, d.stride)


class _DevArray:
    """Minimal __cuda_array_interface__ carrier so torch can view library-owned memory."""

    def __init__(self, ptr, rows, cols, stride):
        self.__cuda_array_interface__ = {
            "shape": (rows, cols), "typestr": "<f4", "data": (ptr, False), "version": 2,
            "strides": (stride * 4, 4),
        }


def tensor_view(ptr, rows, cols, stride):
    import torch
    if rows == 0 or cols == 0:
        return torch.empty((rows, cols), dtype=torch.float32, device="cuda")
    return torch.as_tensor(_DevArray(ptr, rows, cols, stride), device="cuda")


def _out4():
    return ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()


class Component:
    def __init__(self, handle, owned=True):
        if not handle:
            raise KcnnError(capi.load().kcnn_last_error().decode())
        self.h = ctypes.c_void_p(handle)
        self.owned = owned

    def __del__(self):
        try:
            if self.owned and self.h:
                capi.load().kcnn_component_delete(self.h)
        except Exception:
            pass

    @classmethod
    def from_string(cls, line):
        return cls(_lib().kcnn_component_new_from_string(line.encode()))

    @classmethod
    def read(cls, data, binary=True):
        return cls(_lib().kcnn_component_read(data, len(data), int(binary)))

    def write(self, binary=True):
        L = _lib()
        buf, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(L.kcnn_component_write(self.h, int(binary), ctypes.byref(buf), ctypes.byref(n)))
        out = ctypes.string_at(buf, n.value)
        L.kcnn_free(buf)
        return out

    def copy(self):
        return Component(_lib().kcnn_component_copy(self.h))

    @property
    def type(self):
        return _lib().kcnn_component_type(self.h).decode()

    def info(self):
        buf = ctypes.create_string_buffer(4096)
        _lib().kcnn_component_info(self.h, buf, 4096)
        return buf.value.decode()

    @property
    def input_dim(self):
        return _lib().kcnn_component_input_dim(self.h)

    @property
    def output_dim(self):
        return _lib().kcnn_component_output_dim(self.h)

    def propagate(self, x, out=None, num_chunks=None):
        import torch
        if out is None:
            out = torch.empty((x.shape[0], self.output_dim), dtype=torch.float32, device=x.device)
        use_current_stream()
        nc = x.shape[0] if num_chunks is None else num_chunks
        _check(_lib().kcnn_component_propagate(self.h, nc, *_mat(x), *_mat(out)))
        return out

    def backprop(self, in_value, out_value, out_deriv, update=True, to_update=None, in_deriv=None,
                 num_chunks=None):
        import torch
        if in_deriv is None:
            in_deriv = torch.empty((out_deriv.shape[0], self.input_dim), dtype=torch.float32,
                                   device=out_deriv.device)
        use_current_stream()
        tu = to_update.h if to_update is not None else (self.h if update else ctypes.c_void_p(0))
        nc = out_deriv.shape[0] if num_chunks is None else num_chunks
        iv, ov, od, idv = _mat(in_value), _mat(out_value), _mat(out_deriv), _mat(in_deriv)
        _check(_lib().kcnn_component_backprop(self.h, nc, iv[0], iv[3], ov[0], ov[3], od[0], od[1], od[3],
                                              tu, idv[0], idv[3]))
        return in_deriv

    def params(self, which):
        """0: linear_params_, 1: bias_params_ (1 x dim), 2: prev_grad_ -- live device views."""
        p, r, c, s = _out4()
        _check(_lib().kcnn_component_params(self.h, which, ctypes.byref(p), ctypes.byref(r), ctypes.byref(c),
                                            ctypes.byref(s)))
        return tensor_view(p.value, r.value, c.value, s.value)

    def gradient(self, which):
        p, r, c, s = _out4()
        _check(_lib().kcnn_component_gradient(self.h, which, ctypes.byref(p), ctypes.byref(r), ctypes.byref(c),
                                              ctypes.byref(s)))
        return tensor_view(p.value, r.value, c.value, s.value)

    def set_learning_rate(self, lr):
        _check(_lib().kcnn_component_set_learning_rate(self.h, lr))

    def set_weight_decay_momentum(self, wd, mom):
        _check(_lib().kcnn_component_set_weight_decay_momentum(self.h, wd, mom))

    def weight_decay_momentum(self):
        a, b = ctypes.c_float(), ctypes.c_float()
        _check(_lib().kcnn_component_get_weight_decay_momentum(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def set_index_routing(self, on):
        _check(_lib().kcnn_component_set_index_routing(self.h, int(on)))

    def set_deferred_update(self, on):
        _check(_lib().kcnn_component_set_deferred_update(self.h, int(on)))

    def apply_gradient(self, total_rows):
        use_current_stream()
        _check(_lib().kcnn_component_apply_gradient(self.h, int(total_rows)))


class Nnet:
    """A component sequence plus the minibatch step (kcnn_nnet_* in include/kcnn_capi.h)."""

    def __init__(self, handle):
        if not handle:
            raise KcnnError(capi.load().kcnn_last_error().decode())
        self.h = ctypes.c_void_p(handle)
        self._arena = None

    def __del__(self):
        try:
            if self.h:
                capi.load().kcnn_nnet_delete(self.h)
        except Exception:
            pass

    @classmethod
    def from_config(cls, text, skip_splice=True):
        return cls(_lib().kcnn_nnet_new_from_config(text.encode(), int(skip_splice)))

    @classmethod
    def read(cls, data, binary=True):
        return cls(_lib().kcnn_nnet_read(data, len(data), int(binary)))

    def write(self, binary=True):
        L = _lib()
        buf, n = ctypes.c_void_p(), ctypes.c_size_t()
        _check(L.kcnn_nnet_write(self.h, int(binary), ctypes.byref(buf), ctypes.byref(n)))
        out = ctypes.string_at(buf, n.value)
        L.kcnn_free(buf)
        return out

    @property
    def num_components(self):
        return _lib().kcnn_nnet_num_components(self.h)

    def component(self, i):
        return Component(_lib().kcnn_nnet_component(self.h, i), owned=False)

    @property
    def input_dim(self):
        return _lib().kcnn_nnet_input_dim(self.h)

    @property
    def output_dim(self):
        return _lib().kcnn_nnet_output_dim(self.h)

    def forward(self, feats):
        use_current_stream()
        d = capi.mdim(feats)
        _check(_lib().kcnn_nnet_forward(self.h, ctypes.c_void_p(feats.data_ptr()), d.rows, d.stride))

    def forward_range(self, feats, first, last):
        """Propagate through components [first, last] (first == 0 binds feats as the input)."""
        use_current_stream()
        d = capi.mdim(feats)
        _check(_lib().kcnn_nnet_forward_range(self.h, ctypes.c_void_p(feats.data_ptr()), d.rows, d.stride,
                                              int(first), int(last)))

    def objf_and_deriv(self, labels):
        _check(_lib().kcnn_nnet_objf_and_deriv(self.h, ctypes.c_void_p(labels.data_ptr())))

    def backward(self, last=-1, first=0):
        _check(_lib().kcnn_nnet_backward(self.h, last, first))

    def activation(self, i):
        p, r, c, s = _out4()
        _check(_lib().kcnn_nnet_activation(self.h, i, ctypes.byref(p), ctypes.byref(r), ctypes.byref(c),
                                           ctypes.byref(s)))
        return tensor_view(p.value, r.value, c.value, s.value)

    def input_deriv(self):
        p, r, c, s = _out4()
        _check(_lib().kcnn_nnet_input_deriv(self.h, ctypes.byref(p), ctypes.byref(r), ctypes.byref(c),
                                            ctypes.byref(s)))
        return tensor_view(p.value, r.value, c.value, s.value)

    def objf_and_reset(self):
        return _lib().kcnn_nnet_objf_and_reset(self.h)

    def train_step(self, feats, labels):
        """forward + objective + backward (update inside Backprop, as nnet2 does)."""
        self.forward(feats)
        self.objf_and_deriv(labels)
        self.backward()

    def train_minibatch_host(self, feats_np, labels_np):
        use_current_stream()
        objf = ctypes.c_double()
        _check(_lib().kcnn_nnet_train_minibatch_host(
            self.h, feats_np.ctypes.data_as(ctypes.c_void_p), labels_np.ctypes.data_as(ctypes.c_void_p),
            labels_np.shape[0], ctypes.byref(objf)))
        return objf.value

    def train_minibatch_host_async(self, feats_np, labels_np):
        """Pipelined host step: stages the buffers, enqueues copy + step, returns without waiting.
        Read the objective with objf_and_reset() (waits) or running_objf (does not)."""
        use_current_stream()
        _check(_lib().kcnn_nnet_train_minibatch_host_async(
            self.h, feats_np.ctypes.data_as(ctypes.c_void_p), labels_np.ctypes.data_as(ctypes.c_void_p),
            labels_np.shape[0]))

    @property
    def running_objf(self):
        return _lib().kcnn_nnet_running_objf(self.h)

    def train_step_graph(self, feats, labels):
        """The same step through NnetMinibatchUpdater::TrainStep: recorded into a CUDA graph by the
        library on the second call with the same buffers, replayed afterwards."""
        use_current_stream()
        d = capi.mdim(feats)
        _check(_lib().kcnn_nnet_train_step(self.h, ctypes.c_void_p(feats.data_ptr()), d.rows, d.stride,
                                           ctypes.c_void_p(labels.data_ptr())))

    def set_fusion(self, on):
        """[Convolution | FullyConnected] + ReLU as one launch in the forward pass (default on)."""
        _check(_lib().kcnn_nnet_set_fusion(self.h, int(bool(on))))

    def set_graphs(self, on):
        """CUDA-graph recording of train_step_graph / train_minibatch_host* (default on)."""
        _check(_lib().kcnn_nnet_set_graphs(self.h, int(bool(on))))

    @property
    def last_step_replayed(self):
        return bool(_lib().kcnn_nnet_last_step_replayed(self.h))

    @property
    def frames_per_example(self):
        """Input rows per training example (the Splice front end's context span; 1 without one)."""
        return int(_lib().kcnn_nnet_frames_per_example(self.h))

    @property
    def fused_active(self):
        """True when the step runs as the fused plan (csrc/nnet2/nnet-fused.cc); valid after a forward."""
        return bool(_lib().kcnn_nnet_fused_active(self.h))

    # ---- data parallel -----------------------------------------------------------
    def gradient_floats(self):
        return int(_lib().kcnn_nnet_gradient_floats(self.h))

    def enable_data_parallel(self, arena=None):
        """Deferred updates + one flat gradient arena (a torch tensor, so torch.distributed
        can all-reduce it); returns the arena.  arena: caller-provided storage of at least
        gradient_floats() floats (e.g. NVLink peer memory, dp.PeerMemoryAllReduce.arena)."""
        import torch
        L = _lib()
        _check(L.kcnn_nnet_set_deferred_update(self.h, 1))
        n = L.kcnn_nnet_gradient_floats(self.h)
        if arena is not None:
            if arena.numel() < n or arena.dtype != torch.float32 or not arena.is_cuda:
                raise ValueError("arena must be a CUDA float32 tensor of >= %d elements" % n)
            self._arena = arena
        else:
            self._arena = torch.zeros(n, dtype=torch.float32, device="cuda")
        _check(L.kcnn_nnet_set_gradient_arena(self.h, ctypes.c_void_p(self._arena.data_ptr())))
        return self._arena

    def gradient_bucket(self, component):
        off, ln = ctypes.c_size_t(), ctypes.c_size_t()
        _check(_lib().kcnn_nnet_gradient_bucket(self.h, component, ctypes.byref(off), ctypes.byref(ln)))
        return off.value, ln.value

    def apply_gradients(self, total_rows):
        _check(_lib().kcnn_nnet_apply_gradients(self.h, int(total_rows)))

    def apply_component_gradient(self, component, total_rows):
        """The deferred SGD step of ONE updatable component (dp.py pipelines these under the
        all-reduces of the layers below)."""
        self.component(component).apply_gradient(total_rows)
