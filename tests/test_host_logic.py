"""Host logic of the C++ layer through the C ABI, on a machine WITHOUT a GPU (the `-m "not gpu"` suite).

What can be checked without launching anything:
  * configuration errors of the three nnet0 components are raised where the reference raises them, with its
    messages (nnet0/nnet-component-nnet0.cc:379-383 "Could not process these elements" / "Bad initializer",
    :861-863 "Invalid initializer for layer of type" for the max-pool, :1092-1093 for the FC layer -- whose
    left-over-key check comes after Init() as in the reference (:1094-1098), i.e. behind the device), and the factory rejects unknown types (nnet2/nnet-component.cc:97-105);
  * MaxpoolComponent has no parameters, so its whole host side runs here: InitFromString, the shape asserts
    of :783-812, Info, Copy, Type, dims, BackpropNeedsInput / Output (nnet0/nnet-component-nnet0.h:172-173),
    and the Read / Write token streams against the byte-level fixtures of tests/golden/mdl (Write order of
    :936-959, the pre-overlap stream of :917-934);
  * everything that needs device memory FAILS LOUDLY instead of computing on the CPU (there is no CPU path);
  * argument validation of the peer-memory kernels' launchers (bad rank / world / alignment -> -1, no launch).
Nothing here touches oracle/ and nothing computes."""
import ctypes
import os

import pytest

from kaldi_cnn_b200 import capi

MDL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mdl")
POOL = ("MaxpoolComponent in-height=1 in-width=8 in-channel=128 pool-height-dim=1 pool-width-dim=2 "
        "pool-channel-dim=2")
CONV = ("ConvolutionComponent in-height=40 in-width=11 in-channel=3 kernel-height=40 kernel-width=4 stride=1 "
        "group=128 out-height=1 out-width=8 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5")
FC = "FullyConnectedComponent input-dim=256 output-dim=1024 learning-rate=0.02 param-stddev=0.01 bias-stddev=0.5"


def blob(name):
    return open(os.path.join(MDL, name), "rb").read()


@pytest.fixture(scope="module")
def L():
    import torch
    if torch.cuda.is_available():
        pytest.skip("host-only checks: they assert the behaviour of a machine without a CUDA device")
    if not os.path.exists(capi.LIB_PATH):                # a fresh checkout: build first (nvcc cross-compiles)
        import __graft_entry__
        __graft_entry__.build()
    lib = capi.load()
    assert lib.kcnn_select_gpu(b"optional") == 1       # 1 = no device, library stays disabled (no CPU path)
    return lib


def err(L):
    return L.kcnn_last_error().decode()


def new(L, line):
    return L.kcnn_component_new_from_string(line.encode())


def write(L, h, binary):
    buf, n = ctypes.c_void_p(), ctypes.c_size_t()
    assert L.kcnn_component_write(ctypes.c_void_p(h), int(binary), ctypes.byref(buf), ctypes.byref(n)) == 0, err(L)
    out = ctypes.string_at(buf, n.value)
    L.kcnn_free(buf)
    return out


def test_select_gpu_yes_fails_without_a_device(L):
    assert L.kcnn_select_gpu(b"yes") == -1
    assert "No CUDA device" in err(L)
    assert L.kcnn_select_gpu(b"no") == 1


@pytest.mark.parametrize("line, message", [
    ("ConvolutionComponent in-height=40", "Bad initializer"),                               # :379-380
    (CONV + " bogus=1", "Could not process these elements in initializer: bogus=1"),       # :382-383
    ("FullyConnectedComponent input-dim=256", "Bad initializer"),                           # :1092-1093
    ("MaxpoolComponent in-height=1 in-width=8", "Invalid initializer for layer of type MaxpoolComponent"),   # :861-863
    (POOL + " stride=3", "Invalid initializer for layer of type MaxpoolComponent"),         # left-over key, same branch
    ("NoSuchComponent a=1", "no such type of Component"),
    ("", "Bad initializer line"),
])
def test_config_errors_are_the_references(L, line, message):
    assert not new(L, line)
    assert message in err(L), err(L)


@pytest.mark.parametrize("line", [
    POOL.replace("pool-width-dim=2", "pool-width-dim=3"),       # 8 % 3 != 0            (:793-795)
    POOL.replace("pool-channel-dim=2", "pool-channel-dim=3"),   # 128 % 3 != 0
    POOL.replace("in-width=8", "in-width=0"),
])
def test_maxpool_shape_asserts(L, line):
    """Geometry the reference rejects (the asserts of Init :791-810, or output-dim <= 0 at :861) is rejected
    here -- by the same assert or by the "Invalid initializer" branch, never accepted."""
    assert not new(L, line)
    assert "KALDI_ASSERT" in err(L) or "Invalid initializer for layer of type MaxpoolComponent" in err(L), err(L)


def test_maxpool_component_host_side(L):
    h = new(L, POOL)
    assert h, err(L)
    try:
        assert L.kcnn_component_type(ctypes.c_void_p(h)) == b"MaxpoolComponent"
        assert L.kcnn_component_input_dim(ctypes.c_void_p(h)) == 1 * 8 * 128
        assert L.kcnn_component_output_dim(ctypes.c_void_p(h)) == 1 * 4 * 64
        # BackpropNeedsInput / BackpropNeedsOutput: true / true (nnet0/nnet-component-nnet0.h:172-173)
        assert L.kcnn_component_backprop_needs_input(ctypes.c_void_p(h)) == 1
        assert L.kcnn_component_backprop_needs_output(ctypes.c_void_p(h)) == 1
        buf = ctypes.create_string_buffer(4096)
        L.kcnn_component_info(ctypes.c_void_p(h), buf, 4096)
        info = buf.value.decode()
        assert info.startswith("MaxpoolComponent") and "pool-width-dim=2" in info and "pool-channel-dim=2" in info
        c = L.kcnn_component_copy(ctypes.c_void_p(h))
        assert c, err(L)
        assert write(L, c, True) == write(L, h, True) and write(L, c, False) == write(L, h, False)
        L.kcnn_component_delete(ctypes.c_void_p(c))
        # not updatable: no parameters, no learning rate
        p, r, cc, s = ctypes.c_void_p(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert L.kcnn_component_params(ctypes.c_void_p(h), 0, ctypes.byref(p), ctypes.byref(r), ctypes.byref(cc),
                                       ctypes.byref(s)) == -1
        assert L.kcnn_component_gradient_floats(ctypes.c_void_p(h)) == 0
    finally:
        L.kcnn_component_delete(ctypes.c_void_p(h))


def test_maxpool_model_file_bytes(L):
    """Read(golden) -> Write reproduces the stream assembled from the reference's Write() (:936-959), both
    modes, and text <-> binary carry the same values; the stream from before the overlap flags loads (:917-934)."""
    txt, bin_ = blob("maxpool.txt"), blob("maxpool.bin")
    a = L.kcnn_component_read(txt, len(txt), 0)
    b = L.kcnn_component_read(bin_, len(bin_), 1)
    assert a and b, err(L)
    assert write(L, a, False) == txt and write(L, b, True) == bin_
    assert write(L, a, True) == bin_ and write(L, b, False) == txt
    old = blob("maxpool_old.txt")
    o = L.kcnn_component_read(old, len(old), 0)
    assert o, err(L)
    assert write(L, o, False) == txt
    for h in (a, b, o):
        L.kcnn_component_delete(ctypes.c_void_p(h))
    # a truncated stream is an error, not a half-initialised component
    assert not L.kcnn_component_read(bin_[: len(bin_) // 2], len(bin_) // 2, 1)
    assert not L.kcnn_component_read(b"<NoSuchComponent> ", 18, 0)


@pytest.mark.parametrize("line", [CONV, FC])
def test_components_with_parameters_need_the_device(L, line):
    assert not new(L, line)
    assert "no CUDA device is selected" in err(L) and "no CPU fallback" in err(L)


@pytest.mark.parametrize("name", ["conv.txt", "conv.bin", "fc.txt", "fc.bin"])
def test_reading_a_model_with_parameters_needs_the_device(L, name):
    data = blob(name)
    assert not L.kcnn_component_read(data, len(data), int(name.endswith(".bin")))
    assert "no CUDA device is selected" in err(L)


def test_propagate_fails_loudly_without_a_device(L):
    """No CPU path: the max-pool component exists here (no parameters), but running it raises."""
    h = new(L, POOL)
    assert h, err(L)
    x = (ctypes.c_float * (4 * 1024))()
    y = (ctypes.c_float * (4 * 256))()
    rc = L.kcnn_component_propagate(ctypes.c_void_p(h), 4, x, 4, 1024, 1024, y, 4, 256, 256)
    assert rc == -1 and "no CUDA device is selected" in err(L)
    L.kcnn_component_delete(ctypes.c_void_p(h))
    rc = L.kcnn_mat_maxpool_prop(x, 4, 1024, 1024, 1, 8, 1, 2, 2, 0, 0, y, 4, 256, 256)
    assert rc == -1 and "no CUDA device is selected" in err(L)


def test_network_needs_the_device(L):
    cfg = open(os.path.join(capi.HERE, "configs", "nnet_c2_intermap.config")).read()
    assert not L.kcnn_nnet_new_from_config(cfg.encode(), 0)
    assert "no CUDA device is selected" in err(L)
    # a network without parameters is pure host logic: dimensions chain, Splice lines can be dropped
    two = POOL + "\n" + "MaxpoolComponent in-height=1 in-width=4 in-channel=64 pool-height-dim=1 pool-width-dim=2 pool-channel-dim=1\n"
    n = L.kcnn_nnet_new_from_config(two.encode(), 0)
    assert n, err(L)
    assert L.kcnn_nnet_num_components(ctypes.c_void_p(n)) == 2
    assert L.kcnn_nnet_input_dim(ctypes.c_void_p(n)) == 1024 and L.kcnn_nnet_output_dim(ctypes.c_void_p(n)) == 128
    L.kcnn_nnet_delete(ctypes.c_void_p(n))
    # dimension mismatch between consecutive components is an error (Nnet::Check)
    bad = POOL + "\n" + POOL + "\n"
    assert not L.kcnn_nnet_new_from_config(bad.encode(), 0)
    assert err(L)


def test_peer_memory_launchers_validate_their_arguments(L):
    """kernels_p2p.cu: rank / world / channel / 16-byte alignment are checked before anything is launched."""
    bases = (ctypes.c_ulonglong * 8)(*([0x1000] * 8))
    f = L.kcnn_p2p_allreduce_f32
    assert L.kcnn_p2p_flag_floats() == 8192
    assert f(None, bases, 0, 0, 0, 1024, 4096, 0) == -1          # world < 1
    assert f(None, bases, 0, 9, 0, 1024, 4096, 0) == -1          # world > 8
    assert f(None, bases, 2, 2, 0, 1024, 4096, 0) == -1          # rank >= world
    assert f(None, bases, 0, 2, 0, 1024, 4096, 2) == -1          # channel
    assert f(None, bases, 0, 2, 0, 1022, 4096, 0) == -1          # count not a multiple of 4 floats
    assert f(None, bases, 0, 2, 2, 1024, 4096, 0) == -1          # offset not a multiple of 4 floats
    assert f(None, bases, 0, 2, 0, 0, 4096, 0) == 0              # nothing to do
    assert f(None, bases, 0, 1, 0, 1024, 4096, 0) == 0           # one rank: the sum is the input
    g = L.kcnn_p2p_reduce_sgd_f32
    args = dict(momentum=ctypes.c_float(0.9), decay=ctypes.c_float(0.0), grad=ctypes.c_float(0.1))
    def call(rank, world, off, count, wfloats, delta, ch):
        return g(None, bases, ctypes.c_ulonglong(0), rank, world, ctypes.c_size_t(off), ctypes.c_size_t(count),
                 ctypes.c_size_t(wfloats), ctypes.c_size_t(delta), None, args["momentum"], args["decay"], args["grad"],
                 ctypes.c_size_t(4096), ch)
    assert call(0, 2, 0, 1024, 2048, 4096, 0) == -1              # weight part larger than the bucket
    assert call(0, 2, 0, 1024, 1022, 4096, 0) == -1              # weight part not 16-byte aligned
    assert call(0, 2, 0, 1024, 1024, 4098, 0) == -1              # parameter arena offset not aligned
    assert call(3, 2, 0, 1024, 1024, 4096, 0) == -1
    assert call(0, 2, 0, 0, 0, 4096, 0) == 0
    assert L.kcnn_p2p_allreduce_multicast_f32(None, bases, ctypes.c_ulonglong(0), 0, 2, 0, 1024, 4096, 0) == -1


# ---- the element-wise glue of the C2 model: no parameters, so the host side runs here ----

def _i32(v):
    import struct
    return b"\x04" + struct.pack("<i", v)


def _f32(v):
    import struct
    return b"\x04" + struct.pack("<f", v)


def _nonlinear_stream(kind, dim, binary):
    """NonlinearComponent::Write (nnet2/nnet-component.cc:398-412) of a component that has not seen data: empty
    double-precision statistics vectors ("DV" + int32 0 in binary mode, " [ ]\\n" in text mode), count 0.0."""
    if binary:
        return (b"<%s> <Dim> " % kind + _i32(dim) + b"<ValueSum> DV " + _i32(0) + b"<DerivSum> DV " + _i32(0) +
                b"<Count> \x08" + b"\x00" * 8 + b"</%s> " % kind)
    return b"<%s> <Dim> %d <ValueSum>  [ ]\n<DerivSum>  [ ]\n<Count> 0 </%s> " % (kind, dim, kind)


@pytest.mark.parametrize("kind, dim", [(b"RectifiedLinearComponent", 256), (b"SoftmaxComponent", 3454),
                                       (b"NormalizeComponent", 32)])
def test_nonlinear_components_host_side(L, kind, dim):
    h = new(L, "%s dim=%d" % (kind.decode(), dim))                    # InitFromString :418-427
    assert h, err(L)
    assert L.kcnn_component_type(ctypes.c_void_p(h)) == kind
    assert L.kcnn_component_input_dim(ctypes.c_void_p(h)) == dim == L.kcnn_component_output_dim(ctypes.c_void_p(h))
    for binary in (False, True):
        want = _nonlinear_stream(kind, dim, binary)
        assert write(L, h, binary) == want
        r = L.kcnn_component_read(want, len(want), int(binary))
        assert r, err(L)
        assert write(L, r, not binary) == _nonlinear_stream(kind, dim, not binary)
        L.kcnn_component_delete(ctypes.c_void_p(r))
    L.kcnn_component_delete(ctypes.c_void_p(h))
    assert not new(L, "%s dim=%d extra=1" % (kind.decode(), dim))
    assert "Invalid initializer for layer of type " + kind.decode() in err(L)
    assert not new(L, kind.decode())
    assert "Invalid initializer" in err(L)


def test_dropout_component_host_side(L):
    """DropoutComponent::InitFromString / Write (nnet2/nnet-component.cc:3548-3581): <Dim>, <DropoutScale>,
    <DropoutProportion> in that order; defaults dropout-proportion 0.5, dropout-scale 0."""
    h = new(L, "DropoutComponent dim=16 dropout-proportion=0.2 dropout-scale=0.5")
    assert h, err(L)
    assert write(L, h, False) == b"<DropoutComponent> <Dim> 16 <DropoutScale> 0.5 <DropoutProportion> 0.2 </DropoutComponent> "
    assert write(L, h, True) == (b"<DropoutComponent> <Dim> " + _i32(16) + b"<DropoutScale> " + _f32(0.5) +
                                 b"<DropoutProportion> " + _f32(0.2) + b"</DropoutComponent> ")
    # BackpropNeedsInput / Output: the backward pass is d * y / x (:3634-3636)
    assert L.kcnn_component_backprop_needs_input(ctypes.c_void_p(h)) == 1
    assert L.kcnn_component_backprop_needs_output(ctypes.c_void_p(h)) == 1
    L.kcnn_component_delete(ctypes.c_void_p(h))
    d = new(L, "DropoutComponent dim=16")                  # defaults: proportion 0.5, scale 0 (:3551)
    assert d, err(L)
    assert write(L, d, False) == b"<DropoutComponent> <Dim> 16 <DropoutScale> 0 <DropoutProportion> 0.5 </DropoutComponent> "
    L.kcnn_component_delete(ctypes.c_void_p(d))
    for bad in ("DropoutComponent dropout-proportion=0.2", "DropoutComponent dim=0", "DropoutComponent dim=16 keep=1"):
        assert not new(L, bad)
        assert "Invalid initializer for layer of type DropoutComponent" in err(L)


def test_splice_component_host_side(L):
    """SpliceComponent::InitFromString / Write (nnet2/nnet-component.cc:2549-2572, 2854-2863): left / right
    context expands to consecutive offsets, `context=` takes an explicit list, const-component-dim columns are
    not spliced; output-dim = (input-dim - const) * |context| + const."""
    h = new(L, "SpliceComponent input-dim=40 left-context=10 right-context=10")     # egs/exp/nnet/nnet.config:1
    assert h, err(L)
    assert L.kcnn_component_input_dim(ctypes.c_void_p(h)) == 40
    assert L.kcnn_component_output_dim(ctypes.c_void_p(h)) == 40 * 21
    ctx = " ".join(str(i) for i in range(-10, 11))
    assert write(L, h, False) == ("<SpliceComponent> <InputDim> 40 <Context> [ %s ]\n<ConstComponentDim> 0 "
                                  "</SpliceComponent> " % ctx).encode()
    L.kcnn_component_delete(ctypes.c_void_p(h))
    g = new(L, "SpliceComponent input-dim=40 context=-2:0:3 const-component-dim=4")
    assert g, err(L)
    assert L.kcnn_component_output_dim(ctypes.c_void_p(g)) == 36 * 3 + 4
    import struct
    want = (b"<SpliceComponent> <InputDim> " + _i32(40) + b"<Context> " + _i32(3) + struct.pack("<iii", -2, 0, 3) +
            b"<ConstComponentDim> " + _i32(4) + b"</SpliceComponent> ")
    assert write(L, g, True) == want
    r = L.kcnn_component_read(want, len(want), 1)
    assert r, err(L)
    assert write(L, r, False) == b"<SpliceComponent> <InputDim> 40 <Context> [ -2 0 3 ]\n<ConstComponentDim> 4 </SpliceComponent> "
    L.kcnn_component_delete(ctypes.c_void_p(r))
    L.kcnn_component_delete(ctypes.c_void_p(g))
    for bad in ("SpliceComponent input-dim=40", "SpliceComponent left-context=1 right-context=1",
                "SpliceComponent input-dim=0 left-context=1 right-context=1",
                "SpliceComponent input-dim=40 left-context=1 right-context=1 more=1"):
        assert not new(L, bad)
        assert "Invalid initializer for layer of type SpliceComponent" in err(L), err(L)


def test_chunk_info_checks_come_before_the_device(L):
    """ChunkInfo::CheckSize / Check (nnet2/nnet-component.cc:2580-2622): MaxpoolComponent::Propagate checks both
    matrices against their ChunkInfo first (nnet0/nnet-component-nnet0.cc:874-875) -- a wrong chunk count or an
    inverted offset range is an assertion, a consistent call reaches the device layer (and fails: none here)."""
    h = ctypes.c_void_p(new(L, POOL))
    x = (ctypes.c_float * (8 * 1024))()
    y = (ctypes.c_float * (8 * 256))()
    assert L.kcnn_component_propagate(h, 3, x, 4, 1024, 1024, y, 4, 256, 256) == -1      # 4 rows are not 3 chunks
    assert "CheckSize" in err(L)
    assert L.kcnn_component_propagate_chunks(h, 2, 0, 2, 0, 1, x, 4, 1024, 1024, y, 4, 256, 256) == -1
    assert "CheckSize" in err(L)                                                       # 2 chunks x 3 frames != 4 rows
    assert L.kcnn_component_propagate_chunks(h, 2, 1, 0, 0, 1, x, 4, 1024, 1024, y, 4, 256, 256) == -1
    assert "KALDI_ASSERT" in err(L)                                                    # last offset < first offset
    assert L.kcnn_component_propagate_chunks(h, 2, 0, 1, 0, 1, x, 4, 1024, 1024, y, 4, 256, 256) == -1
    assert "no CUDA device is selected" in err(L)                                      # consistent: reaches the device
    L.kcnn_component_delete(h)


def test_model_with_only_glue_parses_like_nnet_config(L):
    """Nnet::Init over config text: one component per line, '#' comments and blank lines skipped, dimensions
    must chain (Nnet::Check), skip_splice drops the SpliceComponent line."""
    cfg = ("# front end\nSpliceComponent input-dim=40 left-context=10 right-context=10\n\n"
           "RectifiedLinearComponent dim=840\nNormalizeComponent dim=840\nSoftmaxComponent dim=840\n")
    n = L.kcnn_nnet_new_from_config(cfg.encode(), 0)
    assert n, err(L)
    assert L.kcnn_nnet_num_components(ctypes.c_void_p(n)) == 4
    assert L.kcnn_nnet_input_dim(ctypes.c_void_p(n)) == 40 and L.kcnn_nnet_output_dim(ctypes.c_void_p(n)) == 840
    # (kcnn_nnet_frames_per_example belongs to the updater, which owns streams and buffers: device only)
    assert L.kcnn_nnet_frames_per_example(ctypes.c_void_p(n)) == -1 and "no CUDA device" in err(L)
    L.kcnn_nnet_delete(ctypes.c_void_p(n))
    m = L.kcnn_nnet_new_from_config(cfg.encode(), 1)
    assert m, err(L)
    assert L.kcnn_nnet_num_components(ctypes.c_void_p(m)) == 3 and L.kcnn_nnet_input_dim(ctypes.c_void_p(m)) == 840
    L.kcnn_nnet_delete(ctypes.c_void_p(m))
    assert not L.kcnn_nnet_new_from_config(cfg.replace("dim=840\nSoftmax", "dim=841\nSoftmax").encode(), 0)


def test_reals_of_either_width_are_read(L):
    """Upstream's ReadBasicType<float> accepts 4- and 8-byte reals (a model written by a double-precision Kaldi);
    in text mode a real may be followed directly by the next token's '<' (peek-based parsing, no multi-char putback)."""
    import struct
    wide = (b"<DropoutComponent> <Dim> " + _i32(16) + b"<DropoutScale> \x08" + struct.pack("<d", 0.5) +
            b"<DropoutProportion> \x08" + struct.pack("<d", float(struct.unpack("<f", struct.pack("<f", 0.2))[0])) +
            b"</DropoutComponent> ")
    r = L.kcnn_component_read(wide, len(wide), 1)
    assert r, err(L)
    assert write(L, r, False) == b"<DropoutComponent> <Dim> 16 <DropoutScale> 0.5 <DropoutProportion> 0.2 </DropoutComponent> "
    L.kcnn_component_delete(ctypes.c_void_p(r))
    bad = wide.replace(b"<DropoutScale> \x08", b"<DropoutScale> \x03")
    assert not L.kcnn_component_read(bad, len(bad), 1)                 # a size byte that is neither 4 nor 8
    for txt in (b"<DropoutComponent> <Dim> 16 <DropoutScale> 5e-1 <DropoutProportion> 0.2 </DropoutComponent> ",
                b"<DropoutComponent>  <Dim>  16\n<DropoutScale>\t0.5 <DropoutProportion> .2 </DropoutComponent>\n"):
        t = L.kcnn_component_read(txt, len(txt), 0)
        assert t, err(L)
        assert write(L, t, True) == (b"<DropoutComponent> <Dim> " + _i32(16) + b"<DropoutScale> " + _f32(0.5) +
                                     b"<DropoutProportion> " + _f32(0.2) + b"</DropoutComponent> ")
        L.kcnn_component_delete(ctypes.c_void_p(t))
    for txt in (b"<DropoutComponent> <Dim> 16 <DropoutScale> abc <DropoutProportion> 0.2 </DropoutComponent> ",
                b"<DropoutComponent> <Dim> 16 <DropoutProportion> 0.2 <DropoutScale> 0.5 </DropoutComponent> ",
                b"<DropoutComponent> <Dim> 16 <DropoutScale> 0.5 <DropoutProportion> 0.2 "):
        assert not L.kcnn_component_read(txt, len(txt), 0)             # bad number / token order / missing end token
