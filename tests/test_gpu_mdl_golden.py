"""Checkpoint byte compatibility: the component Read / Write token streams against
fixtures assembled BY HAND from the reference's Write() statements and Kaldi's stream conventions
(tests/golden/make_mdl_golden.py -- token order of nnet0/nnet-component-nnet0.cc:621-666, 936-959,
1001-1020; the <AvgInput> back-compatibility branch :603-618; the pre-overlap MaxpoolComponent stream
:917-934).  Read(golden) then Write must reproduce the golden bytes exactly, in text and binary mode; a
text fixture re-written in binary mode must equal the binary fixture (same values both ways).
GPU-marked only because the components keep their parameters in device memory."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from kaldi_cnn_b200 import components as kc  # noqa: E402

MDL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mdl")


def blob(name):
    return open(os.path.join(MDL, name), "rb").read()


def test_fixtures_are_what_the_generator_writes():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(MDL, "..", "make_mdl_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    for name, fn in mk.FILES.items():
        assert fn() == blob(name), name


@pytest.mark.parametrize("stem", ["conv", "maxpool", "fc"])
def test_write_reproduces_the_reference_byte_stream(stem):
    txt, bin_ = blob(stem + ".txt"), blob(stem + ".bin")
    c_txt = kc.Component.read(txt, binary=False)
    c_bin = kc.Component.read(bin_, binary=True)
    assert c_txt.write(binary=False) == txt
    assert c_bin.write(binary=True) == bin_
    assert c_txt.write(binary=True) == bin_            # text -> binary: the same values
    assert c_bin.write(binary=False) == txt
    assert c_txt.type == {"conv": "ConvolutionComponent", "maxpool": "MaxpoolComponent", "fc": "FullyConnectedComponent"}[stem]


def test_conv_fields_land_where_the_reference_puts_them():
    c = kc.Component.read(blob("conv.bin"), binary=True)
    assert c.input_dim == 4 * 5 * 2 and c.output_dim == 3 * 5 * 3
    lin, bias, prev = (c.params(k).cpu().numpy() for k in range(3))
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(MDL, "..", "make_mdl_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    v = mk.conv_values()
    assert np.array_equal(lin, v["lin"]) and np.array_equal(bias[0], v["bias"]) and np.array_equal(prev, v["prev"])
    assert c.weight_decay_momentum() == (np.float32(0.0002), np.float32(0.9))
    assert abs(c.info().count("padding-width=1")) == 1


def test_back_compatibility_branches():
    canon_txt, canon_bin = blob("conv.txt"), blob("conv.bin")
    # <AvgInput> / <AvgInputCount> of old model files are read and dropped
    assert kc.Component.read(blob("conv_avginput.txt"), binary=False).write(binary=False) == canon_txt
    assert kc.Component.read(blob("conv_avginput.bin"), binary=True).write(binary=True) == canon_bin
    # <IsGradient> T survives a round trip
    g = blob("conv_gradient.txt")
    assert kc.Component.read(g, binary=False).write(binary=False) == g
    # MaxpoolComponent streams written before the overlap flags existed
    assert kc.Component.read(blob("maxpool_old.txt"), binary=False).write(binary=False) == blob("maxpool.txt")


def test_double_precision_reals_are_accepted():
    """A Kaldi built with --double-precision writes 8-byte reals; upstream's readers accept either width."""
    import struct
    b = blob("fc.bin")
    lr = b"<LearningRate> \x04" + struct.pack("<f", 0.008)
    assert lr in b
    wide = b.replace(lr, b"<LearningRate> \x08" + struct.pack("<d", float(np.float32(0.008))))
    assert kc.Component.read(wide, binary=True).write(binary=True) == b
