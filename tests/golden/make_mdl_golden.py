"""Byte-level golden model fragments for the three nnet0 components.

The reference ships no model file and cannot be built here, so the fixtures are ASSEMBLED BY HAND from
the reference's Write() statements -- token order of
    ConvolutionComponent::Write     src/nnet0/nnet-component-nnet0.cc:621-666
    MaxpoolComponent::Write         :936-959
    FullyConnectedComponent::Write  :1001-1020
and the back-compatibility branch of ConvolutionComponent::Read (:603-618, <AvgInput>) -- and from
Kaldi's stream conventions (base/io-funcs-inl.h, matrix/kaldi-matrix.cc, kaldi-vector.cc upstream):
    token            "<Name> "
    int32 / float    binary: size byte (4) + little-endian value;  text: "%d " / "%.7g " (precision 7,
                     base/kaldi-io.cc InitKaldiOutputStream)
    bool             'T' / 'F' (text: followed by a space)
    matrix           binary: "FM " + int32 rows + int32 cols + row-major float32;
                     text:   " [" + per row "\n  " + values each followed by " " + "]\n"
    vector           binary: "FV " + int32 dim + float32;  text: " [ " + values + "]\n"
Nothing here imports the product or the oracle: the bytes are written by this script alone.

    python tests/golden/make_mdl_golden.py        # rewrites tests/golden/mdl/*
"""
import os
import struct

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mdl")


class W:
    def __init__(self, binary):
        self.b, self.out = binary, bytearray()

    def tok(self, t):
        self.out += (t + " ").encode()

    def i32(self, v):
        self.out += (b"\x04" + struct.pack("<i", v)) if self.b else ("%d " % v).encode()

    def f32(self, v):
        self.out += (b"\x04" + struct.pack("<f", v)) if self.b else (("%.7g" % np.float32(v)) + " ").encode()

    def boolean(self, v):
        self.out += (b"T" if v else b"F") + (b"" if self.b else b" ")

    def mat(self, m):
        m = np.asarray(m, np.float32)
        if self.b:
            self.tok("FM")
            self.i32(m.shape[0]); self.i32(m.shape[1])
            self.out += m.tobytes()
        else:
            s = " ["
            for row in m:
                s += "\n  " + "".join(("%.7g" % v) + " " for v in row)
            self.out += (s + "]\n").encode()

    def vec(self, v):
        v = np.asarray(v, np.float32)
        if self.b:
            self.tok("FV")
            self.i32(v.shape[0])
            self.out += v.tobytes()
        else:
            self.out += (" [ " + "".join(("%.7g" % x) + " " for x in v) + "]\n").encode()


def conv_values():
    rng = np.random.default_rng(2024)
    KH, KW, C, G = 2, 3, 2, 3
    lin = (rng.integers(-40, 40, (KH * KW * C, G)) / 64.0).astype(np.float32)      # exact in "%.7g"
    bias = np.array([0.5, -0.25, 0.125], np.float32)
    prev = (rng.integers(-40, 40, (KH * KW * C, G)) / 128.0).astype(np.float32)       # <= 7 significant digits: exact in "%.7g"
    return dict(in_height=4, in_width=5, in_channel=C, kernel_height=KH, kernel_width=KW, stride=1, padding_height=0,
                padding_width=1, group=G, out_height=3, out_width=5, lr=0.02, wd=0.0002, mom=0.9, lin=lin, bias=bias,
                prev=prev)


def conv_stream(binary, avg_input=False, is_gradient=False):
    v, w = conv_values(), W(binary)
    w.tok("<ConvolutionComponent>")
    for name in ("in_height", "in_width", "in_channel", "kernel_height", "kernel_width", "stride", "padding_height",
                 "padding_width", "group", "out_height", "out_width"):
        w.tok("<%s>" % name); w.i32(v[name])
    w.tok("<LearningRate>"); w.f32(v["lr"])
    w.tok("<WeightDecay>"); w.f32(v["wd"])
    w.tok("<Momentum>"); w.f32(v["mom"])
    w.tok("<LinearParams>"); w.mat(v["lin"])
    w.tok("<BiasParams>"); w.vec(v["bias"])
    w.tok("<PrevGrad>"); w.mat(v["prev"])
    if avg_input:                                   # old files: read and discarded (:603-610)
        w.tok("<AvgInput>"); w.vec(np.arange(v["in_height"] * v["in_width"] * v["in_channel"]) / 8.0)
        w.tok("<AvgInputCount>"); w.f32(17.0)
    w.tok("<IsGradient>"); w.boolean(is_gradient)
    w.tok("</ConvolutionComponent>")
    return bytes(w.out)


def maxpool_stream(binary, with_overlap_tokens=True):
    w = W(binary)
    w.tok("<MaxpoolComponent>")
    for name, val in (("InputDim", 1 * 12 * 8), ("in_height", 1), ("in_width", 12), ("in_channel", 8), ("OutputDim", 6 * 4),
                      ("PoolHeightDim", 1), ("PoolWidthDim", 2), ("PoolChannelDim", 2)):
        w.tok("<%s>" % name); w.i32(val)
    if with_overlap_tokens:
        w.tok("<Overlap>"); w.boolean(False)
        w.tok("<Overlap2D>"); w.boolean(False)
    w.tok("</MaxpoolComponent>")
    return bytes(w.out)


def fc_stream(binary):
    rng = np.random.default_rng(7)
    lin = (rng.integers(-64, 64, (3, 4)) / 128.0).astype(np.float32)
    bias = np.array([1.0, 1.0, 1.0], np.float32)
    prev = (rng.integers(-64, 64, (3, 4)) / 128.0).astype(np.float32)
    w = W(binary)
    w.tok("<FullyConnectedComponent>")
    w.tok("<LearningRate>"); w.f32(0.008)
    w.tok("<LinearParams>"); w.mat(lin)
    w.tok("<BiasParams>"); w.vec(bias)
    w.tok("<WeightDecay>"); w.f32(0.0005)
    w.tok("<Momentum>"); w.f32(0.9)
    w.tok("<PrevGrad>"); w.mat(prev)
    w.tok("</FullyConnectedComponent>")
    return bytes(w.out)


FILES = {
    "conv.txt": lambda: conv_stream(False), "conv.bin": lambda: conv_stream(True),
    "conv_gradient.txt": lambda: conv_stream(False, is_gradient=True),
    "conv_avginput.txt": lambda: conv_stream(False, avg_input=True), "conv_avginput.bin": lambda: conv_stream(True, avg_input=True),
    "maxpool.txt": lambda: maxpool_stream(False), "maxpool.bin": lambda: maxpool_stream(True),
    "maxpool_old.txt": lambda: maxpool_stream(False, with_overlap_tokens=False),
    "fc.txt": lambda: fc_stream(False), "fc.bin": lambda: fc_stream(True),
}

if __name__ == "__main__":
    os.makedirs(HERE, exist_ok=True)
    for name, fn in FILES.items():
        with open(os.path.join(HERE, name), "wb") as f:
            f.write(fn())
    print("wrote", len(FILES), "fixtures to", HERE)
