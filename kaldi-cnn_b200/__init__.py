"""kaldi-cnn_b200: B200-native (sm_100a) implementation of the kaldi-cnn CNN-layer
hot path -- the CuMatrixBase extensions of src/cnslmat/conv2D.cc and the nnet0
ConvolutionComponent / MaxpoolComponent / FullyConnectedComponent.

The product is the C-ABI shared library ``lib/libkaldicnn_b200.so`` (headers in
``/include``); host code is Kaldi-style C++ (csrc/).  This Python package is
only the loader plus thin ctypes views used by tests and bench.py.  There is no
CPU fallback: if the library is missing or there is no GPU, calls fail loudly.
"""
import os

from . import capi  # noqa: F401
from .capi import lib, load, LIB_PATH  # noqa: F401

__all__ = ["capi", "lib", "load", "LIB_PATH"]
PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
