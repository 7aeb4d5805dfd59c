"""Generates tests/golden/*.npz: fixed-seed input / output vectors of the hot path.

    python tests/golden/make_golden.py

Outputs come from the INDEPENDENT NumPy formulation (oracle/oracle_np.py: einsum / reshape on
the layout of SURVEY Appendix A) evaluated in float64, so the fixtures pin both the C
restatement (oracle/kcnn_oracle_impl.h, checked on CPU in tests/test_golden.py) and the CUDA
path (checked with -m gpu).  The reference ships no golden vectors and cannot be built in
this image, so these are the committed fixtures parity is anchored on.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_np as onp  # noqa: E402

# name -> (N, H, W, C, pad_h, pad_w, KH, KW, G)
CONV = {
    "conv_c1a": (6, 40, 11, 3, 0, 0, 40, 4, 16),
    "conv_c1b": (3, 40, 11, 3, 0, 0, 8, 3, 8),
    "conv_time_c32": (5, 1, 18, 32, 0, 0, 1, 3, 32),      # TMA-eligible time-axis layer
    "conv_time_pad": (9, 1, 10, 64, 0, 1, 1, 3, 96),      # TMA-eligible, zero padding
    "conv_pad2d": (2, 6, 7, 2, 1, 1, 3, 3, 5),
}
# name -> (N, H, W, C, ph, pw, pc, kind)
POOL = {
    "pool_3d": (4, 6, 4, 8, 3, 2, 2, "randn"),
    "pool_relu_ties": (5, 1, 8, 16, 1, 2, 2, "relu"),
    "pool_quantised": (3, 3, 9, 4, 3, 3, 2, "quant"),
}
FC = {"fc_small": (7, 24, 40), "fc_odd": (5, 30, 13)}
LR, WD, MOM = 0.02, 0.0002, 0.9


def main():
    for name, (N, H, W, C, ph, pw, KH, KW, G) in CONV.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        OH, OW = H + 2 * ph - KH + 1, W + 2 * pw - KW + 1
        x = rng.standard_normal((N, H * W * C)).astype(np.float32)
        k = (rng.standard_normal((KH * KW * C, G)) * 0.1).astype(np.float32)
        b = rng.standard_normal(G).astype(np.float32)
        dy = rng.standard_normal((N, OH * OW * G)).astype(np.float32)
        prev = (rng.standard_normal(k.shape) * 0.01).astype(np.float32)
        x64, k64, b64, dy64, p64 = (a.astype(np.float64) for a in (x, k, b, dy, prev))
        y = onp.conv_fprop(x64, k64, b64, H, W, C, ph, pw, KH, KW, G)
        dx = onp.conv_dgrad(dy64, k64, H, W, C, ph, pw, KH, KW, G)
        dk, db = onp.conv_wgrad(x64, dy64, H, W, C, ph, pw, KH, KW, G)
        k2, b2, p2 = onp.sgd(k64, b64, p64, dk, db, N, LR, WD, MOM)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array([N, H, W, C, ph, pw, KH, KW, G]),
                            x=x, k=k, b=b, dy=dy, prev=prev, y=y, dx=dx, dk=dk, db=db, k_new=k2, b_new=b2,
                            prev_new=p2, hyper=np.array([LR, WD, MOM]))
    for name, (N, H, W, C, ph, pw, pc, kind) in POOL.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        x = rng.standard_normal((N, H * W * C)).astype(np.float32)
        if kind == "relu":
            x = np.maximum(x, 0)
        elif kind == "quant":
            x = np.round(x * 2) / 4
        y = onp.maxpool_fwd(x, H, W, C, ph, pw, pc).astype(np.float32)
        dy = rng.standard_normal(y.shape).astype(np.float32)
        dx = onp.maxpool_bwd_ties(x, y, dy, H, W, C, ph, pw, pc).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), shape=np.array([N, H, W, C, ph, pw, pc]),
                            x=x, y=y, dy=dy, dx=dx)
    for name, (N, din, dout) in FC.items():
        rng = np.random.default_rng(sum(map(ord, name)))
        x = rng.standard_normal((N, din)).astype(np.float32)
        w = (rng.standard_normal((dout, din)) * 0.1).astype(np.float32)
        b = rng.standard_normal(dout).astype(np.float32)
        dy = rng.standard_normal((N, dout)).astype(np.float32)
        prev = (rng.standard_normal(w.shape) * 0.01).astype(np.float32)
        x64, w64, b64, dy64, p64 = (a.astype(np.float64) for a in (x, w, b, dy, prev))
        y = x64 @ w64.T + b64
        dx = dy64 @ w64
        dw, db = dy64.T @ x64, dy64.sum(0)
        lr = LR / N
        p2 = MOM * p64 - lr * WD * w64 + lr * dw
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, w=w, b=b, dy=dy, prev=prev, y=y, dx=dx, dw=dw,
                            db=db, w_new=w64 + p2, b_new=b64 + lr * db, prev_new=p2, hyper=np.array([LR, WD, MOM]))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
