"""Independent numpy formulation of the hot path (TEST INFRASTRUCTURE ONLY).

Where kcnn_oracle_impl.h restates the reference's loops op for op, this module
writes the same mathematics a second way -- einsum over the tensor layout the
reference uses -- so the two can be cross-checked (the literal
transcription and the einsum form agreed to 1e-15 in FP64).

Layout (cnsl-cu-kernels.cu:28-32, 51-54, 243-249): every activation row is a
[C][W][H] tensor with H fastest; the kernel matrix [(C*KW*KH) x G] has rows
ordered [c][kw][kh] with kh fastest; forward is cross-correlation.
"""
import numpy as np


def act(x, H, W, C):
    """[N x H*W*C] -> [N, C, W, H]."""
    return np.asarray(x).reshape(x.shape[0], C, W, H)


def kern(k, KH, KW, C, G):
    """[(KH*KW*C) x G] -> [C, KW, KH, G]."""
    return np.asarray(k).reshape(C, KW, KH, G)


def _windows(xp, KH, KW):
    """[N, C, Wp, Hp] -> [N, C, OW, OH, KW, KH] sliding windows (view)."""
    return np.lib.stride_tricks.sliding_window_view(xp, (KW, KH), axis=(2, 3))


def conv_fprop(x, k, bias, H, W, C, pad_h, pad_w, KH, KW, G):
    """Y[n,g,ow,oh] = b[g] + sum_{c,kw,kh} Xp[n,c,ow+kw,oh+kh] K[c,kw,kh,g]."""
    xp = np.pad(act(x, H, W, C), ((0, 0), (0, 0), (pad_w, pad_w), (pad_h, pad_h)))
    win = _windows(xp, KH, KW)
    y = np.einsum("ncwhab,cabg->ngwh", win, kern(k, KH, KW, C, G), optimize=True)
    if bias is not None:
        y = y + np.asarray(bias).reshape(1, G, 1, 1)
    return y.reshape(x.shape[0], -1)


def conv_dgrad(dy, k, H, W, C, pad_h, pad_w, KH, KW, G):
    """dX[n,c,w,h] = sum_{g,kw,kh} dY[n,g,w+pw-kw,h+ph-kh] K[c,kw,kh,g]."""
    OH, OW = H + 2 * pad_h - KH + 1, W + 2 * pad_w - KW + 1
    d = np.asarray(dy).reshape(dy.shape[0], G, OW, OH)
    kk = kern(k, KH, KW, C, G)
    dxp = np.zeros((dy.shape[0], C, W + 2 * pad_w, H + 2 * pad_h), dtype=d.dtype)
    for kw in range(KW):
        for kh in range(KH):
            dxp[:, :, kw:kw + OW, kh:kh + OH] += np.einsum("ngwh,cg->ncwh", d, kk[:, kw, kh, :])
    dx = dxp[:, :, pad_w:pad_w + W, pad_h:pad_h + H]
    return np.ascontiguousarray(dx).reshape(dy.shape[0], -1)


def conv_wgrad(x, dy, H, W, C, pad_h, pad_w, KH, KW, G):
    """dK[c,kw,kh,g] = sum_{n,ow,oh} Xp[n,c,ow+kw,oh+kh] dY[n,g,ow,oh]; db[g] = sum dY."""
    OH, OW = H + 2 * pad_h - KH + 1, W + 2 * pad_w - KW + 1
    xp = np.pad(act(x, H, W, C), ((0, 0), (0, 0), (pad_w, pad_w), (pad_h, pad_h)))
    d = np.asarray(dy).reshape(dy.shape[0], G, OW, OH)
    win = _windows(xp, KH, KW)
    dk = np.einsum("ncwhab,ngwh->cabg", win, d, optimize=True)
    return dk.reshape(C * KW * KH, G), d.sum(axis=(0, 2, 3))


def sgd(lin, bias, prev, dk, db, n, learning_rate, weight_decay, momentum):
    """SURVEY App. A 'update' (ascent): lr = learning_rate / N."""
    lr = learning_rate / n
    prev = momentum * prev - lr * weight_decay * lin + lr * dk
    return lin + prev, bias + lr * db, prev


def maxpool_fwd(x, H, W, C, ph, pw, pc):
    t = np.asarray(x).reshape(x.shape[0], C // pc, pc, W // pw, pw, H // ph, ph)
    return t.max(axis=(2, 4, 6)).reshape(x.shape[0], -1)


def maxpool_bwd_ties(x, y, dy, H, W, C, ph, pw, pc):
    """err goes to EVERY window element equal to the max (cnsl-cu-kernels.cu:302-303)."""
    n = x.shape[0]
    t = np.asarray(x).reshape(n, C // pc, pc, W // pw, pw, H // ph, ph)
    yy = np.asarray(y).reshape(n, C // pc, 1, W // pw, 1, H // ph, 1)
    dd = np.asarray(dy).reshape(n, C // pc, 1, W // pw, 1, H // ph, 1)
    return np.where(t == yy, dd, np.zeros((), dtype=t.dtype)).reshape(n, -1)
