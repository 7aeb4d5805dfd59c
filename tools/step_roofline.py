"""Every launch of one training step of the C2 model (bench.py's workload, N = 512) against its
roofline: joins the per-step ncu launch list (profiles/r01_launches_bench_current.csv,
`gpu__time_duration.sum`, cold caches, serialised) with the ALGORITHMIC work of each launch
(conv / FC 2*M*N*K flop; HBM kernels 4 x (elements read + written)).

    python tools/step_roofline.py [launches.csv] > profiles/r01_step_roofline.md

The launch order is the updater's: forward conv1 .. FC3, softmax / cross-entropy, backward FC3 ..
conv1.  Peaks: MEASURED_PEAKS.json (HBM copy GB/s, bf16 dense TFLOP/s / 2 for TF32), else the
fall-backs of B200_PROFILING.md.  Durations under ncu are cold-cache: activations that are
L2-resident in the replayed step (everything but the FC weights) come from HBM here, so the
HBM-bound rows are a LOWER bound on what the step sees; the shares are what to read.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 512
# (name, C, W, KW, G, OW, H) ; H = kernel height = input height (1 except conv1: 40)
CONVS = [("conv1", 1, 21, 4, 128, 18, 40), ("conv2", 64, 18, 3, 128, 16, 1), ("conv3", 128, 16, 3, 256, 14, 1),
         ("conv4", 256, 14, 3, 256, 12, 1), ("conv5", 256, 6, 3, 512, 4, 1), ("conv6", 512, 4, 3, 512, 2, 1)]
FCS = [("FC1", 1024, 4096), ("FC2", 4096, 4096), ("FC3", 4096, 3454)]


def conv_flop(c):
    _, C, W, KW, G, OW, H = c
    return 2.0 * N * OW * G * KW * H * C


def conv_w(c):
    _, C, W, KW, G, OW, H = c
    return KW * H * C * G


def act(cols):
    return 4.0 * N * cols


def plan():
    """[(label, 'tensor' | 'hbm' | '-', flop or bytes)] in launch order."""
    c1, c2, c3, c4, c5, c6 = CONVS
    f1, f2, f3 = FCS
    out = lambda c: c[4] * c[5]                    # noqa: E731  output columns of a convolution
    inp = lambda c: c[1] * c[2] * c[6]             # noqa: E731
    p = []
    T, B = "tensor", "hbm"
    # ---- forward
    p += [("conv1 fprop + bias + ReLU", T, conv_flop(c1)),
          ("maxpool 1x18x128 pc=2 fwd", B, act(2304 + 1152)),
          ("pack conv2 input (channels-last)", B, 2 * act(inp(c2))), ("conv2 fprop + bias + ReLU", T, conv_flop(c2)),
          ("pack conv3 input", B, 2 * act(inp(c3))), ("conv3 fprop + bias + ReLU", T, conv_flop(c3)),
          ("pack conv4 input", B, 2 * act(inp(c4))), ("conv4 fprop + bias", T, conv_flop(c4)),
          ("maxpool 1x12x256 pw=2 fwd", B, act(3072 + 1536)), ("ReLU fwd 1536", B, 2 * act(1536)),
          ("pack conv5 input", B, 2 * act(inp(c5))), ("conv5 fprop, split-K 2", T, conv_flop(c5)),
          ("conv5 split reduce + bias + ReLU", B, 3 * act(out(c5))),
          ("pack conv6 input", B, 2 * act(inp(c6))), ("conv6 fprop, split-K 4", T, conv_flop(c6)),
          ("conv6 split reduce + bias + ReLU", B, 5 * act(out(c6)))]
    for i, (nm, din, dout) in enumerate(FCS):
        p.append(("%s fprop + bias%s" % (nm, " + ReLU" if i < 2 else ""), T, 2.0 * N * din * dout))
        if i < 2:
            p += [("dropout fwd 4096", B, 2 * act(4096)), ("dropout seed bump", "-", 0)]
    p += [("softmax fwd 3454", B, 2 * act(3454)), ("cross-entropy derivative", B, act(3454)),
          ("softmax bwd", B, 3 * act(3454)), ("softmax statistics (column sums)", B, act(3454))]
    # ---- backward: FC stack
    for i, (nm, din, dout) in reversed(list(enumerate(FCS))):
        if nm == "FC1":
            p += [("FC1 dgrad, split-K", T, 2.0 * N * din * dout), ("FC1 dgrad split reduce", B, 3 * act(din))]
        else:
            p.append(("%s dgrad" % nm, T, 2.0 * N * din * dout))
        p.append(("%s wgrad + momentum SGD (W, prev read + written)" % nm, B, 16.0 * din * dout + act(din) + act(dout)))
        p.append(("%s bias gradient (column sums)" % nm, B, act(dout)))
        if i > 0:
            p += [("dropout bwd 4096", B, 2 * act(4096)), ("ReLU bwd + statistics 4096", B, 3 * act(4096))]
    p.append(("ReLU bwd + statistics 1024", B, 3 * act(1024)))
    # ---- backward: convolutions (dY pack, dgrad [+ reduce], wgrad, split reduce + SGD), glue between
    glue = {"conv6": [("ReLU bwd + statistics 2048", B, 3 * act(2048))],
            "conv5": [("ReLU bwd + statistics 1536", B, 3 * act(1536)),
                      ("maxpool 1x12x256 pw=2 bwd (exact)", B, act(2 * 3072 + 2 * 1536))],
            "conv4": [("ReLU bwd + statistics 3584", B, 3 * act(3584))],
            "conv3": [("ReLU bwd + statistics 2048", B, 3 * act(2048))],
            "conv2": [("maxpool 1x18x128 pc=2 bwd (exact)", B, act(2 * 2304 + 2 * 1152)),
                      ("ReLU bwd + statistics 2304", B, 3 * act(2304))],
            "conv1": []}
    dgrad_split = {"conv6": 2, "conv5": 2, "conv3": 2}
    for c in reversed(CONVS):
        nm = c[0]
        p.append(("pack %s out_deriv (+ bias-gradient partials)" % nm, B, 2 * act(out(c))))
        if nm in dgrad_split:
            p += [("%s dgrad, split-K %d" % (nm, dgrad_split[nm]), T, conv_flop(c)),
                  ("%s dgrad split reduce" % nm, B, (dgrad_split[nm] + 1) * act(inp(c)))]
        else:
            p.append(("%s dgrad" % nm, T, conv_flop(c)))
        p.append(("%s wgrad, split-K" % nm, T, conv_flop(c)))
        p.append(("%s split reduce + momentum SGD + bias" % nm, B, 16.0 * conv_w(c)))
        p += glue[nm]
    return p


def peaks():
    hbm, tf32, src = 6542.1, 841.35, "fallback"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            hbm = float(d.get("hbm_gbs", hbm))
            tf32 = float(d.get("bf16_tflops", 2 * tf32)) / 2       # burst figure: each kernel is timed alone
            src = "MEASURED_PEAKS.json; TF32 = bf16 / 2"
        except Exception:
            pass
    return hbm, tf32, src


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_launches_bench_current.csv")
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    seq = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(",", ""))
        t = t / 1000.0 if r[ui] == "ns" else t * 1000.0 if r[ui] == "ms" else t
        seq.append((r[ki], t))
    pl = plan()
    if len(pl) != len(seq):
        raise SystemExit("plan has %d launches, the list %d: not the C2 step this script describes" % (len(pl), len(seq)))
    hbm, tf32, src = peaks()
    total = sum(t for _, t in seq)
    print("# One training step of the C2 model (N = 512), launch by launch, against the roofline\n")
    print("Source: `%s` (ncu `gpu__time_duration.sum`, cold caches, serialised: sum %.0f us; the replayed step "
          "takes 746 us in the same build). Peaks (%s): HBM %.0f GB/s, TF32 dense %.0f TFLOP/s. "
          "Generated by `tools/step_roofline.py`.\n" % (os.path.relpath(path, ROOT), total, src, hbm, tf32))
    print("| # | launch | kernel | us | share | bound | algorithmic | achieved | of peak |")
    print("|---|---|---|---|---|---|---|---|---|")
    agg = {}
    for i, ((label, bound, work), (kname, us)) in enumerate(zip(pl, seq), 1):
        short = kname.replace("void ", "").split("(")[0][:44]
        if bound == "tensor":
            ach = work / (us * 1e-6) / 1e12
            cell = "%.0f MFLOP | %.0f TFLOP/s | %.0f %%" % (work / 1e6, ach, 100 * ach / tf32)
        elif bound == "hbm":
            ach = work / (us * 1e-6) / 1e9
            cell = "%.1f MB | %.0f GB/s | %.0f %%" % (work / 1e6, ach, 100 * ach / hbm)
        else:
            cell = "- | - | -"
        print("| %d | %s | `%s` | %.1f | %.1f %% | %s | %s |" % (i, label, short, us, 100 * us / total, bound, cell))
        key = ("GEMMs (tensor-bound)" if bound == "tensor" else
               "FC wgrad + SGD (HBM-bound GEMM epilogue)" if "wgrad + momentum" in label else
               "split-K reductions (+ SGD)" if "reduce" in label else
               "staging packs" if label.startswith("pack") else "elementwise glue / pooling")
        a = agg.setdefault(key, [0.0, 0])
        a[0] += us
        a[1] += 1
    print("\n| group | launches | us | share |\n|---|---|---|---|")
    for k, (us, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("| %s | %d | %.0f | %.0f %% |" % (k, n, us, 100 * us / total))


if __name__ == "__main__":
    main()
