import os, sys, json, traceback
sys.path.insert(0, "/root/repo")
import numpy as np
import tests.test_gpu_fused_step as T
cfg_full = open(os.path.join(T.ROOT, "kaldi-cnn_b200", "configs", "nnet_c2_intermap.config")).read()
cfg_full = "\n".join(l for l in cfg_full.splitlines() if not l.startswith("SpliceComponent"))
for mode in ("rne", "rna"):
    for name, cfg, N, seed, steps in (("small", T.CFG, 96, 3, 3), ("conv2fc", T.CFG_CONV_TO_FC, 80, 5, 2), ("full", cfg_full, 512, 42, 2)):
        try:
            rep = T.step_vs_oracle(cfg, N, seed, steps=steps, tol_out=1.0, tol_step=10.0, gemm_operands=mode)
        except Exception:
            traceback.print_exc()
            continue
        worst_val = max(v[0] for k, v in rep.items() if isinstance(v, tuple))
        worst_step = max(v[1] for k, v in rep.items() if isinstance(v, tuple))
        print("MODEL", mode, name, "objf %.2e post %.2e worst_val %.2e worst_step %.2e" % (rep["objf"], rep["posteriors"], worst_val, worst_step), flush=True)
        print("   ", {k: ("%.1e/%.1e" % v if isinstance(v, tuple) else "%.1e" % v) for k, v in rep.items()}, flush=True)
