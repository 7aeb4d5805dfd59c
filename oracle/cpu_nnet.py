"""CPU training step of an nnet.config model, op for op as the reference's CPU build does it
(TEST INFRASTRUCTURE / CPU BASELINE ONLY -- used by tests/ and by bench.py's cpu_baseline and
--impl reference legs; never by the product).

Every layer calls the oracle's restatement of the reference component
(oracle/kcnn_oracle_impl.h: conv_propagate / conv_backprop / conv_update / maxpool_* / fc_*),
or, when oracle/_ref is built, the UNMODIFIED reference code compiled for CPU.
The loop is nnet2's NnetUpdater: Propagate through all components, cross-entropy objective and
derivative, Backprop in reverse with the update inside Backprop.
"""
import numpy as np

from . import oracle as ora


def parse_config(text, skip_splice=True):
    layers = []
    for line in text.replace("\r", "").split("\n"):
        line = line.strip()
        if not line or line.startswith("#"):
            continue
        parts = line.split()
        kind, kv = parts[0], {}
        for p in parts[1:]:
            k, v = p.split("=", 1)
            kv[k] = v
        if skip_splice and kind == "SpliceComponent":
            continue
        layers.append((kind, kv))
    return layers


def tf32_operand(a, mode):
    """What a TF32 tensor-core instruction sees of an fp32 GEMM operand: sign, 8 exponent bits, 10 mantissa
    bits.  mode "trunc": the low 13 bits are ignored; "rna": round to nearest, ties away from zero
    (cvt.rna.tf32.f32); "rne": round to nearest, ties to even.  Test-only model of the DEVICE arithmetic;
    the reference itself is fp32."""
    if mode is None:
        return a
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    if mode == "rna":
        u = u + np.uint32(0x1000)
    elif mode == "rne":
        u = u + np.uint32(0x0FFF) + ((u >> np.uint32(13)) & np.uint32(1))
    return (u & np.uint32(0xFFFFE000)).view(np.float32)


class CpuNnet:
    def __init__(self, config_text, seed=42, dtype=np.float32, backend=None, gemm_operands=None):
        """backend: None = C restatement (oracle.py); or an object with the same functions
        (oracle/ref.py wraps the compiled reference).
        gemm_operands: None = the reference's fp32 arithmetic; "trunc" / "rna" = the same op chain with the
        two operands of every matrix product reduced to TF32 first (tf32_operand) -- the arithmetic of the
        device's TF32 tensor-core path, so a test can hold that path to a rounding-error bound instead of
        a bound that has to absorb flipped ReLU gates and max-pool winners."""
        self.o = backend or ora
        self.ops = gemm_operands
        self.dtype = dtype
        rng = np.random.default_rng(seed)
        self.layers = []
        for kind, kv in parse_config(config_text):
            L = {"kind": kind}
            if kind == "ConvolutionComponent":
                g = lambda k, d=None: int(kv[k]) if k in kv else d
                L.update(H=g("in-height"), W=g("in-width"), C=g("in-channel"), ph=g("in-pad-height", 0),
                         pw=g("in-pad-width", 0), KH=g("kernel-height"), KW=g("kernel-width"), G=g("group"),
                         lr=float(kv["learning-rate"]))
                ps = float(kv.get("param-stddev", 1.0 / np.sqrt(L["KH"] * L["KW"])))
                bs = float(kv.get("bias-stddev", 1.0))
                L["lin"] = (rng.standard_normal((L["KH"] * L["KW"] * L["C"], L["G"])) * ps).astype(dtype)
                L["bias"] = (rng.standard_normal(L["G"]) * bs).astype(dtype)
                L["prev"] = np.zeros_like(L["lin"])
                L["wd"], L["mom"] = 0.0002, 0.9          # config values not applied
                L["OH"] = L["H"] + 2 * L["ph"] - L["KH"] + 1
                L["OW"] = L["W"] + 2 * L["pw"] - L["KW"] + 1
            elif kind == "MaxpoolComponent":
                g = lambda k: int(kv[k])
                L.update(H=g("in-height"), W=g("in-width"), C=g("in-channel"), ph=g("pool-height-dim"),
                         pw=g("pool-width-dim"), pc=g("pool-channel-dim"))
            elif kind == "FullyConnectedComponent":
                din, dout = int(kv["input-dim"]), int(kv["output-dim"])
                ps = float(kv.get("param-stddev", 1.0 / np.sqrt(din)))
                L.update(din=din, dout=dout, lr=float(kv.get("learning-rate", 0.001)),
                         wd=float(kv.get("weight-decay", 0.0002)), mom=float(kv.get("momentum", 0.9)))
                L["W"] = (rng.standard_normal((dout, din)) * ps).astype(dtype)
                L["bias"] = np.full(dout, float(kv.get("bias-stddev", 1.0)), dtype=dtype)
                L["prev"] = np.zeros_like(L["W"])
            elif kind == "DropoutComponent":
                L.update(dp=float(kv.get("dropout-proportion", 0.5)), scale=float(kv.get("dropout-scale", 0.0)))
            elif kind in ("RectifiedLinearComponent", "SoftmaxComponent"):
                L.update(dim=int(kv["dim"]))
            else:
                raise ValueError("unsupported component " + kind)
            self.layers.append(L)
        self.rng = rng

    @property
    def input_dim(self):
        L = self.layers[0]
        return L["H"] * L["W"] * L["C"] if L["kind"] == "ConvolutionComponent" else L["din"]

    def forward(self, x, dropout_masks=None):
        o, acts = self.o, [np.ascontiguousarray(x, dtype=self.dtype)]
        t = lambda m: tf32_operand(m, self.ops)
        self.masks = []
        for i, L in enumerate(self.layers):
            a, k = acts[-1], L["kind"]
            if k == "ConvolutionComponent":
                y = o.conv_propagate(t(a), t(L["lin"]), L["bias"], L["H"], L["W"], L["C"], L["ph"], L["pw"], L["KH"],
                                     L["KW"], L["G"], dtype=self.dtype)
            elif k == "MaxpoolComponent":
                y = o.maxpool_prop(a, L["H"], L["W"], L["ph"], L["pw"], L["pc"], dtype=self.dtype)
            elif k == "FullyConnectedComponent":
                y = o.fc_propagate(t(a), t(L["W"]), L["bias"], dtype=self.dtype)
            elif k == "RectifiedLinearComponent":
                y = np.maximum(a, 0)
            elif k == "DropoutComponent":
                if dropout_masks is not None:
                    m = dropout_masks[len(self.masks)]
                else:
                    hi = (1.0 - L["dp"] * L["scale"]) / (1.0 - L["dp"])
                    m = np.where(self.rng.random(a.shape) > L["dp"], hi, L["scale"]).astype(self.dtype)
                self.masks.append(m)
                y = a * m
            elif k == "SoftmaxComponent":
                e = np.exp(a - a.max(axis=1, keepdims=True))
                y = np.maximum(e / e.sum(axis=1, keepdims=True), 1e-20).astype(self.dtype)
            acts.append(y)
        self.acts = acts
        return acts[-1]

    def backward(self, labels, update=True):
        o = self.o
        t = lambda m: tf32_operand(m, self.ops)
        post = self.acts[-1]
        n = post.shape[0]
        objf = float(np.log(post[np.arange(n), labels].astype(np.float64)).sum())
        d = np.zeros_like(post)
        d[np.arange(n), labels] = 1.0 / post[np.arange(n), labels]
        mi = len(self.masks)
        for i in range(len(self.layers) - 1, -1, -1):
            L, a, y, k = self.layers[i], self.acts[i], self.acts[i + 1], self.layers[i]["kind"]
            if k == "ConvolutionComponent":
                din = o.conv_backprop(t(d), t(L["lin"]), L["H"], L["W"], L["C"], L["ph"], L["pw"], L["KH"], L["KW"],
                                      L["G"], dtype=self.dtype)
                if update:
                    args = (L["H"], L["W"], L["C"], L["ph"], L["pw"], L["KH"], L["KW"], L["G"], L["lr"], L["wd"],
                            L["mom"])
                    lin, bias, prev, _, _ = o.conv_update(t(a), t(d), L["lin"], L["bias"], L["prev"], *args,
                                                          dtype=self.dtype)
                    if self.ops is not None:     # the bias gradient is a column sum, not a matrix product
                        bias = o.conv_update(a, d, L["lin"], L["bias"], L["prev"], *args, dtype=self.dtype)[1]
                    L["lin"], L["bias"], L["prev"] = lin, bias, prev
            elif k == "MaxpoolComponent":
                din = o.maxpool_backprop(a, y, d, L["H"], L["W"], L["ph"], L["pw"], L["pc"], dtype=self.dtype)
            elif k == "FullyConnectedComponent":
                din = o.fc_backprop(t(d), t(L["W"]), dtype=self.dtype)
                if update:
                    Wn, bias, prev = o.fc_update(t(a), t(d), L["W"], L["bias"], L["prev"], L["lr"], L["wd"],
                                                 L["mom"], dtype=self.dtype)
                    if self.ops is not None:
                        bias = o.fc_update(a, d, L["W"], L["bias"], L["prev"], L["lr"], L["wd"], L["mom"],
                                           dtype=self.dtype)[1]
                    L["W"], L["bias"], L["prev"] = Wn, bias, prev
            elif k == "RectifiedLinearComponent":
                din = np.where(y > 0, d, 0).astype(self.dtype)
            elif k == "DropoutComponent":
                mi -= 1
                din = d * self.masks[mi]
            elif k == "SoftmaxComponent":
                dot = (y.astype(np.float64) * d).sum(axis=1, keepdims=True)
                din = (y * (d - dot)).astype(self.dtype)
            d = din
        self.input_deriv = d
        return objf

    def train_step(self, x, labels):
        self.forward(x)
        return self.backward(labels, update=True)
