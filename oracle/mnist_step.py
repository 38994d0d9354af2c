"""CPU restatement of ONE training step of the reference's MNIST model -- TEST INFRASTRUCTURE ONLY
(imported by tests/ and by bench.py's `mnist_cpu` leg; never by the product).

BASELINE.json configs[0]: Flatten -> CustomDense(128) -> relu -> CustomDense(10) -> softmax
(MNIST/nested_quantization_layer/experiment.py:121-173), batch 32, Adam(1e-4)
(experiment.py:303-305), sparse categorical cross-entropy, on top of the operator restated in
oracle/lq_oracle.py:
    variant "nested"       forward NQ:55-60, backward NQ:62-118 (threshold rule), lambda
    variant "maxbin" / "difference" / "inverse"
                           forward CL:55-59, backward CL:61-62 (ds = 0) plus gamma * penalty of
                           MNIST/custom_loss_terms/custom_components/custom_loss_functions.py
                           (LF:75-116 / 161-195 / 240-275) differentiated with TF's autodiff rules
followed by Keras-2.11 Adam and MinValueConstraint (NQ:35-46) on the scales.
"PARITY UNPINNED" against TensorFlow binaries, like the rest of the oracle (see lq_oracle.py);
the matmul / softmax / cross-entropy part is ordinary numpy fp32.
"""
from __future__ import annotations

import numpy as np

from . import lq_oracle as o

F32 = np.float32
NAMES = ("W1", "b1", "W2", "b2")


class MnistOracle:
    def __init__(self, P: dict, S: dict, *, variant: str = "nested", value: float = 1e-10, lr: float = 1e-4):
        """P / S: parameters and scales by name (W1 (784,128), b1 (128,), W2 (128,10), b2 (10,)).
        value: lambda (nested) or gamma (loss-term variants)."""
        if variant not in ("nested", "maxbin", "difference", "inverse"):
            raise ValueError(variant)
        self.P = {k: np.array(P[k], F32) for k in NAMES}
        self.S = {k: np.array(S[k], F32) for k in NAMES}
        self.variant, self.value, self.lr = variant, float(value), float(lr)
        self.mP = {k: np.zeros_like(v) for k, v in self.P.items()}
        self.vP = {k: np.zeros_like(v) for k, v in self.P.items()}
        self.mS = {k: np.zeros_like(v) for k, v in self.S.items()}
        self.vS = {k: np.zeros_like(v) for k, v in self.S.items()}
        self.it = 0

    @classmethod
    def init(cls, seed: int = 42, orientation: str = "scalar", **kw):
        """RandomNormal(0, 0.05) for weights AND biases (experiment.py:131-139 hands one initializer to the
        layer, which uses it for both, NQ:246-257); every scale starts at 100*eps (NQ:156-158)."""
        rng = np.random.default_rng(seed)
        P = {"W1": rng.normal(0, 0.05, (784, 128)), "b1": rng.normal(0, 0.05, (128,)),
             "W2": rng.normal(0, 0.05, (128, 10)), "b2": rng.normal(0, 0.05, (10,))}
        S = {k: np.full(o.scale_shape(np.shape(P[k]), orientation if k[0] == "W" else "scalar"), o.SCALE_INIT, F32)
             for k in NAMES}
        return cls(P, S, **kw)

    def step(self, X: np.ndarray, Y: np.ndarray) -> float:
        """X: (B, 784) fp32, Y: (B,) int labels.  Returns the loss the step was taken on."""
        P, S = self.P, self.S
        B = X.shape[0]
        self.it += 1
        Q = {k: o.fq_forward(P[k], S[k]) for k in NAMES}                      # (codes, fake-quantised)
        yq = {k: Q[k][1] for k in NAMES}
        h = X @ yq["W1"] + yq["b1"]
        a = np.maximum(h, 0)
        z = a @ yq["W2"] + yq["b2"]
        e = np.exp(z - z.max(axis=1, keepdims=True))
        p = e / e.sum(axis=1, keepdims=True)
        loss = float(np.mean(-np.log(np.clip(p[np.arange(B), Y], 1e-7, 1 - 1e-7))))
        dz = p.copy()
        dz[np.arange(B), Y] -= 1
        dz = (dz / B).astype(F32)
        g = {"W2": a.T @ dz, "b2": dz.sum(0)}
        dh = (dz @ yq["W2"].T) * (h > 0)
        g["W1"], g["b1"] = X.T @ dh, dh.sum(0)
        gs = {}
        if self.variant == "nested":
            for k in NAMES:
                _, ds = o.nested_backward(g[k].astype(F32), Q[k][0], Q[k][1], S[k], self.value, accum="fp64")
                gs[k] = np.asarray(ds, F32).reshape(S[k].shape)
        else:
            layers = [(P["W1"], S["W1"], P["b1"], S["b1"]), (P["W2"], S["W2"], P["b2"], S["b2"])]
            pen = {"maxbin": o.maxbin_penalty, "difference": o.difference_penalty, "inverse": o.inverse_penalty}[self.variant]
            grads = {"maxbin": o.maxbin_penalty_grads, "difference": o.difference_penalty_grads,
                     "inverse": o.inverse_penalty_grads}[self.variant](layers, upstream=self.value, accum="fp64")
            loss += self.value * float(pen(layers))                            # LF:58: scce + gamma * penalty
            for (kW, kb), (gK, gsK, gb, gsb) in zip((("W1", "b1"), ("W2", "b2")), grads):
                g[kW] = g[kW] + gK
                g[kb] = g[kb] + gb
                gs[kW], gs[kb] = gsK.reshape(S[kW].shape), gsb.reshape(S[kb].shape)   # op's own ds is zeros (CL:61-62)
        for k in NAMES:
            P[k], self.mP[k], self.vP[k] = o.adam_step(P[k], g[k].astype(F32), self.mP[k], self.vP[k], self.lr, self.it)
            S[k], self.mS[k], self.vS[k] = o.adam_step(S[k], gs[k], self.mS[k], self.vS[k], self.lr, self.it,
                                                       min_value=o.SCALE_INIT)
        return loss
