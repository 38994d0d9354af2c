"""Stand-in for keras.regularizers.l2 as the reference's drivers use it
(`regularizers.l2(1e-4)`, IMAGENETTE/nested_quantization_layer/experiment.py:638)."""
from __future__ import annotations


class L2:
    def __init__(self, l2: float = 0.01):
        self.l2 = float(l2)

    def __call__(self, x):
        return self.l2 * (x * x).sum()

    def get_config(self):
        return {"l2": self.l2}


def l2(l2: float = 0.01) -> L2:   # noqa: A001  (Keras name)
    return L2(l2)
