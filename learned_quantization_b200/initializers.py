"""Minimal stand-ins for the Keras initializers the reference's drivers pass to the layers
(`RandomNormal(seed=seed)`, experiment.py:637,717; `Constant`, NQ:156).  A layer accepts any
callable `f(shape) -> array-like`."""
from __future__ import annotations

import torch


class RandomNormal:
    def __init__(self, mean: float = 0.0, stddev: float = 0.05, seed=None):
        self.mean, self.stddev, self.seed = mean, stddev, seed
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(int(seed))

    def __call__(self, shape):
        return torch.randn(tuple(shape), generator=self._gen, dtype=torch.float32) * self.stddev + self.mean


class Constant:
    def __init__(self, value: float):
        self.value = float(value)

    def __call__(self, shape):
        return torch.full(tuple(shape), self.value, dtype=torch.float32)
