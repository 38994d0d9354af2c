"""Drop-in for CIFAR-10/paper_implementation/custom_components/custom_layers.py (PI).

Forward PI:56-60 -> lq_fq_fwd.  The reference's backward returns zeros for the scale
(PI:62-64, marked TODO there); that is the default here.  `set_scale_grad_mode("threshold")`
switches every layer created afterwards to the threshold rule of the nested variant
(NQ:62-118), which is what the reference's TODO announces and what the batch-256
per-channel benchmark configuration of BASELINE.json uses.
"""
from __future__ import annotations

from typing import Literal, Tuple

from ...custom_components import _impl
from ...custom_components._impl import MinValueConstraint, apply_scale_constraints  # noqa: F401

eps_float32 = _impl.eps_float32
_MODE = {"mode": _impl.SGRAD_NONE}


def set_scale_grad_mode(mode: str) -> None:
    if mode not in ("none", "threshold"):
        raise ValueError("scale_grad_mode must be 'none' (reference) or 'threshold'")
    _MODE["mode"] = _impl.SGRAD_THRESHOLD if mode == "threshold" else _impl.SGRAD_NONE


def my_custom_gradient(parameter, scale, penalty_threshold):
    return _impl.fake_quant(parameter, scale, penalty_threshold, _MODE["mode"])


class CustomQuantizedScaleLayer(_impl.QuantizedScaleLayerBase):
    """PI:69-146."""

    def __init__(self, penalty_threshold, initializer, orientation):
        super().__init__(initializer, orientation)
        self.penalty_threshold = penalty_threshold
        self._sgrad_mode = _MODE["mode"]

    def _threshold(self):
        return self.penalty_threshold


class CustomDenseLayer(_impl.DenseLayerBase):
    """PI:149-214."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_threshold"

    def __init__(self, seed: int, units: int, penalty_threshold: float,
                 orientation: Literal["rowwise", "columnwise", "scalar"], initializer, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, units, penalty_threshold, orientation, initializer, regularizer, trained_weights)


class CustomConv2DLayer(_impl.Conv2DLayerBase):
    """PI:217-296."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_threshold"

    def __init__(self, seed: int, penalty_threshold: float,
                 orientation: Literal["rowwise", "columnwise", "channelwise", "scalar"], initializer,
                 filters: int, kernel_size: Tuple[int, int], strides: Tuple[int, int], padding: str, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, penalty_threshold, orientation, initializer, filters, kernel_size, strides, padding,
                    regularizer, trained_weights, kwargs.get("kernel_storage", "hwio"))


class CustomConv2DLayerNoBias(_impl.Conv2DLayerBase):
    """PI:299-365."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_threshold"
    _use_bias = False

    def __init__(self, seed: int, penalty_threshold: float,
                 orientation: Literal["rowwise", "columnwise", "channelwise", "scalar"], initializer,
                 filters: int, kernel_size: Tuple[int, int], strides: Tuple[int, int], padding: str, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, penalty_threshold, orientation, initializer, filters, kernel_size, strides, padding,
                    regularizer, trained_weights, kwargs.get("kernel_storage", "hwio"))
