"""Drop-in for */nested_quantization_layer/custom_components/custom_layers.py (NQ).

Same names and constructor arguments as the reference; the arithmetic runs in liblq.so:
forward NQ:55-60 -> lq_fq_fwd, backward NQ:62-118 -> lq_fq_bwd_threshold.
"""
from __future__ import annotations

from typing import Literal, Optional, Tuple

from ...custom_components import _impl
from ...custom_components._impl import MinValueConstraint, apply_scale_constraints  # noqa: F401

eps_float32 = _impl.eps_float32


def my_custom_gradient(parameter, scale, penalty_threshold):
    """Fake-quantise `parameter` with `scale`; the backward penalises large |grad|/|param|
    ratios through the learned-scale gradient (NQ:49-120)."""
    return _impl.fake_quant(parameter, scale, penalty_threshold, _impl.SGRAD_THRESHOLD)


class CustomQuantizedScaleLayer(_impl.QuantizedScaleLayerBase):
    """NQ:123-200."""
    _sgrad_mode = _impl.SGRAD_THRESHOLD

    def __init__(self, penalty_threshold, initializer, orientation):
        super().__init__(initializer, orientation)
        self.penalty_threshold = penalty_threshold

    def _threshold(self):
        return self.penalty_threshold


class CustomDenseLayer(_impl.DenseLayerBase):
    """NQ:203-268."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_threshold"

    def __init__(self, seed: int, units: int, penalty_threshold: float,
                 orientation: Literal["rowwise", "columnwise", "scalar"], initializer, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, units, penalty_threshold, orientation, initializer, regularizer, trained_weights)


class CustomConv2DLayer(_impl.Conv2DLayerBase):
    """NQ:271-350."""
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
