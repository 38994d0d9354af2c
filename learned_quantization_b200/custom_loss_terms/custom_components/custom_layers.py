"""Drop-in for */custom_loss_terms/custom_components/custom_layers.py (CL).

Forward CL:55-59 -> lq_fq_fwd; backward CL:61-62 is the plain straight-through estimator
(dP = dy, dscale = 0) -> lq_fq_bwd_ste.  The scale gradients come from the loss terms
(custom_loss_functions.py).
"""
from __future__ import annotations

from typing import Literal, Tuple

from ...custom_components import _impl
from ...custom_components._impl import MinValueConstraint, apply_scale_constraints  # noqa: F401

eps_float32 = _impl.eps_float32


def my_custom_gradient(parameter, scale):
    """Custom gradient function that simply enables STE (CL:49-64)."""
    return _impl.fake_quant(parameter, scale, None, _impl.SGRAD_NONE)


class CustomQuantizedScaleLayer(_impl.QuantizedScaleLayerBase):
    """CL:67-144 (the MNIST copy keeps penalty_rate as a non-trainable weight, MNIST CL:97-102)."""
    _sgrad_mode = _impl.SGRAD_NONE

    def __init__(self, penalty_rate, initializer, orientation):
        super().__init__(initializer, orientation)
        self.penalty_rate = penalty_rate


class CustomDenseLayer(_impl.DenseLayerBase):
    """CL:147-212."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_rate"

    def __init__(self, seed: int, units: int, penalty_rate: float,
                 orientation: Literal["rowwise", "columnwise", "scalar"], initializer, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, units, penalty_rate, orientation, initializer, regularizer, trained_weights)


class CustomConv2DLayer(_impl.Conv2DLayerBase):
    """CL:215-294."""
    _q_cls = CustomQuantizedScaleLayer
    _q_arg = "penalty_rate"

    def __init__(self, seed: int, penalty_rate: float,
                 orientation: Literal["rowwise", "columnwise", "channelwise", "scalar"], initializer,
                 filters: int, kernel_size: Tuple[int, int], strides: Tuple[int, int], padding: str, name: str,
                 regularizer=None, trained_weights=None, **kwargs):
        super().__init__()
        self.given_name = name
        self._setup(seed, penalty_rate, orientation, initializer, filters, kernel_size, strides, padding,
                    regularizer, trained_weights, kwargs.get("kernel_storage", "hwio"))
