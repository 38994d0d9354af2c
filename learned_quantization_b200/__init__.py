"""learned_quantization_b200 -- the learned-quantization hot path of
anuunchin/learned-quantization on B200 (sm_100a): fake-quant forward, straight-through /
threshold backward with the learned-scale gradient, and the custom loss-term reductions,
behind the reference's layer and loss API.

Importing this package loads csrc/liblq.so and fails loudly if it has not been built;
there is no CPU or framework fallback for the hot path.

Drop-in modules (same class / function names as the reference trees):
    learned_quantization_b200.nested_quantization_layer.custom_components.custom_layers
    learned_quantization_b200.custom_loss_terms.custom_components.custom_layers
    learned_quantization_b200.custom_loss_terms.custom_components.custom_loss_functions
    learned_quantization_b200.paper_implementation.custom_components.custom_layers
"""
from . import _lib, ops  # noqa: F401  (loads liblq.so)
from ._lib import LqError, version  # noqa: F401
from . import regularizers  # noqa: F401
from .initializers import Constant, RandomNormal  # noqa: F401

__all__ = ["ops", "LqError", "version", "RandomNormal", "Constant"]
