"""Shared implementation behind the three reference layer files (NQ / CL / PI):
torch.autograd plumbing over the C-ABI kernels, and nn.Module hosts that keep the
reference's names, constructor arguments, attributes and parameter layouts
(dense `(in, units)`, conv HWIO, NHWC activations).

Reference: */nested_quantization_layer/custom_components/custom_layers.py (NQ),
*/custom_loss_terms/custom_components/custom_layers.py (CL),
CIFAR-10/paper_implementation/custom_components/custom_layers.py (PI).
"""
from __future__ import annotations

import re
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib, ops

eps_float32 = ops.EPS_FLOAT32          # NQ:11

SGRAD_NONE, SGRAD_THRESHOLD = _lib.LQ_SGRAD_NONE, _lib.LQ_SGRAD_THRESHOLD


class _FakeQuant(torch.autograd.Function):
    """my_custom_gradient: forward NQ:55-60; backward NQ:62-118 (threshold) or CL:61-62 / PI:62-64."""

    @staticmethod
    def forward(ctx, parameter, scale, penalty_threshold, sgrad_mode):
        P = ops.dense(parameter.detach())            # stays as it lies in memory (e.g. an OHWI-stored kernel)
        s = scale.detach().contiguous()
        layout = ops.tensor_layout(P, s.shape)
        y = ops.fq_fwd(P, s, layout)
        ctx.save_for_backward(P, s)
        ctx.layout, ctx.lam, ctx.mode, ctx.s_shape = layout, penalty_threshold, sgrad_mode, scale.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        P, s = ctx.saved_tensors
        dy = ops.match_strides(dy, P)
        if ctx.mode == SGRAD_THRESHOLD:
            dX, ds = ops.fq_bwd_threshold(dy, P, s, ctx.lam, ctx.layout)
        else:
            dX, ds = ops.fq_bwd_ste(dy, s)
        return dX, ds.view(ctx.s_shape), None, None


def fake_quant(parameter, scale, penalty_threshold, sgrad_mode):
    return _FakeQuant.apply(parameter, scale, 0.0 if penalty_threshold is None else float(penalty_threshold), sgrad_mode)


class MinValueConstraint:
    """Ensures the scale factor values stay above a defined minimum value (NQ:35-46)."""

    def __init__(self, min_value):
        self.min_value = min_value

    def __call__(self, w):
        out = w.detach().clone().contiguous()
        return ops.scale_constraint_(out, float(self.min_value))

    def apply_(self, w):
        """In-place form for use after an optimizer step."""
        return ops.scale_constraint_(w.data, float(self.min_value))

    def get_config(self):
        return {"min_value": self.min_value}


_UIDS: dict = {}


def _auto_name(cls_name: str) -> str:
    # Keras auto-naming: snake_case(class name) + ("" | "_N") -- the reference never forwards
    # name= to Keras (NQ:219), and its drivers filter layers on these auto-names.
    base = re.sub(r"(?<!^)(?=[A-Z])", "_", cls_name).lower().replace("2_d", "2d")
    n = _UIDS.get(base, 0)
    _UIDS[base] = n + 1
    return base if n == 0 else f"{base}_{n}"


def _as_param(value, shape, device) -> nn.Parameter:
    t = torch.as_tensor(value, dtype=torch.float32).reshape(tuple(shape)).clone()
    return nn.Parameter(t.to(device).contiguous())


class QuantizedScaleLayerBase(nn.Module):
    """CustomQuantizedScaleLayer (NQ:123-200 / CL:67-144 / PI:69-146)."""
    _sgrad_mode = SGRAD_THRESHOLD

    def __init__(self, initializer, orientation):
        super().__init__()
        self.initializer = initializer
        self.orientation = orientation
        self.built = False
        self.constraint = MinValueConstraint(eps_float32 * 100)
        self._banked = None        # set by training.step.QuantBank: pre-computed y for this step

    def _threshold(self) -> float:
        return 0.0

    def build(self, input_shape, device=None):
        shape = ops.scale_shape(tuple(input_shape), self.orientation)   # raises ValueError (NQ:194-197)
        self.scale = nn.Parameter(torch.full(shape, ops.SCALE_INIT, dtype=torch.float32, device=device))
        self.built = True

    def forward(self, inputs):
        if not self.built:
            self.build(tuple(inputs.shape), inputs.device)
        if self._banked is not None:
            # step-level multi-tensor path: y was produced by ONE lq_mt_fq_fwd launch over all
            # layers; its gradient is consumed by ONE lq_mt_fq_bwd launch (training/step.py)
            return self._banked
        return fake_quant(inputs, self.scale, self._threshold(), self._sgrad_mode)

    call = forward


def _same_pad(n: int, k: int, s: int) -> Tuple[int, int]:
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


class _ConvBias(torch.autograd.Function):
    """conv2d + bias on cuDNN with the bias gradient from lq_colsum: the output gradient of an NHWC
    convolution is a [N*H*W, Cout] matrix in memory and d(bias) its column sums; the framework's generic
    reduction took ~43 us per layer for it (13 % of the IMAGENETTE step)."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.conv = (tuple(stride), tuple(padding))
        return F.conv2d(x, w, bias, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.conv
        gx, gw, _ = torch.ops.aten.convolution_backward(
            gy, x, w, None, stride, padding, (1, 1), False, (0, 0), 1,
            (ctx.needs_input_grad[0], ctx.needs_input_grad[1], False))
        gb = None
        if ctx.needs_input_grad[2]:
            g = gy.permute(0, 2, 3, 1)                   # NHWC: contiguous when gy is channels-last
            if not g.is_contiguous():
                g = g.contiguous()
            gb = ops.colsum(g.reshape(-1, g.shape[-1]))
        return gx, gw, gb, None, None


def _conv(x, w, bias, stride, padding):
    if bias is None or not x.is_cuda:
        return F.conv2d(x, w, bias, stride=stride, padding=padding)
    return _ConvBias.apply(x, w, bias, stride, padding)


def conv2d_nhwc(inputs, kernel_hwio, strides, padding: str, bias=None):
    """tf.nn.conv2d semantics (NHWC activations, HWIO filters, TF 'SAME' padding) on cuDNN.
    `bias` is added in the convolution's epilogue: the same single fp32 add per
    output as NQ:350's tf.add(conv_output, qb), without the extra pass over the activation
    and with the bias gradient produced by cuDNN's backward instead of a separate reduction."""
    if isinstance(strides, int):
        strides = (strides, strides)
    sh, sw = int(strides[0]), int(strides[1])
    x = inputs.permute(0, 3, 1, 2)                       # NCHW view of channels-last memory
    w = kernel_hwio.permute(3, 2, 0, 1)                  # OIHW view
    if padding.upper() == "SAME":
        pt, pb = _same_pad(x.shape[2], w.shape[2], sh)
        pl, pr = _same_pad(x.shape[3], w.shape[3], sw)
        if pt == pb and pl == pr:
            y = _conv(x, w, bias, (sh, sw), (pt, pl))
        else:
            y = _conv(F.pad(x, (pl, pr, pt, pb)), w, bias, (sh, sw), (0, 0))
    elif padding.upper() == "VALID":
        y = _conv(x, w, bias, (sh, sw), (0, 0))
    else:
        raise ValueError(f"padding must be 'same' or 'valid', got {padding!r}")
    return y.permute(0, 2, 3, 1)


class _HostLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.name = _auto_name(type(self).__name__)
        self.built = False

    def _init_weight(self, shape, trained, device):
        if trained is not None:
            return _as_param(trained, shape, device)
        if self.initializer is None:
            # Keras' add_weight default for float weights: glorot_uniform with Keras' fans (_compute_fans): a vector
            # has fan_in = fan_out = n, a matrix (in, out), a kernel receptive_field * (Cin, Cout)
            shape = tuple(int(d) for d in shape)
            if len(shape) == 1:
                fan_in = fan_out = shape[0]
            elif len(shape) == 2:
                fan_in, fan_out = shape
            else:
                receptive = 1
                for d in shape[:-2]:
                    receptive *= d
                fan_in, fan_out = receptive * shape[-2], receptive * shape[-1]
            limit = (6.0 / max(1.0, (fan_in + fan_out))) ** 0.5
            t = torch.empty(shape, dtype=torch.float32).uniform_(-limit, limit)
            return nn.Parameter(t.to(device))
        return _as_param(self.initializer(tuple(shape)), shape, device)

    def regularization_loss(self):
        """Sum of regularizer(param) over this layer's parameters (Keras adds these to the loss)."""
        if self.regularizer is None:
            return None
        terms = [self.regularizer(p) for p in self._regularized()]
        return sum(terms[1:], terms[0])


class DenseLayerBase(_HostLayer):
    """CustomDenseLayer (NQ:203-268): y = x @ q(W) + q(b); W is (in, units), bias scale is 'scalar'."""
    _q_cls = QuantizedScaleLayerBase
    _q_arg = "penalty_threshold"

    def _setup(self, seed, units, q_value, orientation, initializer, regularizer, trained_weights):
        self.seed = seed
        self.nested_q_w_layer = self._q_cls(**{self._q_arg: q_value}, initializer=None, orientation=orientation)
        self.nested_q_b_layer = self._q_cls(**{self._q_arg: q_value}, initializer=None, orientation="scalar")
        self.units = units
        self.initializer = initializer
        self.regularizer = regularizer
        self.trained_weights = trained_weights

    def build(self, input_shape, device=None):
        tw = self.trained_weights if self.trained_weights else None      # NQ:240
        self.W = self._init_weight((input_shape[-1], self.units), tw[0] if tw else None, device)
        self.b = self._init_weight((self.units,), tw[1] if tw else None, device)
        self.nested_q_w_layer.build(tuple(self.W.shape), device)
        self.nested_q_b_layer.build(tuple(self.b.shape), device)
        self.built = True

    def _regularized(self):
        return [self.W, self.b]

    def forward(self, inputs):
        if not self.built:
            self.build(tuple(inputs.shape), inputs.device)
        qw = self.nested_q_w_layer(self.W)                               # NQ:265
        qb = self.nested_q_b_layer(self.b)                               # NQ:266
        return torch.add(torch.matmul(inputs, qw), qb)                   # NQ:268

    call = forward


class Conv2DLayerBase(_HostLayer):
    """CustomConv2DLayer (NQ:271-350) / CustomConv2DLayerNoBias (PI:299-365): HWIO kernel, NHWC input."""
    _q_cls = QuantizedScaleLayerBase
    _q_arg = "penalty_threshold"
    _use_bias = True

    def _setup(self, seed, q_value, orientation, initializer, filters, kernel_size, strides, padding,
               regularizer, trained_weights, kernel_storage: str = "hwio"):
        """kernel_storage: "hwio" keeps the kernel in memory as the reference's TensorFlow
        variable lies (kh, kw, Cin, Cout); "ohwi" keeps the SAME logical HWIO tensor -- same shape, same
        values, same groups for every orientation -- in (Cout, kh, kw, Cin) memory order, which is what
        cuDNN's channels-last convolutions read and write: the per-step transposes of every kernel and
        of every kernel gradient disappear.  The quantisation kernels work on memory order either way
        (ops.tensor_layout)."""
        if kernel_storage not in ("hwio", "ohwi"):
            raise ValueError("kernel_storage must be 'hwio' or 'ohwi'")
        self.kernel_storage = kernel_storage
        self.seed = seed
        self.nested_q_k_layer = self._q_cls(**{self._q_arg: q_value}, initializer=None, orientation=orientation)
        if self._use_bias:
            self.nested_q_b_layer = self._q_cls(**{self._q_arg: q_value}, initializer=None, orientation="scalar")
        self.initializer = initializer
        self.filters = filters
        self.kernel_size = tuple(kernel_size)
        self.strides = strides
        self.padding = padding.upper()                                   # NQ:305
        self.regularizer = regularizer
        self.trained_weights = trained_weights

    def build(self, input_shape, device=None):
        tw = self.trained_weights if self.trained_weights else None      # NQ:314
        kernel_shape = (*self.kernel_size, input_shape[-1], self.filters)  # NQ:321
        self.kernel = self._init_weight(kernel_shape, tw[0] if tw else None, device)
        if self.kernel_storage == "ohwi":
            ohwi = self.kernel.detach().permute(3, 0, 1, 2).contiguous()
            self.kernel = nn.Parameter(ohwi.permute(1, 2, 3, 0))          # logical HWIO, OHWI in memory
        self.nested_q_k_layer.build(tuple(self.kernel.shape), device)
        if self._use_bias:
            self.b = self._init_weight((self.filters,), tw[1] if tw else None, device)
            self.nested_q_b_layer.build(tuple(self.b.shape), device)
        self.built = True

    def _regularized(self):
        return [self.kernel] + ([self.b] if self._use_bias else [])

    def forward(self, inputs):
        if not self.built:
            self.build(tuple(inputs.shape), inputs.device)
        qk = self.nested_q_k_layer(self.kernel)                          # NQ:340
        if not self._use_bias:
            return conv2d_nhwc(inputs, qk, self.strides, self.padding)   # PI:356-365
        qb = self.nested_q_b_layer(self.b)                               # NQ:341
        return conv2d_nhwc(inputs, qk, self.strides, self.padding, bias=qb)  # NQ:343-350 (add fused)

    call = forward


def apply_scale_constraints(module: nn.Module) -> None:
    """What Keras does after every optimizer step: s <- max(s, 100*eps) (NQ:158 + NQ:35-46)."""
    for m in module.modules():
        if isinstance(m, QuantizedScaleLayerBase) and m.built:
            m.constraint.apply_(m.scale)
