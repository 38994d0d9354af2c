"""A torch-CPU stand-in for the few `tensorflow` symbols the reference's hot-path files touch.

Purpose: TensorFlow 2.11 (requirements.txt:1 of the reference) is not installable in
this container, so `make_golden.py` installs this module under the names
`tensorflow`, `tensorflow.keras.*` and then imports the reference's OWN source files
(custom_layers.py, custom_loss_functions.py) unmodified from /root/reference.  Every
tf.* call in those files lands here and is executed by the equivalent torch CPU fp32
op; gradients the reference gets from TF autodiff come from torch autograd, whose
rules for abs / amax (even tie split) / div / mean / where agree with TF's
math_grad.py.  IEEE fp32 add/sub/mul/div/floor/compare are correctly rounded in both
libraries, so forward codes and outputs are bit-identical to any conforming backend;
tanh and summation order are implementation details covered by the 1e-5 tolerance.

This file is test infrastructure.  It implements semantics from the public TF API
documentation; it contains no reference code.
"""
from __future__ import annotations

import logging
import re
import sys
import types

import numpy as np
import torch

torch.set_grad_enabled(True)

TRACE = {"floor": []}   # make_golden.py reads the integer codes out of here


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return (int(axis),)


# ----------------------------------------------------------------------- tf.* ops
def floor(x):
    out = torch.floor(_t(x))
    TRACE["floor"].append(out.detach().clone())
    return out


def stop_gradient(x):
    return _t(x).detach()


def where(cond, a, b):
    return torch.where(_t(cond), _t(a), _t(b))


def abs_(x):
    return torch.abs(_t(x))


def maximum(a, b):
    return torch.maximum(_t(a), _t(b))


def reduce_max(x, axis=None):
    x = _t(x)
    ax = _axes(axis)
    return torch.amax(x) if ax is None else torch.amax(x, dim=ax)


def reduce_all(x, axis=None):
    x = _t(x)
    ax = _axes(axis)
    if ax is None:
        return torch.all(x)
    for a in sorted((a % x.ndim for a in ax), reverse=True):
        x = torch.all(x, dim=a)
    return x


def reduce_mean(x, axis=None):
    x = _t(x)
    ax = _axes(axis)
    return torch.mean(x) if ax is None else torch.mean(x, dim=ax)


def expand_dims(x, axis):
    x = _t(x)
    ax = _axes(axis)
    for a in sorted(ax):
        x = torch.unsqueeze(x, a)
    return x


def broadcast_to(x, shape):
    return torch.broadcast_to(_t(x), tuple(shape))


def reshape(x, shape):
    return torch.reshape(_t(x), tuple(shape))


def zeros_like(x):
    return torch.zeros_like(_t(x))


def tanh(x):
    return torch.tanh(_t(x))


def matmul(a, b):
    return torch.matmul(_t(a), _t(b))


def add(a, b):
    return torch.add(_t(a), _t(b))


def tf_print(*a, **k):
    return None


def get_logger():
    return logging.getLogger("tf_shim")


def Variable(value, trainable=True, dtype=None, **k):
    return torch.tensor(float(value), dtype=torch.float32, requires_grad=bool(trainable))


def constant_initializer(value):
    arr = np.asarray(value, dtype=np.float32)

    def init(shape):
        return torch.as_tensor(arr.reshape(tuple(shape)).copy())
    return init


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return total // 2, total - total // 2


def conv2d(inputs, filters, strides, padding):
    """tf.nn.conv2d: NHWC input, HWIO filters, 'SAME' (TF asymmetric) or 'VALID'."""
    x = _t(inputs).permute(0, 3, 1, 2)
    w = _t(filters).permute(3, 2, 0, 1)
    if isinstance(strides, int):
        strides = (strides, strides)
    sh, sw = int(strides[0]), int(strides[1])
    if padding.upper() == "SAME":
        pt, pb = _same_pad(x.shape[2], w.shape[2], sh)
        pl, pr = _same_pad(x.shape[3], w.shape[3], sw)
        x = torch.nn.functional.pad(x, (pl, pr, pt, pb))
    y = torch.nn.functional.conv2d(x, w, stride=(sh, sw))
    return y.permute(0, 2, 3, 1)


def custom_gradient(f):
    def wrapper(*args):
        idx = [i for i, a in enumerate(args) if isinstance(a, torch.Tensor)]

        class _Fn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, *tensors):
                full = list(args)
                for i, t in zip(idx, tensors):
                    full[i] = t
                y, grad_fn = f(*full)
                ctx.user_grad = grad_fn
                return y.clone()

            @staticmethod
            def backward(ctx, dy):
                grads = ctx.user_grad(dy)
                return tuple(grads[i] for i in idx)

        return _Fn.apply(*[args[i] for i in idx])
    return wrapper


def sparse_categorical_crossentropy(y_true, y_pred):
    y_pred = _t(y_pred)
    idx = torch.as_tensor(np.asarray(y_true)).long().reshape(-1, 1)
    p = torch.gather(y_pred, 1, idx).squeeze(1)
    return -torch.log(torch.clamp(p, 1e-7, 1.0))


# ------------------------------------------------------------------------- keras
class Constraint:
    pass


class Initializer:
    pass


class Regularizer:
    pass


class Constant(Initializer):
    def __init__(self, value):
        self.value = value

    def __call__(self, shape):
        return torch.full(tuple(shape), float(np.float32(self.value)), dtype=torch.float32)


class RandomNormal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev
        self.gen = torch.Generator().manual_seed(0 if seed is None else int(seed))

    def __call__(self, shape):
        return torch.randn(tuple(shape), generator=self.gen) * self.stddev + self.mean


class Layer:
    _uids: dict = {}

    def __init__(self, name=None, **kwargs):
        if name is None:
            base = re.sub(r"(?<!^)(?=[A-Z])", "_", type(self).__name__).lower()
            base = base.replace("2_d", "2d")
            n = Layer._uids.get(base, 0)
            Layer._uids[base] = n + 1
            name = base if n == 0 else f"{base}_{n}"
        self.name = name
        self.built = False

    def add_weight(self, name=None, shape=None, initializer=None, trainable=True,
                   constraint=None, regularizer=None, **k):
        w = initializer(tuple(shape)) if shape != () else torch.tensor(float(initializer.value))
        w = w.clone().detach().to(torch.float32).requires_grad_(bool(trainable))
        w.constraint = constraint
        return w

    def __call__(self, inputs):
        if not self.built:
            if hasattr(self, "build"):
                self.build(tuple(inputs.shape))
            self.built = True
        return self.call(inputs)


def install():
    """Register the shim under the module names the reference imports."""
    tf = types.ModuleType("tensorflow")
    tf.floor, tf.stop_gradient, tf.where, tf.abs = floor, stop_gradient, where, abs_
    tf.maximum, tf.reduce_max, tf.reduce_all, tf.reduce_mean = maximum, reduce_max, reduce_all, reduce_mean
    tf.expand_dims, tf.broadcast_to, tf.reshape, tf.zeros_like = expand_dims, broadcast_to, reshape, zeros_like
    tf.matmul, tf.add, tf.print, tf.get_logger = matmul, add, tf_print, get_logger
    tf.Variable, tf.constant_initializer, tf.custom_gradient = Variable, constant_initializer, custom_gradient
    tf.float32 = torch.float32

    math = types.ModuleType("tensorflow.math")
    math.tanh = tanh
    nn = types.ModuleType("tensorflow.nn")
    nn.conv2d = conv2d

    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer = Layer
    constraints = types.ModuleType("tensorflow.keras.constraints")
    constraints.Constraint = Constraint
    initializers = types.ModuleType("tensorflow.keras.initializers")
    initializers.Initializer, initializers.RandomNormal, initializers.Constant = Initializer, RandomNormal, Constant
    regularizers = types.ModuleType("tensorflow.keras.regularizers")
    regularizers.Regularizer = Regularizer
    losses = types.ModuleType("tensorflow.keras.losses")
    losses.sparse_categorical_crossentropy = sparse_categorical_crossentropy

    keras.layers, keras.constraints, keras.initializers = layers, constraints, initializers
    keras.regularizers, keras.losses = regularizers, losses
    tf.math, tf.nn, tf.keras = math, nn, keras

    for name, mod in {
        "tensorflow": tf, "tensorflow.math": math, "tensorflow.nn": nn, "tensorflow.keras": keras,
        "tensorflow.keras.layers": layers, "tensorflow.keras.constraints": constraints,
        "tensorflow.keras.initializers": initializers, "tensorflow.keras.regularizers": regularizers,
        "tensorflow.keras.losses": losses,
    }.items():
        sys.modules[name] = mod
    return tf
