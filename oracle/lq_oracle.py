"""CPU oracle for the learned-quantization hot path  --  TEST INFRASTRUCTURE ONLY.

This file restates, op for op and in the reference's own order, the arithmetic of
anuunchin/learned-quantization's quantization layers and custom loss terms in numpy
fp32.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import it.  The product (learned_quantization_b200/) never does: it calls the
CUDA C-ABI library and fails loudly when that library is missing.

PARITY STATUS: the reference has no tests, golden vectors or fixtures for this path
 and its arithmetic lives in TensorFlow 2.11, which is not
installed here, so parity against TensorFlow *binaries* is UNPINNED.  What pins this
oracle instead: tests/golden/make_golden.py imports the reference's own
custom_layers.py / custom_loss_functions.py unmodified from /root/reference and
executes them on a torch-CPU shim of the handful of tf.* calls they make; the
committed fixtures under tests/golden/ are those outputs and tests/test_oracle.py
checks this file against every one of them (plus the SURVEY Appendix B known-answer
vector).

Citations (relative to the reference root):
  NQ = IMAGENETTE/nested_quantization_layer/custom_components/custom_layers.py
  CL = IMAGENETTE/custom_loss_terms/custom_components/custom_layers.py
  PI = CIFAR-10/paper_implementation/custom_components/custom_layers.py
  LF = IMAGENETTE/custom_loss_terms/custom_components/custom_loss_functions.py
TF gradient rules restated for the loss terms (tensorflow/python/ops/math_grad.py,
TF 2.11, not vendored in the reference): _AbsGrad (grad*sign), _MinOrMaxGrad (even
split over ties), _RealDivGrad (grad/y ; sum(grad*((-x/y)/y)) over broadcast axes),
_MeanGrad (grad/N), _SelectGrad.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
# NQ:11  eps_float32 = np.finfo(np.float32).eps
EPS_FLOAT32 = np.finfo(np.float32).eps
# NQ:156-158  Constant(eps_float32 * 100) initialiser and MinValueConstraint floor
SCALE_INIT = F32(EPS_FLOAT32 * 100)

ROUND_FLOOR = 0   # the reference (NQ:59)
ROUND_RNE = 1     # extension: round-half-to-even (BASELINE.json wording)


# ----------------------------------------------------------------------------------
# a4  CustomQuantizedScaleLayer.build  (NQ:147-197)
# ----------------------------------------------------------------------------------
def scale_shape(param_shape, orientation):
    """Shape of the learned scale for a parameter of `param_shape` (NQ:149-197)."""
    n = len(param_shape)
    if orientation == "rowwise":        # NQ:149-152
        return tuple(param_shape[i] if i == 0 else 1 for i in range(n))
    if orientation == "columnwise":     # NQ:161-164
        return tuple(param_shape[i] if i == 1 else 1 for i in range(n))
    if orientation == "channelwise":    # NQ:173-176
        return tuple(param_shape[i] if i == 2 else 1 for i in range(n))
    if orientation == "scalar":         # NQ:185-188
        return (1,)
    raise ValueError(                   # NQ:194-197 (message kept verbatim)
        f"Invalid scaler application: {orientation}. Expected rowwise, columnwise or scalar."
    )


def group_layout(param_shape, s_shape):
    """View P as [outer, C, inner] with group(idx) = (idx // inner) % C  (SURVEY 8a).

    A scale has at most one non-1 axis (NQ:149-183) or is shape (1,) (NQ:185-192).
    """
    param_shape = tuple(int(d) for d in param_shape)
    s_shape = tuple(int(d) for d in s_shape)
    total = int(np.prod(param_shape, dtype=np.int64)) if param_shape else 1
    if len(s_shape) == 1 and len(param_shape) != 1:
        if s_shape[0] != 1:
            raise ValueError("rank-1 scale on a rank-%d parameter must be (1,)" % len(param_shape))
        return 1, 1, total
    if len(s_shape) != len(param_shape):
        raise ValueError("scale rank %d does not broadcast over parameter rank %d"
                         % (len(s_shape), len(param_shape)))
    axes = [i for i, d in enumerate(s_shape) if d != 1]
    if len(axes) > 1:
        raise ValueError("scale may have at most one non-1 axis")
    if not axes:
        return 1, 1, total
    ax = axes[0]
    if s_shape[ax] != param_shape[ax]:
        raise ValueError("scale axis %d has %d entries, parameter has %d"
                         % (ax, s_shape[ax], param_shape[ax]))
    outer = int(np.prod(param_shape[:ax], dtype=np.int64)) if ax > 0 else 1
    inner = int(np.prod(param_shape[ax + 1:], dtype=np.int64)) if ax + 1 < len(param_shape) else 1
    return outer, param_shape[ax], inner


# ----------------------------------------------------------------------------------
# a5  MinValueConstraint.__call__  (NQ:35-46)
# ----------------------------------------------------------------------------------
def min_value_constraint(w, min_value=SCALE_INIT):
    return np.maximum(np.asarray(w, F32), F32(min_value))      # NQ:43


# ----------------------------------------------------------------------------------
# a1  my_custom_gradient forward  (NQ:55-60 = CL:55-59 = PI:56-60)
# ----------------------------------------------------------------------------------
def fq_forward(P, s, round_mode=ROUND_FLOOR, qmin=None, qmax=None):
    """Returns (q, y): integer-valued fp32 codes and the fake-quantised parameter."""
    P = np.asarray(P, F32)
    s = np.asarray(s, F32)
    nonrounded = P / s                                         # NQ:56-58  tf.truediv, IEEE RN
    if round_mode == ROUND_FLOOR:
        q = np.floor(nonrounded)                               # NQ:59
    elif round_mode == ROUND_RNE:
        q = np.rint(nonrounded)                                # extension
    else:
        raise ValueError("bad round_mode")
    if qmin is not None or qmax is not None:                   # extension
        lo = F32(-np.inf if qmin is None else qmin)
        hi = F32(np.inf if qmax is None else qmax)
        q = np.minimum(np.maximum(q, lo), hi)
    y = q * s                                                  # NQ:60
    return q.astype(F32), y.astype(F32)


# ----------------------------------------------------------------------------------
# a2  custom_grad, nested variant  (NQ:62-118)
# ----------------------------------------------------------------------------------
def _mean(x, axis, accum):
    if accum == "fp64":
        return np.mean(x.astype(np.float64), axis=axis).astype(F32)
    return np.mean(x, axis=axis, dtype=F32)


def nested_backward(dy, q, y, s, penalty_threshold, accum="fp32"):
    """(dP, ds) of NQ:62-118.  accum="fp64" accumulates the mean in double (the
    'truth' the 1e-5 tolerance is measured against); "fp32" is the literal restatement."""
    dy = np.asarray(dy, F32)
    s = np.asarray(s, F32)
    lam = F32(penalty_threshold)                               # NQ:55 (python float -> fp32 tensor)
    non_zero_param = np.where(y == F32(0.0), F32(EPS_FLOAT32), y)          # NQ:63
    with np.errstate(all="ignore"):
        ratio = np.abs(dy) / np.abs(non_zero_param)                        # NQ:64
        c0 = F32(-1.0) * np.abs(np.tanh(lam))                              # NQ:79 / NQ:105
        if s.ndim == 1:                                                    # NQ:67
            maxvalue = np.max(np.abs(q))                                   # NQ:68
            all_above = np.all(ratio >= lam)                               # NQ:70
            all_above_broad = np.broadcast_to(np.expand_dims(all_above, -1), dy.shape)  # NQ:71-73
            scale_grads = np.where(                                        # NQ:77-85
                all_above_broad, c0,
                np.where(ratio >= lam, F32(0.0), F32(-1.0) * np.abs(np.tanh(lam - ratio))))
            reduced = _mean(scale_grads.astype(F32), None, accum)          # NQ:87
            reduced = np.reshape(reduced, s.shape)                         # NQ:88
        else:                                                              # NQ:91
            axes = tuple(i for i in range(s.ndim) if s.shape[i] == 1)      # NQ:92
            maxvalue = np.max(np.abs(q), axis=axes)                        # NQ:94
            maxvalue = np.reshape(maxvalue, s.shape)                       # NQ:95
            all_above = np.all(ratio >= lam, axis=axes)                    # NQ:97
            all_above_broad = np.broadcast_to(np.reshape(all_above, s.shape), dy.shape)  # NQ:98-99
            scale_grads = np.where(                                        # NQ:103-111
                all_above_broad, c0,
                np.where(ratio >= lam, F32(0.0), F32(-1.0) * np.abs(np.tanh(lam - ratio))))
            reduced = _mean(scale_grads.astype(F32), axes, accum)          # NQ:113
            reduced = np.reshape(reduced, s.shape)                         # NQ:114
        final = (reduced * maxvalue).astype(F32)                           # NQ:116
    return dy, final                                                       # NQ:118 (dP is dy itself)


# ----------------------------------------------------------------------------------
# a3  custom_grad, STE-only variant  (CL:61-62, PI:62-64)
# ----------------------------------------------------------------------------------
def ste_backward(dy, s):
    return np.asarray(dy, F32), np.zeros_like(np.asarray(s, F32))


# ----------------------------------------------------------------------------------
# a7-a9  custom loss terms (LF:75-116, 161-195, 240-275) and their TF-autodiff grads
# A "layer" here is a 4-tuple (K, sK, b, sb) of numpy arrays -- the four attributes
# the loss reads from each layer object (LF:85-88).
# ----------------------------------------------------------------------------------
def _dim(shape):
    d = 1.0                                                    # LF:102-108 (python float product)
    for n in shape:
        d *= n
    return d


def _maxbin_axes(s):
    # LF:90 / LF:96
    return tuple(i for i in range(s.ndim) if s.shape[i] == 1 and s.ndim > 1)


def _maxbin_of(P, s):
    axes = _maxbin_axes(s)
    d = np.abs(P) / s
    if axes != ():                                             # LF:91-92
        return np.max(d, axis=axes), d, axes
    return np.max(d), d, None                                  # LF:93-94


def maxbin_penalty(layers):
    total, normalizer = F32(0.0), 0.0
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        k_mb, _, _ = _maxbin_of(K, sK)
        b_mb, _, _ = _maxbin_of(b, sb)
        k_dim, b_dim = _dim(K.shape), _dim(b.shape)
        layer_penalty = np.mean(k_mb, dtype=F32) * F32(k_dim) + np.mean(b_mb, dtype=F32) * F32(b_dim)  # LF:110
        total = F32(total + layer_penalty)                     # LF:112
        normalizer += k_dim + b_dim                            # LF:114
    return F32(total / F32(normalizer))                        # LF:116


def difference_penalty(layers):
    total, normalizer = F32(0.0), 0.0
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        k_pen = np.mean(np.abs(K - K / sK), dtype=F32)         # LF:175,178 (un-floored K/s)
        b_pen = np.mean(np.abs(b - b / sb), dtype=F32)         # LF:176,179
        k_dim, b_dim = _dim(K.shape), _dim(b.shape)
        total = F32(total + (k_pen * F32(k_dim) + b_pen * F32(b_dim)))   # LF:189-191
        normalizer += k_dim + b_dim
    return F32(total / F32(normalizer))                        # LF:195


def inverse_penalty(layers):
    total, normalizer = F32(0.0), 0.0
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        sK_nz = np.where(sK == F32(0.0), F32(EPS_FLOAT32), sK)  # LF:255
        sb_nz = np.where(sb == F32(0.0), F32(EPS_FLOAT32), sb)  # LF:256
        k_inv = np.mean(F32(1.0) / sK_nz, dtype=F32)            # LF:258
        b_inv = np.mean(F32(1.0) / sb_nz, dtype=F32)            # LF:259
        k_dim, b_dim = _dim(K.shape), _dim(b.shape)
        total = F32(total + (k_inv * F32(k_dim) + b_inv * F32(b_dim)))   # LF:269-271
        normalizer += k_dim + b_dim
    return F32(total / F32(normalizer))                        # LF:275


def _normalizer(layers):
    n = 0.0
    for K, _, b, _ in layers:
        n += _dim(np.shape(K)) + _dim(np.shape(b))
    return n


def _sum_to_shape(g, shape, f64=False):
    """_RealDivGrad's reduce_sum over broadcast axes + reshape to the divisor's shape."""
    g = np.asarray(g)
    acc = np.float64 if f64 else F32
    if len(shape) == 1 and g.ndim != 1:          # scale (1,) against a rank-n tensor
        return np.reshape(np.sum(g, dtype=acc), shape).astype(F32)
    axes = tuple(i for i in range(len(shape)) if shape[i] == 1 and g.shape[i] != 1)
    return np.reshape(np.sum(g, axis=axes, dtype=acc), shape).astype(F32)


def _maxbin_grad_one(P, s, dim, upstream_total, accum):
    """Gradients of  reduce_mean(reduce_max(abs(P)/s, axes)) * dim  w.r.t. P and s."""
    mb, d, axes = _maxbin_of(P, s)
    n_groups = mb.size
    g_mean = F32(upstream_total * F32(dim))                    # d(mean*dim)/d mean
    g_mb = F32(g_mean / F32(n_groups))                         # _MeanGrad
    mb_b = np.reshape(mb, s.shape) if axes is not None else mb
    indicators = (d == mb_b).astype(F32)                       # _MinOrMaxGrad
    if axes is not None:
        num_selected = np.reshape(np.sum(indicators, axis=axes, dtype=F32), s.shape)
    else:
        num_selected = np.sum(indicators, dtype=F32)
    g_d = indicators / num_selected * g_mb
    a = np.abs(P)
    g_a = g_d / s                                              # _RealDivGrad, numerator
    g_s = _sum_to_shape(g_d * ((-a / s) / s), s.shape, accum == "fp64")   # denominator
    g_P = g_a * np.sign(P)                                     # _AbsGrad
    return g_P.astype(F32), g_s


def maxbin_penalty_grads(layers, upstream=1.0, accum="fp32"):
    """d(upstream * maxbin_penalty)/d(K, sK, b, sb) per layer, TF autodiff rules."""
    g_total = F32(F32(upstream) / F32(_normalizer(layers)))    # LF:116
    out = []
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        gK, gsK = _maxbin_grad_one(K, sK, _dim(K.shape), g_total, accum)
        gb, gsb = _maxbin_grad_one(b, sb, _dim(b.shape), g_total, accum)
        out.append((gK, gsK, gb, gsb))
    return out


def _difference_grad_one(P, s, dim, upstream_total, accum):
    n = P.size
    Pq = P / s                                                 # LF:175
    d = P - Pq                                                 # LF:178
    g_mean = F32(upstream_total * F32(dim))
    g_abs = F32(g_mean / F32(n))                               # _MeanGrad
    g_d = g_abs * np.sign(d)                                   # _AbsGrad
    g_P = g_d + (-g_d) / s                                     # _SubGrad + _RealDivGrad numerator
    g_s = _sum_to_shape((-g_d) * ((-P / s) / s), s.shape, accum == "fp64")
    return g_P.astype(F32), g_s


def difference_penalty_grads(layers, upstream=1.0, accum="fp32"):
    g_total = F32(F32(upstream) / F32(_normalizer(layers)))
    out = []
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        gK, gsK = _difference_grad_one(K, sK, _dim(K.shape), g_total, accum)
        gb, gsb = _difference_grad_one(b, sb, _dim(b.shape), g_total, accum)
        out.append((gK, gsK, gb, gsb))
    return out


def _inverse_grad_one(s, dim, upstream_total):
    nz = np.where(s == F32(0.0), F32(EPS_FLOAT32), s)
    g_mean = F32(upstream_total * F32(dim))
    g_inv = F32(g_mean / F32(s.size))
    g_nz = g_inv * ((F32(-1.0) / nz) / nz)                     # _RealDivGrad, x = 1.0
    return np.where(s == F32(0.0), F32(0.0), g_nz).astype(F32)  # _SelectGrad


def inverse_penalty_grads(layers, upstream=1.0, accum="fp32"):
    g_total = F32(F32(upstream) / F32(_normalizer(layers)))
    out = []
    for K, sK, b, sb in layers:
        K, sK, b, sb = (np.asarray(a, F32) for a in (K, sK, b, sb))
        out.append((np.zeros_like(K), _inverse_grad_one(sK, _dim(K.shape), g_total),
                    np.zeros_like(b), _inverse_grad_one(sb, _dim(b.shape), g_total)))
    return out


# ----------------------------------------------------------------------------------
# f2 helpers: what the callbacks / export recompute on the host
# (custom_callbacks.py:84-129, utils/log_scripts.py:61-97)
# ----------------------------------------------------------------------------------
def codes_int8(P, s):
    """floor(P/s).astype(int8) exactly as log_scripts.py:74-79 does (wraps silently)."""
    q, _ = fq_forward(P, s)
    return q.astype(np.int64).astype(np.int8)


def code_stats(P, s):
    """(number of unique codes, max|q|) as custom_callbacks.py:84-129 logs them."""
    q, _ = fq_forward(P, s)
    return int(np.unique(q).size), float(np.max(np.abs(q)))


# ----------------------------------------------------------------------------------
# f1  optimizer step + constraint (the step right after the hot path)
#
# PARITY STATUS of this block: UNPINNED.  The update rules live in Keras 2.11
# (tensorflow==2.11.0, requirements.txt:1; keras/optimizers/optimizer_experimental/adam.py
# and sgd.py, `update_step`), which is neither vendored in the reference nor installable
# here; they are restated from the published source, one numpy fp32 op per TF op.  Call
# sites: IMAGENETTE/nested_quantization_layer/experiment.py:763-771 (Adam(1e-4)),
# CIFAR-10/paper_implementation/experiment.py:130-141 (SGD 0.1 / 0.9 / nesterov / wd 5e-4);
# the constraint is NQ:35-46 applied after the update (Optimizer.apply_gradients).
# ----------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, step, beta_1=0.9, beta_2=0.999, epsilon=1e-7, l2=0.0, min_value=None,
              weight_decay=0.0, grad_scale=1.0):
    """One Keras-2.11 Adam update; returns (p, m, v).  `l2` adds the keras.regularizers.l2
    gradient 2*l2*p; `min_value` applies MinValueConstraint afterwards."""
    p, g, m, v = (np.asarray(a, F32) for a in (p, g, m, v))
    lr, b1, b2, eps = F32(lr), F32(beta_1), F32(beta_2), F32(epsilon)
    g = g * F32(grad_scale)
    if l2:
        g = g + (F32(2.0) * F32(l2)) * p
    if weight_decay:
        p = p - p * F32(weight_decay) * lr
    local_step = F32(step)
    beta_1_power = np.power(b1, local_step, dtype=F32)
    beta_2_power = np.power(b2, local_step, dtype=F32)
    alpha = lr * np.sqrt(F32(1) - beta_2_power) / (F32(1) - beta_1_power)
    m = m + (g - m) * (F32(1) - b1)
    v = v + (np.square(g) - v) * (F32(1) - b2)
    p = p - (m * alpha) / (np.sqrt(v) + eps)
    if min_value is not None:
        p = min_value_constraint(p, F32(min_value))
    return p.astype(F32), m.astype(F32), v.astype(F32)


def sgd_step(p, g, m, lr, momentum=0.0, nesterov=False, l2=0.0, min_value=None, weight_decay=0.0, grad_scale=1.0):
    """One Keras-2.11 SGD update; returns (p, m)."""
    p, g = np.asarray(p, F32), np.asarray(g, F32)
    lr, mom = F32(lr), F32(momentum)
    g = g * F32(grad_scale)
    if l2:
        g = g + (F32(2.0) * F32(l2)) * p
    if weight_decay:
        p = p - p * F32(weight_decay) * lr
    if momentum:
        m = -g * lr + np.asarray(m, F32) * mom
        p = p + ((-g * lr + m * mom) if nesterov else m)
    else:
        p = p + (-g * lr)
    if min_value is not None:
        p = min_value_constraint(p, F32(min_value))
    return p.astype(F32), (None if m is None else np.asarray(m, F32))
