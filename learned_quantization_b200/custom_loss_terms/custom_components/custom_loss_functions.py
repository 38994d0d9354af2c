"""Drop-in for */custom_loss_terms/custom_components/custom_loss_functions.py (LF).

SCCEMaxBin / SCCEDifference / SCCEInverse keep the reference's constructor and methods.
The penalty reductions over all layers run as multi-tensor launches in liblq.so
(lq_penalty_fwd) and their gradients -- what TensorFlow's autodiff derives in the
reference -- come from lq_penalty_bwd through a torch.autograd.Function.

Layers may expose either `.kernel / .nested_q_k_layer` (CIFAR-10 / IMAGENETTE files) or
`.W / .nested_q_w_layer` (MNIST file, LF_MNIST:85-86).
"""
from __future__ import annotations

import os

import torch

from ... import _lib, ops

eps_float32 = ops.EPS_FLOAT32


class _Penalty(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, *tensors):
        Ps = [ops.dense(t.detach()) for t in tensors[0::2]]
        ss = [t.detach().contiguous() for t in tensors[1::2]]
        layouts = [ops.tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
        out, ws = ops.penalty_fwd(kind, Ps, ss, layouts)
        ctx.kind, ctx.layouts, ctx.ws = kind, layouts, ws
        ctx.Ps, ctx.ss = Ps, ss
        ctx.s_shapes = [t.shape for t in tensors[1::2]]
        return out

    @staticmethod
    def backward(ctx, g):
        dXs, dss = ops.penalty_bwd(ctx.kind, ctx.Ps, ctx.ss, ctx.ws, 1.0,
                                   upstream_dev=g.detach().contiguous().to(torch.float32), layouts=ctx.layouts)
        grads = [None]
        for dx, ds, shp in zip(dXs, dss, ctx.s_shapes):
            grads += [dx, ds.view(shp)]
        return tuple(grads)


class _PenaltyFromForward(torch.autograd.Function):
    """The same term when the step's multi-tensor forward has already produced its value and workspace as by-products
    (lq_mt_fq_fwd_penalty through training.step.QuantBank): no second pass over the tensors; the backward is _Penalty's."""

    @staticmethod
    def forward(ctx, kind, fused, *tensors):
        ctx.kind, ctx.fused = kind, fused
        ctx.s_shapes = [t.shape for t in tensors[1::2]]
        return fused["out"].clone()

    @staticmethod
    def backward(ctx, g):
        f = ctx.fused
        dXs, dss = ops.penalty_bwd(ctx.kind, f["Ps"], f["ss"], f["ws"], 1.0,
                                   upstream_dev=g.detach().contiguous().to(torch.float32), layouts=f["layouts"])
        grads = [None, None]
        for dx, ds, shp in zip(dXs, dss, ctx.s_shapes):
            grads += [dx, ds.view(shp)]
        return tuple(grads)


def _layer_tensors(layers):
    tensors = []
    for layer in layers:
        if hasattr(layer, "kernel"):
            tensors += [layer.kernel, layer.nested_q_k_layer.scale]      # LF:85-86
        else:
            tensors += [layer.W, layer.nested_q_w_layer.scale]           # LF_MNIST:85-86
        tensors += [layer.b, layer.nested_q_b_layer.scale]               # LF:87-88
    return tensors


def sparse_categorical_crossentropy(y_true, y_pred):
    """Keras' from_logits=False path: per-sample -log(clip(p[y], 1e-7, 1 - 1e-7))."""
    idx = y_true.reshape(-1, 1).to(torch.int64)
    p = torch.gather(y_pred, -1, idx).squeeze(-1)
    return -torch.log(torch.clamp(p, 1e-7, 1.0 - 1e-7))


class _SCCEBase:
    _kind = None
    _log_name = ""

    def __init__(self, layers, penalty_rate, log_dir, l2_lambda=0.01):
        self.layers = layers
        self.penalty_rate = penalty_rate
        self.custom_loss_dir = os.path.join(log_dir, "custom_losses")
        os.makedirs(self.custom_loss_dir, exist_ok=True)               # LF:13-30 setup_logger
        self._logs = ["total_loss.log", "scce_loss.log", self._log_name]
        for log in self._logs:
            with open(os.path.join(self.custom_loss_dir, log), "w"):
                pass
        self.log_to_files = True
        self._pending = []

    _fused = None      # set by QuantBank.forward for the duration of one training step (value + workspace of this term)

    def _penalty(self):
        fused, self._fused = self._fused, None
        if fused is not None:
            return _PenaltyFromForward.apply(self._kind, fused, *_layer_tensors(self.layers))
        return _Penalty.apply(self._kind, *_layer_tensors(self.layers))

    def compute_total_loss(self, y_true, y_pred):
        """scce + penalty_rate * penalty, per sample (LF:47-73); the three tf.print log lines
        of the reference are buffered on the device and written by flush_logs()."""
        cross_entropy_loss = sparse_categorical_crossentropy(y_true, y_pred)
        penalty = self._penalty()
        total_loss = cross_entropy_loss + self.penalty_rate * penalty
        if self.log_to_files:
            self._pending.append(torch.stack([total_loss.detach().mean(), cross_entropy_loss.detach().mean(),
                                              (self.penalty_rate * penalty).detach()]))
            if len(self._pending) >= 256:
                self.flush_logs()
        return total_loss

    def flush_logs(self):
        if not self._pending:
            return
        rows = torch.stack(self._pending).cpu().tolist()
        self._pending = []
        for col, log in enumerate(self._logs):
            with open(os.path.join(self.custom_loss_dir, log), "a") as f:
                for r in rows:
                    f.write(f"{r[col]:g}\n")


class SCCEMaxBin(_SCCEBase):
    """LF:33-116."""
    _kind = _lib.LQ_PEN_MAXBIN
    _log_name = "maxbin_loss.log"

    def compute_maxbin_penalty(self):
        return self._penalty()


class SCCEDifference(_SCCEBase):
    """LF:119-195."""
    _kind = _lib.LQ_PEN_DIFFERENCE
    _log_name = "difference_loss.log"

    def compute_difference_penalty(self):
        return self._penalty()


class SCCEInverse(_SCCEBase):
    """LF:198-275."""
    _kind = _lib.LQ_PEN_INVERSE
    _log_name = "inverse_loss.log"

    def compute_inverse_penalty(self):
        return self._penalty()
