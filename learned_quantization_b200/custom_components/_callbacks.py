"""Shared implementation of the reference's per-layer tracking callbacks
(*/custom_components/custom_callbacks.py: NestedScaleTrackingCallback :9-208,
AccuracyLossTrackingCallBack :211-237) with the same log files and line formats, so that the
reference's utils/plot_scripts.py can parse the output.

What changes is where the numbers come from: the reference pulls every weight tensor to the host
each epoch and recomputes floor(P / s), np.unique and max|q| there (:84-129); here the integer-code
statistics are computed on the device (ops.code_stats -> lq_code_range / lq_code_histogram) and only
the non-empty histogram bins travel.  The raw per-epoch weight dumps of the reference (one text line
per weight, :51-82) are kept but can be switched off (`dump_values=False`): they are by far the
largest cost of the reference's epoch end and nothing in its plot scripts reads them.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops


def _shape_str(t) -> str:
    return str(tuple(int(d) for d in t.shape))            # TensorShape prints like a tuple in the file names


def _scale_name(orientation: str) -> str:
    # NQ:154,166,178,187 (the channelwise branch re-uses "Columnwise-scaler")
    return {"rowwise": "Rowwise-scaler", "columnwise": "Columnwise-scaler", "channelwise": "Columnwise-scaler",
            "scalar": "Scalar-scaler"}[orientation]


class NestedScaleTrackingCallback:
    """Keras-callback-shaped object: call on_train_begin / on_epoch_end(epoch) / on_train_end."""

    def __init__(self, layer, log_dir, dump_values: bool = True):
        self.layer, self.dump_values = layer, dump_values
        dense = hasattr(layer, "W")
        self._param = (lambda: layer.W) if dense else (lambda: layer.kernel)
        self.qk_layer = layer.nested_q_w_layer if dense else layer.nested_q_k_layer
        self.qb_layer = layer.nested_q_b_layer
        self.k_scale, self.b_scale = self.qk_layer.scale, self.qb_layer.scale
        kernel_name = "Weights" if dense else f"Kernel_{layer.name}"      # NQ:248 / NQ:324
        bias_name = "Bias" if dense else f"Bias_{layer.name}"             # NQ:256 / NQ:331
        ks, bs = _scale_name(self.qk_layer.orientation), _scale_name(self.qb_layer.orientation)
        kshape, bshape = _shape_str(self._param()), _shape_str(layer.b)
        kss, bss = _shape_str(self.k_scale), _shape_str(self.b_scale)
        e, t, b0 = f"{log_dir}/on_epoch_end", f"{log_dir}/on_train_end", f"{log_dir}/on_train_begin"
        for d in (e, t, b0):
            os.makedirs(d, exist_ok=True)
        self.log_file_path_kernel = f"{e}/{kernel_name}_{kshape}.log"
        self.log_file_path_biases = f"{e}/{bias_name}_{bshape}.log"
        self.log_file_path_k_scale = f"{e}/{kernel_name}_{kshape}_{ks}_{kss}.log"
        self.log_file_path_b_scale = f"{e}/{bias_name}_{bshape}_{bs}_{bss}.log"
        self.log_file_path_qk_epoch = f"{e}/Number_of_unique_{kernel_name}_{kshape}.log"
        self.log_file_path_qb_epoch = f"{e}/Number_of_unique_{bias_name}_{bshape}.log"
        self.log_file_path_qk_epoch_max = f"{e}/Max_{kernel_name}_{kshape}.log"
        self.log_file_path_qb_epoch_max = f"{e}/Max_{bias_name}_{bshape}.log"
        self.log_file_path_qk = f"{t}/Quantized_{kernel_name}_{kshape}_{ks}_{kss}.log"
        self.log_file_path_qb = f"{t}/Quantized_{bias_name}_{bshape}_{bs}_{bss}.log"
        self.log_file_path_qk_unique = f"{t}/Unique_quantized_{kernel_name}_{kshape}_{ks}_{kss}.log"
        self.log_file_path_qb_unique = f"{t}/Unique_quantized_{bias_name}_{bshape}_{bs}_{bss}.log"
        # the IMAGENETTE / CIFAR-10 files join this one name with double underscores (custom_callbacks.py:43), the MNIST file
        # (dense layers, :43 there) with single ones
        sep = "_" if dense else "__"
        self.log_file_path_qk_initial = f"{b0}/Initial_Quantized_{kernel_name}_{kshape}{sep}{ks}{sep}{kss}.log"
        self.log_file_path_qb_initial = f"{b0}/Initial_Quantized_{bias_name}_{bshape}_{bs}_{bss}.log"
        self.log_file_path_qk_initial_unique = f"{b0}/Unique_initial_quantized_{kernel_name}_{kshape}_{ks}_{kss}.log"
        self.log_file_path_qb_initial_unique = f"{b0}/Unique_initial_quantized_{bias_name}_{bshape}_{bs}_{bss}.log"

    # -- helpers
    @staticmethod
    def _append_values(path, header, values):
        with open(path, "a") as f:
            if header is not None:
                f.write(header)
            f.write("".join(f"{v}\n" for v in values))

    def _codes(self, p, scale):
        P, s = ops.dense(p.detach()), scale.detach().contiguous()
        return ops.fq_fwd(P, s, codes_dtype=torch.float32)[1], ops.code_stats(P, s)

    def _dump_codes(self, p, scale, path_values, path_unique):
        codes, st = self._codes(p, scale)
        self._append_values(path_values, None, codes.flatten().cpu().numpy())
        with open(path_unique, "a") as f:
            for v, c in zip(st["unique"].numpy(), st["counts"].numpy()):
                f.write(f"{v}, {c}\n")

    # -- the three hooks
    def on_train_begin(self, logs=None):
        self._dump_codes(self._param(), self.k_scale, self.log_file_path_qk_initial, self.log_file_path_qk_initial_unique)
        self._dump_codes(self.layer.b, self.b_scale, self.log_file_path_qb_initial, self.log_file_path_qb_initial_unique)

    def on_epoch_end(self, epoch, logs=None):
        hdr = f"Epoch {epoch}\n"
        if self.dump_values:                                                # :51-82
            for path, t in ((self.log_file_path_kernel, self._param()), (self.log_file_path_biases, self.layer.b),
                            (self.log_file_path_k_scale, self.k_scale), (self.log_file_path_b_scale, self.b_scale)):
                self._append_values(path, hdr, t.detach().flatten().cpu().numpy())
        codes, st = self._codes(self._param(), self.k_scale)               # :84-95
        self._append_values(self.log_file_path_qk_epoch, hdr, [st["n_unique"]])
        mx = codes.abs().amax(dim=1) if codes.dim() > 1 else codes.abs().amax().reshape(1)   # :98 reduce_max(axis=1)
        self._append_values(self.log_file_path_qk_epoch_max, hdr, mx.flatten().cpu().numpy())
        _, stb = self._codes(self.layer.b, self.b_scale)                    # :106-129
        self._append_values(self.log_file_path_qb_epoch, hdr, [stb["n_unique"]])
        self._append_values(self.log_file_path_qb_epoch_max, hdr, [np.float32(stb["max_abs"])])

    def on_train_end(self, logs=None):
        self._dump_codes(self._param(), self.k_scale, self.log_file_path_qk, self.log_file_path_qk_unique)
        self._dump_codes(self.layer.b, self.b_scale, self.log_file_path_qb, self.log_file_path_qb_unique)


class AccuracyLossTrackingCallBack:
    """:211-237 -- accuracy/ and loss/ logs, 'Epoch N' then the value."""

    def __init__(self, log_dir, accuracy_file="accuracy.log", loss_file="loss.log"):
        os.makedirs(f"{log_dir}/accuracy", exist_ok=True)
        self.accuracy_log_file_path = f"{log_dir}/accuracy/train_{accuracy_file}"
        self.val_accuracy_log_file_path = f"{log_dir}/accuracy/val_{accuracy_file}"
        os.makedirs(f"{log_dir}/loss", exist_ok=True)
        self.loss_log_file_path = f"{log_dir}/loss/train_{loss_file}"
        self.val_loss_log_file_path = f"{log_dir}/loss/val_{loss_file}"

    def on_epoch_end(self, epoch, logs=None):
        for path, key in ((self.val_accuracy_log_file_path, "val_accuracy"), (self.val_loss_log_file_path, "val_loss"),
                          (self.accuracy_log_file_path, "accuracy"), (self.loss_log_file_path, "loss")):
            with open(path, "a") as f:
                f.write(f"Epoch {epoch}\n")
                f.write(f"{logs[key]}\n")
