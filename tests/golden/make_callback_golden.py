#!/usr/bin/env python
"""Generate tests/golden/callback_logs.json by running the REFERENCE'S OWN custom_callbacks.py (imported
unmodified from /root/reference) on the torch-CPU TensorFlow shim, for one conv layer (IMAGENETTE file) and one
dense layer (MNIST file): on_train_begin, two epochs with a parameter change in between, on_train_end, plus
AccuracyLossTrackingCallBack.  The fixture holds every file the reference wrote (relative path -> text) and the
parameters at each stage; tests/test_gpu_training.py makes the drop-in callbacks write the same tree from the
same parameters and compares byte for byte, tests/test_callback_logs.py feeds the tree to the reference's
utils/plot_scripts.py parsers.  Dev container only:

    python tests/golden/make_callback_golden.py
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_shim  # noqa: E402

REF = os.environ.get("LQ_REFERENCE_ROOT", "/root/reference")
CALLBACKS = {"conv": "IMAGENETTE/nested_quantization_layer/custom_components/custom_callbacks.py",
             "dense": "MNIST/nested_quantization_layer/custom_components/custom_callbacks.py"}
LAYERS = {"conv": "IMAGENETTE/nested_quantization_layer/custom_components/custom_layers.py",
          "dense": "MNIST/nested_quantization_layer/custom_components/custom_layers.py"}


class Var:
    """What the callbacks touch of a tf.Variable: .name, .shape (prints like a TensorShape), .numpy(), division."""

    def __init__(self, value: np.ndarray, name: str):
        self.value = torch.as_tensor(np.asarray(value, np.float32))
        self.name = name
        self.shape = tuple(self.value.shape)

    def numpy(self):
        return self.value.numpy()

    def __truediv__(self, other):
        return self.value / (other.value if isinstance(other, Var) else other)


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def install_callbacks_shim():
    tf = tf_shim.install()
    cb = types.ModuleType("tensorflow.keras.callbacks")

    class Callback:
        def __init__(self):
            pass
    cb.Callback = Callback
    tf.keras.callbacks = cb
    sys.modules["tensorflow.keras.callbacks"] = cb
    return tf


def stages(rng, kind):
    """Parameters at train begin, after epoch 0 and after epoch 1 (= train end)."""
    if kind == "conv":
        kshape, sshape, n = (3, 3, 4, 8), (1, 1, 4, 1), 8        # channelwise
    else:
        kshape, sshape, n = (20, 6), (20, 1), 6                  # rowwise
    out = []
    K = rng.normal(0, 0.05, kshape).astype(np.float32)
    b = rng.normal(0, 0.05, (n,)).astype(np.float32)
    sK = rng.uniform(5e-3, 2e-2, sshape).astype(np.float32)
    sb = rng.uniform(5e-3, 2e-2, (1,)).astype(np.float32)
    for _ in range(3):
        out.append(tuple(a.copy() for a in (K, sK, b, sb)))
        K = (K + rng.normal(0, 0.01, kshape)).astype(np.float32)
        b = (b + rng.normal(0, 0.01, (n,))).astype(np.float32)
        sK = (sK * rng.uniform(0.9, 1.2, sshape)).astype(np.float32)
        sb = (sb * rng.uniform(0.9, 1.2, (1,))).astype(np.float32)
    return out


def main():
    install_callbacks_shim()
    rng = np.random.default_rng(2024)
    fixture = {"files": {}, "params": {}, "layer_names": {}}
    arrays = {}
    for kind in ("conv", "dense"):
        layers = load(LAYERS[kind], f"ref_layers_{kind}")
        cbm = load(CALLBACKS[kind], f"ref_callbacks_{kind}")
        # a reference layer, for the names Keras would give its variables (add_weight(name=...), NQ:154-187,248-331)
        if kind == "conv":
            layer = layers.CustomConv2DLayer(seed=1, penalty_threshold=1e-3, orientation="channelwise",
                                             initializer=tf_shim.RandomNormal(seed=1), filters=8, kernel_size=(3, 3),
                                             strides=(1, 1), padding="same", name="x", regularizer=None)
            layer(torch.zeros(1, 5, 5, 4))
            names = (f"{layer.name}/Kernel_{layer.name}:0", f"{layer.name}/Bias_{layer.name}:0",
                     f"{layer.name}/custom_quantized_scale_layer/Columnwise-scaler:0",
                     f"{layer.name}/custom_quantized_scale_layer_1/Scalar-scaler:0")
        else:
            layer = layers.CustomDenseLayer(seed=1, units=6, penalty_threshold=1e-3, orientation="rowwise",
                                            initializer=tf_shim.RandomNormal(seed=1), name="x", regularizer=None)
            layer(torch.zeros(1, 20))
            names = (f"{layer.name}/Weights:0", f"{layer.name}/Bias:0",
                     f"{layer.name}/custom_quantized_scale_layer_2/Rowwise-scaler:0",
                     f"{layer.name}/custom_quantized_scale_layer_3/Scalar-scaler:0")
        fixture["layer_names"][kind] = layer.name
        st = stages(rng, kind)
        for i, (K, sK, b, sb) in enumerate(st):
            for nm, a in (("K", K), ("sK", sK), ("b", b), ("sb", sb)):
                arrays[f"{kind}/stage{i}/{nm}"] = a

        class Host:
            pass
        host, qk, qb = Host(), Host(), Host()

        def put(stage):
            K, sK, b, sb = st[stage]
            setattr(host, "kernel" if kind == "conv" else "W", Var(K, names[0]))
            host.b = Var(b, names[1])
            qk.scale, qb.scale = Var(sK, names[2]), Var(sb, names[3])
        put(0)
        setattr(host, "nested_q_k_layer" if kind == "conv" else "nested_q_w_layer", qk)
        host.nested_q_b_layer = qb
        with tempfile.TemporaryDirectory() as tmp:
            cwd = os.getcwd()
            os.chdir(tmp)
            try:
                os.makedirs("logs/temp", exist_ok=True)           # the reference appends to ./logs/temp/*.log
                cb = cbm.NestedScaleTrackingCallback(host, "run")
                # the reference caches the scale variables in __init__: stage changes go through the same objects
                def restage(stage):
                    K, sK, b, sb = st[stage]
                    (host.kernel if kind == "conv" else host.W).value = torch.as_tensor(K)
                    host.b.value = torch.as_tensor(b)
                    qk.scale.value, qb.scale.value = torch.as_tensor(sK), torch.as_tensor(sb)
                cb.on_train_begin()
                for epoch in (0, 1):
                    restage(epoch + 1)
                    if hasattr(cb, "on_epoch_begin"):
                        cb.on_epoch_begin(epoch)
                    cb.on_epoch_end(epoch)
                cb.on_train_end()
                acc = cbm.AccuracyLossTrackingCallBack("run")
                for epoch, logs in enumerate(({"val_accuracy": 0.5, "val_loss": 1.25, "accuracy": 0.4375, "loss": 1.5},
                                              {"val_accuracy": 0.625, "val_loss": 1.0, "accuracy": 0.5625, "loss": 1.125})):
                    acc.on_epoch_end(epoch, logs)
                files = {}
                for root, _, fs in os.walk("run"):
                    for f in fs:
                        p = os.path.join(root, f)
                        with open(p) as fh:
                            files[os.path.relpath(p, "run")] = fh.read()
                fixture["files"][kind] = files
            finally:
                os.chdir(cwd)
    with open(os.path.join(HERE, "callback_logs.json"), "w") as f:
        json.dump(fixture, f, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "callback_logs.npz"), **arrays)
    for kind, files in fixture["files"].items():
        print(kind, len(files), "files;", sum(len(v) for v in files.values()), "bytes")
        for k in sorted(files):
            print("   ", k)


if __name__ == "__main__":
    main()
