#!/usr/bin/env python
"""Generate tests/golden/reference_vectors.{npz,json} by running the REFERENCE'S OWN
SOURCE FILES (imported unmodified from /root/reference) on the torch-CPU TensorFlow
shim in tf_shim.py.  Run in the dev container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The outputs are committed; tests read only the committed files.
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
FILES = {
    "NQ": "IMAGENETTE/nested_quantization_layer/custom_components/custom_layers.py",
    "NQ_CIFAR": "CIFAR-10/nested_quantization_layer/custom_components/custom_layers.py",
    "NQ_MNIST": "MNIST/nested_quantization_layer/custom_components/custom_layers.py",
    "CL": "IMAGENETTE/custom_loss_terms/custom_components/custom_layers.py",
    "CL_MNIST": "MNIST/custom_loss_terms/custom_components/custom_layers.py",
    "PI": "CIFAR-10/paper_implementation/custom_components/custom_layers.py",
    "LF": "IMAGENETTE/custom_loss_terms/custom_components/custom_loss_functions.py",
    "LF_MNIST": "MNIST/custom_loss_terms/custom_components/custom_loss_functions.py",
}

SCALE_INIT = np.float32(np.finfo(np.float32).eps * 100)
ORIENTS = ("rowwise", "columnwise", "channelwise", "scalar")


def load_reference():
    tf_shim.install()
    mods = {}
    for key, rel in FILES.items():
        spec = importlib.util.spec_from_file_location(f"ref_{key}", os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[key] = m
    return mods


def np32(t):
    return t.detach().cpu().numpy().astype(np.float32)


def make_inputs(rng, shape, s_shape, s_kind):
    P = rng.normal(0.0, 0.05, size=shape).astype(np.float32)
    flat = P.reshape(-1)
    n = flat.size
    flat[rng.integers(0, n, size=max(1, n // 40))] = 0.0          # exact zeros
    flat[rng.integers(0, n, size=max(1, n // 80))] = -0.0
    dy = rng.normal(0.0, 1e-3, size=shape).astype(np.float32)
    dy.reshape(-1)[rng.integers(0, n, size=max(1, n // 40))] = 0.0
    if s_kind == "init":
        s = np.full(s_shape, SCALE_INIT, dtype=np.float32)
    else:
        s = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
        # put a few parameters exactly on the grid q*s (division must return q exactly)
        sb = np.broadcast_to(s if len(s_shape) == len(shape) else s.reshape((1,) * len(shape)), shape)
        idx = rng.integers(0, n, size=max(1, n // 20))
        qs = rng.integers(-40, 40, size=idx.size).astype(np.float32)
        flat[idx] = qs * sb.reshape(-1)[idx]
    return P, s, dy


def scale_shape_of(mod, shape, orient, **ctor):
    layer = mod.CustomQuantizedScaleLayer(initializer=None, orientation=orient, **ctor)
    layer.build(tuple(shape))
    return tuple(layer.scale.shape), layer


def run_op(mod, P, s, dy, lam=None):
    tf_shim.TRACE["floor"].clear()
    Pt = torch.tensor(P, requires_grad=True)
    st = torch.tensor(s, requires_grad=True)
    y = mod.my_custom_gradient(Pt, st, lam) if lam is not None else mod.my_custom_gradient(Pt, st)
    q = np32(tf_shim.TRACE["floor"][-1])
    y.backward(torch.tensor(dy))
    dP = np32(Pt.grad)
    assert np.array_equal(dP.view(np.uint32), dy.view(np.uint32)), "reference dP must be dy itself"
    return q, np32(y), np32(st.grad)


def main():
    out_npz = os.path.join(HERE, "reference_vectors.npz")
    out_json = os.path.join(HERE, "reference_vectors.json")
    arrays, manifest = {}, {"ops": [], "ste": [], "loss": [], "layers": [], "shapes": [], "constraint": []}
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix="lq_golden_")
    os.chdir(tmp)   # the reference's setup_logger() creates logs/temp under cwd
    try:
        mods = load_reference()
        rng = np.random.default_rng(20261018)

        # ---------------- scale shapes per orientation (NQ:147-197) ----------------
        for shape in [(3, 3, 8, 16), (7, 7, 3, 4), (1, 1, 4, 8), (24, 10), (16,)]:
            for orient in ORIENTS:
                if len(shape) == 1 and orient in ("columnwise", "channelwise"):
                    continue  # input_shape[i] with i beyond rank never evaluated -> all ones
                shp, _ = scale_shape_of(mods["NQ"], shape, orient, penalty_threshold=0.0)
                manifest["shapes"].append({"param": list(shape), "orientation": orient, "scale": list(shp)})
        try:
            scale_shape_of(mods["NQ"], (3, 3), "diagonal", penalty_threshold=0.0)
            manifest["bad_orientation_raises"] = None
        except ValueError as e:
            manifest["bad_orientation_raises"] = str(e)

        # ---------------- operator fwd/bwd, nested variant (NQ:49-120) ----------------
        lams = [0.0, 1e-11, 1e-10, 1e-3, 0.05, 1000.0]
        op_shapes = [((3, 3, 8, 16), ORIENTS), ((1, 1, 4, 8), ("channelwise", "scalar")),
                     ((24, 10), ORIENTS), ((16,), ("scalar",)), ((5, 3, 6, 7), ("channelwise", "rowwise", "columnwise"))]
        in_id = 0
        for shape, orients in op_shapes:
            for orient in orients:
                s_shape, _ = scale_shape_of(mods["NQ"], shape, orient, penalty_threshold=0.0)
                for s_kind in ("init", "trained"):
                    P, s, dy = make_inputs(rng, shape, s_shape, s_kind)
                    if s_kind == "trained" and orient == "rowwise" and len(shape) > 1:
                        dy[0] *= np.float32(1e6)   # group 0 becomes "all above threshold" for mid lambdas
                    key = f"in{in_id:03d}"
                    in_id += 1
                    arrays[f"{key}/P"], arrays[f"{key}/s"], arrays[f"{key}/dy"] = P, s, dy
                    for lam in lams:
                        variants = ["NQ"] + (["NQ_CIFAR", "NQ_MNIST"] if (lam in (1e-3, 0.05) and s_kind == "trained") else [])
                        for mk in variants:
                            try:
                                q, y, ds = run_op(mods[mk], P, s, dy, lam)
                            except Exception as e:          # MNIST's expand_dims(axis=list) variant
                                manifest.setdefault("skipped", []).append(f"{mk} {shape} {orient}: {type(e).__name__}")
                                continue
                            cid = f"op{len(manifest['ops']):04d}"
                            arrays[f"{cid}/q"], arrays[f"{cid}/y"], arrays[f"{cid}/ds"] = q, y, ds
                            manifest["ops"].append({"id": cid, "inputs": key, "module": mk, "shape": list(shape),
                                                    "orientation": orient, "s_kind": s_kind, "lam": lam})

        # SURVEY Appendix B known-answer vector, executed by the reference source
        P = np.array([[0.30, -0.26, 0.04], [-0.05, 0.0, 0.11]], np.float32)
        dy = np.array([[0.003, -0.5, 0.0], [0.2, 0.1, -0.3]], np.float32)
        for s in (np.array([[0.1], [0.02]], np.float32), np.array([0.1], np.float32)):
            key = f"in{in_id:03d}"
            in_id += 1
            arrays[f"{key}/P"], arrays[f"{key}/s"], arrays[f"{key}/dy"] = P, s, dy
            for lam in (0.0, 0.05, 1000.0):
                q, y, ds = run_op(mods["NQ"], P, s, dy, lam)
                cid = f"op{len(manifest['ops']):04d}"
                arrays[f"{cid}/q"], arrays[f"{cid}/y"], arrays[f"{cid}/ds"] = q, y, ds
                manifest["ops"].append({"id": cid, "inputs": key, "module": "NQ", "shape": [2, 3],
                                        "orientation": "rowwise" if s.ndim == 2 else "scalar",
                                        "s_kind": "kat", "lam": lam})

        # ---------------- STE-only variants (CL:49-64, PI:49-66) ----------------
        for mk, lam in (("CL", None), ("CL_MNIST", None), ("PI", 1e-3)):
            for shape, orient in (((3, 3, 8, 16), "channelwise"), ((24, 10), "columnwise"), ((16,), "scalar")):
                ctor = {"penalty_rate": 0.0} if mk.startswith("CL") else {"penalty_threshold": 0.0}
                s_shape, _ = scale_shape_of(mods[mk], shape, orient, **ctor)
                P, s, dy = make_inputs(rng, shape, s_shape, "trained")
                q, y, ds = run_op(mods[mk], P, s, dy, lam)
                cid = f"ste{len(manifest['ste']):03d}"
                for n, a in (("P", P), ("s", s), ("dy", dy), ("q", q), ("y", y), ("ds", ds)):
                    arrays[f"{cid}/{n}"] = a
                manifest["ste"].append({"id": cid, "module": mk, "shape": list(shape), "orientation": orient})

        # ---------------- custom loss terms (LF:33-275) ----------------
        def fake_layers(specs, mnist):
            layers, raw = [], []
            for (K, sK, b, sb) in specs:
                t = [torch.tensor(a, requires_grad=True) for a in (K, sK, b, sb)]
                ns = types.SimpleNamespace
                if mnist:
                    layers.append(ns(W=t[0], nested_q_w_layer=ns(scale=t[1]), b=t[2], nested_q_b_layer=ns(scale=t[3])))
                else:
                    layers.append(ns(kernel=t[0], nested_q_k_layer=ns(scale=t[1]), b=t[2], nested_q_b_layer=ns(scale=t[3])))
                raw.append(t)
            return layers, raw

        def layer_set(kind):
            specs = []
            if kind == "kat":
                specs.append((np.array([[0.30, -0.26, 0.04], [-0.05, 0.0, 0.11]], np.float32),
                              np.array([[0.1], [0.02]], np.float32),
                              np.array([0.05, -0.2], np.float32), np.array([0.1], np.float32)))
                return specs
            combos = [((3, 3, 8, 16), "channelwise"), ((3, 3, 16, 16), "rowwise"), ((1, 1, 8, 16), "columnwise"),
                      ((3, 3, 4, 8), "scalar"), ((24, 10), "columnwise"), ((24, 10), "channelwise")]
            for shape, orient in combos:
                s_shape, _ = scale_shape_of(mods["CL"], shape, orient, penalty_rate=0.0)
                K, sK, _ = make_inputs(rng, shape, s_shape, "init" if kind == "init" else "trained")
                b, sb, _ = make_inputs(rng, (shape[-1],), (1,), "init" if kind == "init" else "trained")
                if kind == "ties":
                    # duplicate the arg-max of every group (once with opposite sign) -> even gradient split
                    flat = K.reshape(-1)
                    j = int(np.argmax(np.abs(flat / np.broadcast_to(sK if sK.ndim == K.ndim else sK.reshape((1,) * K.ndim), K.shape).reshape(-1))))
                    stride = 1
                    flat[(j + stride) % flat.size] = -flat[j] if sK.size == 1 else flat[(j + stride) % flat.size]
                    b[1] = -b[int(np.argmax(np.abs(b)))] if int(np.argmax(np.abs(b))) != 1 else b[1]
                specs.append((K, sK, b, sb))
            return specs

        for kind in ("kat", "trained", "init", "ties"):
            specs = layer_set(kind)
            for mk, mnist in (("LF", False), ("LF_MNIST", True)):
                if mnist and kind not in ("kat", "trained"):
                    continue
                for cls, meth in (("SCCEMaxBin", "compute_maxbin_penalty"),
                                  ("SCCEDifference", "compute_difference_penalty"),
                                  ("SCCEInverse", "compute_inverse_penalty")):
                    layers, raw = fake_layers(specs, mnist)
                    gamma = 1e-9
                    obj = getattr(mods[mk], cls)(layers=layers, penalty_rate=gamma, log_dir=os.path.join(tmp, "ld"))
                    pen = getattr(obj, meth)()
                    pen.backward()
                    cid = f"loss{len(manifest['loss']):03d}"
                    arrays[f"{cid}/penalty"] = np32(pen)
                    for li, (t, spec) in enumerate(zip(raw, specs)):
                        for n, tt, a in zip(("K", "sK", "b", "sb"), t, spec):
                            arrays[f"{cid}/L{li}/{n}"] = a
                            g = tt.grad if tt.grad is not None else torch.zeros_like(tt)
                            arrays[f"{cid}/L{li}/g{n}"] = np32(g)
                    # total loss path (LF:47-73): mean over batch of scce + gamma*penalty
                    layers2, _ = fake_layers(specs, mnist)
                    obj2 = getattr(mods[mk], cls)(layers=layers2, penalty_rate=gamma, log_dir=os.path.join(tmp, "ld"))
                    y_true = np.array([1, 0, 3], np.int64)
                    y_pred = np.array([[0.1, 0.7, 0.1, 0.1], [0.25, 0.25, 0.25, 0.25], [0.05, 0.05, 0.1, 0.8]], np.float32)
                    tot = obj2.compute_total_loss(y_true, y_pred)
                    arrays[f"{cid}/y_true"], arrays[f"{cid}/y_pred"] = y_true, y_pred
                    arrays[f"{cid}/total_loss"] = np32(tot)
                    manifest["loss"].append({"id": cid, "module": mk, "class": cls, "method": meth, "set": kind,
                                             "n_layers": len(specs), "gamma": gamma})

        # inverse with an exactly-zero scale (LF:255-256 guard)
        specs = [(np.array([[0.3, -0.2], [0.1, 0.4]], np.float32), np.array([[0.0], [0.02]], np.float32),
                  np.array([0.05, -0.2], np.float32), np.array([0.0], np.float32))]
        layers, raw = fake_layers(specs, False)
        obj = mods["LF"].SCCEInverse(layers=layers, penalty_rate=1e-9, log_dir=os.path.join(tmp, "ld"))
        pen = obj.compute_inverse_penalty()
        pen.backward()
        cid = f"loss{len(manifest['loss']):03d}"
        arrays[f"{cid}/penalty"] = np32(pen)
        for n, tt, a in zip(("K", "sK", "b", "sb"), raw[0], specs[0]):
            arrays[f"{cid}/L0/{n}"] = a
            arrays[f"{cid}/L0/g{n}"] = np32(tt.grad if tt.grad is not None else torch.zeros_like(tt))
        manifest["loss"].append({"id": cid, "module": "LF", "class": "SCCEInverse", "method": "compute_inverse_penalty",
                                 "set": "zero_scale", "n_layers": 1, "gamma": 1e-9, "no_total": True})

        # ---------------- host layers (NQ:203-350, PI:299-365) ----------------
        def run_layer(mk, cls, ctor, x, weights, scales, name):
            mod = mods[mk]
            layer = getattr(mod, cls)(**ctor, trained_weights=[w for w in weights])
            xt = torch.tensor(x, requires_grad=True)
            layer(xt)   # build
            q_layers = [n for n in ("nested_q_w_layer", "nested_q_k_layer", "nested_q_b_layer") if hasattr(layer, n)]
            with torch.no_grad():
                for n, sv in zip(q_layers, scales):
                    getattr(layer, n).scale.copy_(torch.tensor(sv).reshape(getattr(layer, n).scale.shape))
            out = layer(xt)
            G = torch.tensor(rng.normal(0, 1e-3, size=tuple(out.shape)).astype(np.float32))
            (out * G).sum().backward()
            cid = f"layer{len(manifest['layers']):03d}"
            arrays[f"{cid}/x"], arrays[f"{cid}/G"], arrays[f"{cid}/out"] = x, np32(G), np32(out)
            arrays[f"{cid}/dx"] = np32(xt.grad)
            params = [layer.W if hasattr(layer, "W") else layer.kernel] + ([layer.b] if hasattr(layer, "b") else [])
            for i, p in enumerate(params):
                arrays[f"{cid}/w{i}"], arrays[f"{cid}/dw{i}"] = np32(p), np32(p.grad)
            for i, n in enumerate(q_layers):
                sc = getattr(layer, n).scale
                arrays[f"{cid}/s{i}"] = np32(sc)
                arrays[f"{cid}/ds{i}"] = np32(sc.grad if sc.grad is not None else torch.zeros_like(sc))
            manifest["layers"].append({"id": cid, "module": mk, "class": cls, "name": name, "auto_name": layer.name,
                                       "ctor": {k: (list(v) if isinstance(v, tuple) else v) for k, v in ctor.items()
                                                if k not in ("initializer", "regularizer")}})

        for orient in ORIENTS:
            W = rng.normal(0, 0.05, size=(24, 10)).astype(np.float32)
            b = rng.normal(0, 0.05, size=(10,)).astype(np.float32)
            s_shape, _ = scale_shape_of(mods["NQ"], (24, 10), orient, penalty_threshold=0.0)
            sW = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
            sb = rng.uniform(5e-3, 2e-2, size=(1,)).astype(np.float32)
            x = rng.uniform(0, 4, size=(4, 24)).astype(np.float32)
            run_layer("NQ", "CustomDenseLayer",
                      dict(seed=1, units=10, penalty_threshold=1e-3, orientation=orient, initializer=None,
                           name="custom_dense_layer_1", regularizer=None), x, [W, b], [sW, sb], "dense")
            K = rng.normal(0, 0.05, size=(3, 3, 8, 16)).astype(np.float32)
            bk = rng.normal(0, 0.05, size=(16,)).astype(np.float32)
            s_shape, _ = scale_shape_of(mods["NQ"], K.shape, orient, penalty_threshold=0.0)
            sK = rng.uniform(5e-3, 2e-2, size=s_shape).astype(np.float32)
            x = rng.uniform(0, 4, size=(2, 9, 10, 8)).astype(np.float32)
            run_layer("NQ", "CustomConv2DLayer",
                      dict(seed=1, penalty_threshold=1e-3, orientation=orient, initializer=None, filters=16,
                           kernel_size=(3, 3), strides=(2, 2), padding="same", name="custom_conv2d_layer_x",
                           regularizer=None), x, [K, bk], [sK, sb], "conv_same_s2")
            run_layer("CL", "CustomConv2DLayer",
                      dict(seed=1, penalty_rate=1e-9, orientation=orient, initializer=None, filters=16,
                           kernel_size=(3, 3), strides=(1, 1), padding="valid", name="custom_conv2d_layer_y",
                           regularizer=None), x, [K, bk], [sK, sb], "conv_valid_s1_cl")
            run_layer("PI", "CustomConv2DLayerNoBias",
                      dict(seed=1, penalty_threshold=1e-3, orientation=orient, initializer=None, filters=16,
                           kernel_size=(3, 3), strides=(1, 1), padding="same", name="conv1",
                           regularizer=None), x, [K], [sK], "conv_nobias_pi")

        # ---------------- MinValueConstraint (NQ:35-46) ----------------
        w = np.array([0.0, 1e-7, 1.1920929e-05, 1.2e-05, -3.0, 0.5], np.float32)
        c = mods["NQ"].MinValueConstraint(np.finfo(np.float32).eps * 100)
        arrays["constraint/in"], arrays["constraint/out"] = w, np32(c(torch.tensor(w)))
        manifest["constraint"].append({"min_value": float(np.finfo(np.float32).eps * 100),
                                       "config": c.get_config()})
    finally:
        os.chdir(cwd)

    manifest["generator"] = "tests/golden/make_golden.py on tests/golden/tf_shim.py (torch %s CPU)" % torch.__version__
    manifest["reference_files"] = FILES
    np.savez_compressed(out_npz, **arrays)
    with open(out_json, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True, default=float)
    print("wrote", out_npz, os.path.getsize(out_npz), "bytes;", {k: len(v) for k, v in manifest.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
