"""Drop-in for the int8 export of */utils/log_scripts.py (`save_compress_parameters`, :61-97), with
the integer codes produced on the device (lq_fq_fwd, codes_dtype = int8) instead of pulling fp32
weights to the host, and with the saturation check the reference lacks.

Same artefacts: `weights.npy` (a pickled dict name -> int8 array, np.save of a dict like the
reference), `weights.zip` (ZIP_DEFLATED) and `file_sizes.log` with the reference's two lines; one
extra file `int8_overflow.log` lists tensors whose codes do not fit int8 (they wrap, exactly as
`.astype(np.int8)` does in the reference).
"""
from __future__ import annotations

import os
import zipfile

import numpy as np

from .. import ops
from ..custom_components import _impl


def quantized_parameters(model):
    """{name/W, name/b -> int8 codes} for every quantised layer, in module order (log_scripts.py:70-80)."""
    weights, overflow = {}, {}
    for m in model.modules():
        if isinstance(m, _impl.Conv2DLayerBase) and m.built:
            pairs = [("W", m.kernel, m.nested_q_k_layer)] + ([("b", m.b, m.nested_q_b_layer)] if m._use_bias else [])
        elif isinstance(m, _impl.DenseLayerBase) and m.built:
            pairs = [("W", m.W, m.nested_q_w_layer), ("b", m.b, m.nested_q_b_layer)]
        else:
            continue
        for tag, p, q in pairs:
            P, s = ops.dense(p.detach()), q.scale.detach().contiguous()
            weights[f"{m.name}/{tag}"] = ops.codes_int8(P, s).cpu().numpy()
            st = ops.code_stats(P, s)
            if st["int8_overflow"]:
                overflow[f"{m.name}/{tag}"] = (st["int8_overflow"], st["min"], st["max"])
    return weights, overflow


def save_compress_parameters(model, log_dir: str) -> None:
    os.makedirs(log_dir, exist_ok=True)
    weights_path = os.path.join(log_dir, "weights.npy")
    weights, overflow = quantized_parameters(model)
    np.save(weights_path, weights)                                       # :81
    zip_file_path = os.path.join(log_dir, "weights.zip")
    with zipfile.ZipFile(zip_file_path, "w", compression=zipfile.ZIP_DEFLATED) as zipf:
        zipf.write(weights_path, arcname="weights.npy")                  # :84-86
    size = os.path.getsize(weights_path) / (1024 * 1024)
    zip_size = os.path.getsize(zip_file_path) / (1024 * 1024)
    with open(os.path.join(log_dir, "file_sizes.log"), "w") as log_file:  # :92-95
        log_file.write(f"Weights size: {size:.4f} MB\n")
        log_file.write(f"Compressed weights size: {zip_size:.4f} MB\n")
    with open(os.path.join(log_dir, "int8_overflow.log"), "w") as f:
        for name, (cnt, lo, hi) in overflow.items():
            f.write(f"{name}: {cnt} codes outside [-128, 127] (code range {lo:g} .. {hi:g}) wrap in the int8 export\n")
