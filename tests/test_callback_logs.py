"""CPU tier: the callback-log fixture (tests/golden/callback_logs.*, written by the reference's own
custom_callbacks.py on the TF shim) against (a) the oracle's integer-code helpers and (b) -- where the reference
tree is present (the dev container; the GPU box has none) -- the reference's own log parsers in
utils/plot_scripts.py ("the reference's plot scripts run on our outputs").  The drop-in callbacks
are held to the same fixture byte for byte on the GPU (tests/test_gpu_training.py), so what is shown here for the
fixture holds for their files."""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest

from oracle import lq_oracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LQ_REFERENCE_ROOT", "/root/reference")
PLOT = os.path.join(REF, "IMAGENETTE/nested_quantization_layer/utils/plot_scripts.py")


@pytest.fixture(scope="module")
def fx():
    with open(os.path.join(HERE, "golden", "callback_logs.json")) as f:
        return json.load(f), np.load(os.path.join(HERE, "golden", "callback_logs.npz"))


def _find(files, prefix, folder):
    hits = [k for k in files if k.startswith(folder + "/" + prefix)]
    assert len(hits) == 1, (prefix, hits)
    return files[hits[0]]


@pytest.mark.parametrize("kind", ["conv", "dense"])
def test_fixture_numbers_are_the_oracles(fx, kind):
    meta, arr = fx
    files = meta["files"][kind]
    K, sK, b, sb = (arr[f"{kind}/stage2/{n}"] for n in ("K", "sK", "b", "sb"))       # parameters at train end
    qK, _ = o.fq_forward(K, sK)
    qb, _ = o.fq_forward(b, sb)
    name = "Kernel" if kind == "conv" else "Weights"
    vals = [float(l) for l in _find(files, f"Quantized_{name}", "on_train_end").split()]
    assert np.array_equal(np.asarray(vals, np.float32), qK.reshape(-1))
    uniq = [l.split(", ") for l in _find(files, f"Unique_quantized_{name}", "on_train_end").strip().split("\n")]
    u, c = np.unique(qK.reshape(-1), return_counts=True)
    assert np.array_equal(np.asarray([float(x[0]) for x in uniq], np.float32), u) and [int(x[1]) for x in uniq] == list(c)
    n_unique, max_abs = o.code_stats(b, sb)
    last = _find(files, "Number_of_unique_Bias", "on_epoch_end").strip().split("\n")
    assert last[-2] == "Epoch 1" and int(last[-1]) == n_unique
    mx = _find(files, "Max_Bias", "on_epoch_end").strip().split("\n")
    assert float(mx[-1]) == max_abs == float(np.max(np.abs(qb)))


@pytest.mark.skipif(not os.path.exists(PLOT), reason="the reference tree is not present on this machine")
@pytest.mark.parametrize("kind", ["conv", "dense"])
def test_reference_plot_parsers_accept_the_tree(fx, kind, tmp_path):
    meta, arr = fx
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):          # not installed; the parsers do not plot
        m = types.ModuleType(name)
        m.use = lambda *a, **k: None
        sys.modules.setdefault(name, m)
    spec = importlib.util.spec_from_file_location("ref_plot_scripts", PLOT)
    ps = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ps)
    run = tmp_path / "run"
    for rel, text in meta["files"][kind].items():
        p = run / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(text)
    K, sK, b, sb = (arr[f"{kind}/stage2/{n}"] for n in ("K", "sK", "b", "sb"))
    qK, _ = o.fq_forward(K, sK)
    qb, _ = o.fq_forward(b, sb)
    both = np.concatenate([qK.reshape(-1), qb.reshape(-1)])
    assert ps.unique_values_on_end(str(run)) == len(np.unique(both))                 # plot_scripts.py:230-247
    assert ps.value_range_on_end(str(run)) == float(np.max(np.abs(both)))            # :250-268
    assert ps.final_val_acc_on_end(str(run)) == 0.625 and ps.final_val_loss_on_end(str(run)) == 1.0
    per_epoch = ps.process_file_logged_per_epoch(                                    # :13-38
        str(run / "on_epoch_end" / [k for k in meta["files"][kind] if k.startswith("on_epoch_end/Number_of_unique_Bias")][0].split("/")[1]))
    assert [len(e) for e in per_epoch] == [1, 1]
    for epoch in (0, 1):
        bq, _ = o.fq_forward(arr[f"{kind}/stage{epoch + 1}/b"], arr[f"{kind}/stage{epoch + 1}/sb"])
        assert per_epoch[epoch][0] == len(np.unique(bq))
