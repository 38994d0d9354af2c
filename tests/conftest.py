import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """Committed outputs of the reference's own source (see tests/golden/make_golden.py)."""

    def __init__(self):
        d = os.path.join(ROOT, "tests", "golden")
        self.arrays = np.load(os.path.join(d, "reference_vectors.npz"))
        with open(os.path.join(d, "reference_vectors.json")) as f:
            self.manifest = json.load(f)

    def __getitem__(self, k):
        return self.arrays[k]

    def op_cases(self):
        for c in self.manifest["ops"]:
            i = c["inputs"]
            yield c, self[f"{i}/P"], self[f"{i}/s"], self[f"{i}/dy"], self[f"{c['id']}/q"], self[f"{c['id']}/y"], self[f"{c['id']}/ds"]

    def loss_layers(self, c):
        out = []
        for li in range(c["n_layers"]):
            out.append(tuple(self[f"{c['id']}/L{li}/{n}"] for n in ("K", "sK", "b", "sb")))
        return out

    def loss_grads(self, c):
        out = []
        for li in range(c["n_layers"]):
            out.append(tuple(self[f"{c['id']}/L{li}/g{n}"] for n in ("K", "sK", "b", "sb")))
        return out


@pytest.fixture(scope="session")
def golden():
    return Golden()


def rel_err(a, b, tiny=1e-30):
    """max |a-b| / max(|b|, tiny) with +-0 equal and NaN==NaN."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.where(both_nan, 0.0, np.abs(a - b) / np.maximum(np.abs(b), tiny))
    d = np.where((a == 0) & (b == 0), 0.0, d)
    return float(np.max(d)) if d.size else 0.0


def bits_equal(a, b):
    """Bit-exact fp32 comparison that treats +0/-0 as DIFFERENT and NaN payloads as equal."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    if a.shape != b.shape:
        return False
    same = a.view(np.uint32) == b.view(np.uint32)
    same |= np.isnan(a) & np.isnan(b)
    return bool(np.all(same))
