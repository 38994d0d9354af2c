"""ctypes loader for oracle/liblq_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liblq_oracle.so")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def load():
    if not os.path.exists(_PATH):
        raise ImportError(f"{_PATH} not built: run `make -C oracle`")
    lib = C.CDLL(_PATH)
    lib.lqo_threads.restype = C.c_int
    lib.lqo_fq_forward.argtypes = [_fp, _fp, C.c_int64, C.c_int64, C.c_int64, _fp, _fp, _fp]
    lib.lqo_fq_forward.restype = None
    lib.lqo_nested_backward.argtypes = [_fp, _fp, _fp, C.c_float, C.c_int64, C.c_int64, C.c_int64, _fp, _fp]
    lib.lqo_nested_backward.restype = None
    for name in ("lqo_maxbin_term", "lqo_difference_term"):
        getattr(lib, name).argtypes = [_fp, _fp, C.c_int64, C.c_int64, C.c_int64]
        getattr(lib, name).restype = C.c_double
    lib.lqo_inverse_term.argtypes = [_fp, C.c_int64, C.c_int64]
    lib.lqo_inverse_term.restype = C.c_double
    return lib


class COracle:
    def __init__(self):
        self.lib = load()
        self.threads = self.lib.lqo_threads()
        self._scratch = None

    def _buf(self, n):
        if self._scratch is None or self._scratch.size < n:
            self._scratch = np.empty(n, np.float32)
        return self._scratch

    def fwd_bwd(self, P, s, dy, layout, lam, out=None):
        """(q, y, ds) for flat fp32 arrays viewed as layout = (outer, C, inner).  `out` = (q, y) buffers to
        reuse (a timed loop passes them so that fresh pages are not faulted in on every pass -- TensorFlow's
        allocator pools its output tensors too)."""
        n = P.size
        q, y = out if out is not None else (np.empty(n, np.float32), np.empty(n, np.float32))
        ds = np.empty(layout[1], np.float32)
        scratch = self._buf(5 * n)
        self.lib.lqo_fq_forward(P, s, *layout, scratch, q, y)
        self.lib.lqo_nested_backward(dy, q, y, np.float32(lam), *layout, ds, scratch)
        return q, y, ds

    def penalty(self, kind, tensors):
        """tensors: list of (P flat, s flat, layout). Returns the penalty value."""
        total, nrm = 0.0, 0.0
        for P, s, layout in tensors:
            n = layout[0] * layout[1] * layout[2]
            if kind == "maxbin":
                total += self.lib.lqo_maxbin_term(P, s, *layout)
            elif kind == "difference":
                total += self.lib.lqo_difference_term(P, s, *layout)
            else:
                total += self.lib.lqo_inverse_term(s, layout[1], n)
            nrm += n
        return np.float32(total / nrm)
