"""ctypes binding of csrc/liblq.so (the C-ABI declared in include/lq.h).

There is no fallback of any kind: if the shared library has not been built the import
fails loudly, and every compute entry point fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LQ_LIBRARY points at an alternative build of the same ABI (kernel A/B experiments); default: the in-tree build
LIB_PATH = os.environ.get("LQ_LIBRARY") or os.path.join(_HERE, "csrc", "liblq.so")

LQ_OK = 0
LQ_E_NULLPTR, LQ_E_BADSHAPE, LQ_E_BADMODE, LQ_E_WORKSPACE, LQ_E_CUDA, LQ_E_NCCL = -1, -2, -3, -4, -5, -6
LQ_ROUND_FLOOR, LQ_ROUND_RNE = 0, 1
LQ_CODES_NONE, LQ_CODES_F32, LQ_CODES_I32, LQ_CODES_I8 = 0, 1, 2, 3
LQ_PEN_MAXBIN, LQ_PEN_DIFFERENCE, LQ_PEN_INVERSE = 0, 1, 2
LQ_SGRAD_NONE, LQ_SGRAD_THRESHOLD = 0, 1
LQ_PHASE_FWD, LQ_PHASE_BWD, LQ_PHASE_PENALTY_FWD, LQ_PHASE_PENALTY_BWD = 0, 1, 2, 3
LQ_F_ACCUM_DX, LQ_F_ACCUM_DS = 1, 2
STATUS_NAMES = {0: "LQ_OK", -1: "LQ_E_NULLPTR", -2: "LQ_E_BADSHAPE", -3: "LQ_E_BADMODE",
                -4: "LQ_E_WORKSPACE", -5: "LQ_E_CUDA", -6: "LQ_E_NCCL"}


class TensorDesc(C.Structure):
    """struct lq_tensor_desc (include/lq.h)."""
    _fields_ = [("P", C.c_void_p), ("s", C.c_void_p), ("y", C.c_void_p), ("dy", C.c_void_p),
                ("dX", C.c_void_p), ("ds", C.c_void_p),
                ("outer", C.c_int64), ("C", C.c_int64), ("inner", C.c_int64),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class OptDesc(C.Structure):
    """struct lq_opt_desc (include/lq.h)."""
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p),
                ("n", C.c_int64), ("l2", C.c_float), ("min_value", C.c_float), ("weight_decay", C.c_float),
                ("reserved", C.c_uint32)]


class LqError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


_P, _I, _I64, _F, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_DESCS = C.POINTER(TensorDesc)

# name -> (restype, argtypes); one entry per declaration in include/lq.h
SIGNATURES = {
    "lq_version": (_I, []),
    "lq_last_error": (C.c_char_p, []),
    "lq_device_sm_count": (_I, []),
    "lq_workspace_bytes": (_I, [_DESCS, _I, _I, C.POINTER(_SZ)]),
    "lq_fq_fwd": (_I, [_P, _P, _P, _P, _I, _I64, _I64, _I64, _I, _F, _F, _P]),
    "lq_fq_fwd_stats": (_I, [_P, _P, _P, _P, _I, _I64, _I64, _I64, _I, _F, _F, _P, _P, _SZ, _P]),
    "lq_fq_bwd_threshold": (_I, [_P, _P, _P, _F, _P, _P, _I64, _I64, _I64, _I, _F, _F, _P, _SZ, _P]),
    "lq_fq_bwd_threshold_stats": (_I, [_P, _P, _P, _F, _P, _P, _I64, _I64, _I64, _I, _F, _F, _P, _SZ, _P]),
    "lq_fq_bwd_threshold_finish": (_I, [_P, _I, _I64, _F, C.c_double, _P, _P]),
    "lq_fq_bwd_ste": (_I, [_P, _P, _P, _I64, _I64, _P]),
    "lq_scale_constraint": (_I, [_P, _I64, _F, _P]),
    "lq_penalty_fwd": (_I, [_I, _DESCS, _I, _P, _P, _SZ, _P]),
    "lq_penalty_bwd": (_I, [_I, _DESCS, _I, _F, _P, _P, _SZ, _P]),
    "lq_mt_fq_fwd_penalty": (_I, [_I, _DESCS, _I, _I, _F, _F, _P, _P, _SZ, _P]),
    "lq_mt_fq_fwd": (_I, [_DESCS, _I, _I, _F, _F, _P]),
    "lq_mt_fq_bwd": (_I, [_DESCS, _I, _F, _I, _I, _F, _F, _P, _SZ, _P]),
    "lq_host_workspace_bytes": (_I, [_I64, _I64, _I64, C.POINTER(_SZ)]),
    "lq_host_fq_fwd_bwd": (_I, [_P, _P, _P, _F, _P, _P, _P, _I64, _I64, _I64, _P, _SZ]),
    "lq_host_create": (_I, [C.POINTER(_P)]),
    "lq_host_destroy": (_I, [_P]),
    "lq_host_fq_fwd_bwd_on": (_I, [_P, _P, _P, _P, _F, _P, _P, _P, _I64, _I64, _I64, _P, _SZ, _P, _I]),
    "lq_colsum_workspace_bytes": (_I, [_I64, _I64, C.POINTER(_SZ)]),
    "lq_colsum": (_I, [_P, _I64, _I64, _P, _P, _SZ, _P]),
    "lq_code_range": (_I, [_P, _P, _I64, _I64, _I64, _I, _F, _F, _P, _P, _SZ, _P]),
    "lq_code_histogram": (_I, [_P, _P, _I64, _I64, _I64, _I, _F, _F, _I64, _I64, _P, _P, _P]),
    "lq_opt_adam": (_I, [C.POINTER(OptDesc), _I, _F, _F, _F, _F, _I64, _P, _F, _P, _P]),
    "lq_opt_sgd": (_I, [C.POINTER(OptDesc), _I, _F, _F, _I, _F, _P, _P]),
    "lq_debug_divcheck": (_I, [_P, _I64, C.c_uint32, _P, _P]),
    "lq_debug_mt_geometry": (_I, [_DESCS, _I, _I, C.POINTER(_I64)]),
    "lq_p2p_create": (_I, [_I, _I, _I64, C.POINTER(_P), _P]),
    "lq_p2p_connect": (_I, [_P, _P]),
    "lq_p2p_buffer": (_I, [_P, C.POINTER(_P)]),
    "lq_p2p_connect_ptrs": (_I, [_P, C.POINTER(_P)]),
    "lq_fq_bwd_threshold_finish_p2p": (_I, [_P, _P, _I64, _F, C.c_double, _P, _P]),
    "lq_p2p_error": (_I, [_P, C.POINTER(_I)]),
    "lq_p2p_destroy": (_I, [_P]),
    "lq_dp_unique_id": (_I, [_P]),
    "lq_dp_init": (_I, [_P, _I, _I, C.POINTER(_P)]),
    "lq_dp_allreduce": (_I, [_P, _P, _I64, _F, _P]),
    "lq_dp_allreduce_max": (_I, [_P, _P, _I64, _P]),
    "lq_dp_allgather_f64": (_I, [_P, _P, _P, _I64, _P]),
    "lq_dp_destroy": (_I, [_P]),
    "lq_dpx_create": (_I, [_I, _I, _SZ, C.POINTER(_P), _P, C.POINTER(_P)]),
    "lq_dpx_connect": (_I, [_P, _P]),
    "lq_dpx_connect_ptrs": (_I, [_P, C.POINTER(_P)]),
    "lq_dpx_exchange": (_I, [_P, _DESCS, _I, _P, _I64, _F, _I, _I, _I, _P, _SZ, _P]),
    "lq_dpx_allreduce": (_I, [_P, _P, _I64, _I, _I, _P]),
    "lq_dpx_error": (_I, [_P, C.POINTER(_I)]),
    "lq_dpx_destroy": (_I, [_P]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `make -C "
            "learned_quantization_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library mismatch
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()


def check(status: int) -> None:
    if status != LQ_OK:
        raise LqError(status, lib.lq_last_error().decode("utf-8", "replace"))


def version() -> int:
    return lib.lq_version()
