"""Thin torch-side wrappers over the C-ABI: raw device pointers + the current stream.

torch is used for device memory, streams and autograd plumbing only; all arithmetic of
the hot path happens inside liblq.so.  Nothing here falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import TensorDesc, check, lib

EPS_FLOAT32 = 1.1920928955078125e-07           # np.finfo(np.float32).eps (NQ:11)
SCALE_INIT = 1.1920928955078125e-05            # eps_float32 * 100 in fp32 (NQ:156-158)
NEG_INF, POS_INF = -math.inf, math.inf


# --------------------------------------------------------------------- host logic
def scale_shape(param_shape: Sequence[int], orientation: str) -> Tuple[int, ...]:
    """Scale shape per orientation, CustomQuantizedScaleLayer.build (NQ:147-197)."""
    n = len(param_shape)
    if orientation == "rowwise":
        return tuple(int(param_shape[i]) if i == 0 else 1 for i in range(n))
    if orientation == "columnwise":
        return tuple(int(param_shape[i]) if i == 1 else 1 for i in range(n))
    if orientation == "channelwise":
        return tuple(int(param_shape[i]) if i == 2 else 1 for i in range(n))
    if orientation == "scalar":
        return (1,)
    raise ValueError(
        f"Invalid scaler application: {orientation}. Expected rowwise, columnwise or scalar.")


def group_layout(param_shape: Sequence[int], s_shape: Sequence[int]) -> Tuple[int, int, int]:
    """[outer, C, inner] view with group(idx) = (idx // inner) % C for a broadcastable scale."""
    param_shape = tuple(int(d) for d in param_shape)
    s_shape = tuple(int(d) for d in s_shape)
    total = 1
    for d in param_shape:
        total *= d
    if len(s_shape) == 1 and len(param_shape) != 1:
        if s_shape[0] != 1:
            raise ValueError(f"scale shape {s_shape} does not broadcast over parameter shape {param_shape}")
        return 1, 1, total
    if len(s_shape) != len(param_shape):
        raise ValueError(f"scale shape {s_shape} does not broadcast over parameter shape {param_shape}")
    axes = [i for i, d in enumerate(s_shape) if d != 1]
    if len(axes) > 1:
        raise ValueError(f"scale shape {s_shape}: at most one non-1 axis is supported")
    if not axes:
        return 1, 1, total
    ax = axes[0]
    if s_shape[ax] != param_shape[ax]:
        raise ValueError(f"scale shape {s_shape} does not broadcast over parameter shape {param_shape}")
    outer = 1
    for d in param_shape[:ax]:
        outer *= d
    inner = 1
    for d in param_shape[ax + 1:]:
        inner *= d
    return outer, param_shape[ax], inner


def mem_perm(t: torch.Tensor) -> Optional[Tuple[int, ...]]:
    """Dims of `t` ordered from the slowest to the fastest varying in MEMORY, if `t` is a permutation of
    a contiguous tensor (e.g. a conv kernel of logical shape HWIO stored as OHWI, the order cuDNN's
    channels-last kernels want); None otherwise.  Contiguous tensors give the identity."""
    if t.is_contiguous():
        return tuple(range(t.dim()))
    perm = tuple(sorted(range(t.dim()), key=lambda d: (-t.stride(d), d)))
    return perm if t.permute(perm).is_contiguous() else None


def dense(t: torch.Tensor) -> torch.Tensor:
    """`t` itself when its elements occupy one dense block of memory (in any dim order), else a contiguous copy."""
    return t if mem_perm(t) is not None else t.contiguous()


def same_layout(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Same shape and the same element at every memory offset (strides of size-1 dims do not matter)."""
    return a.shape == b.shape and all(sa == sb or n == 1 for sa, sb, n in zip(a.stride(), b.stride(), a.shape))


def match_strides(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """`t` laid out in memory exactly like `like` (same shape): a no-op when it already is."""
    if same_layout(t, like):
        return t
    return torch.empty_like(like).copy_(t)


def _require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("learned_quantization_b200 runs on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")
        if t.dtype != torch.float32:
            raise TypeError(f"fp32 tensors only, got {t.dtype}")
        if mem_perm(t) is None:
            raise ValueError("tensors must be dense in memory (contiguous, or a permutation of a contiguous tensor)")


def tensor_layout(P: torch.Tensor, s_shape: Sequence[int]) -> Tuple[int, int, int]:
    """[outer, C, inner] of `P` AS IT LIES IN MEMORY for a scale of (logical) shape `s_shape`: the groups are
    the same as group_layout's, only their position in the flat memory order changes when `P` is a permuted
    view.  HWIO kernel stored as OHWI: channelwise -> [Co*kh*kw, Ci, 1], rowwise -> [Co, kh, kw*Ci],
    columnwise -> [Co*kh, kw, Ci]."""
    perm = mem_perm(P)
    if perm is None:
        raise ValueError("tensor is not dense in memory")
    shape = tuple(int(d) for d in P.shape)
    s_shape = tuple(int(d) for d in s_shape)
    if perm == tuple(range(P.dim())):
        return group_layout(shape, s_shape)
    if len(s_shape) == 1 and len(shape) != 1 or all(d == 1 for d in s_shape):
        group_layout(shape, s_shape)                   # validates
        return 1, 1, P.numel()
    ax = [i for i, d in enumerate(s_shape) if d != 1]
    group_layout(shape, s_shape)                       # validates (one non-1 axis, sizes agree)
    k = perm.index(ax[0])
    mem_shape = [shape[d] for d in perm]
    outer = 1
    for d in mem_shape[:k]:
        outer *= d
    inner = 1
    for d in mem_shape[k + 1:]:
        inner *= d
    return outer, mem_shape[k], inner


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# One growable workspace per (device, stream): kernels on one stream are ordered, so
# consecutive calls may reuse it.
_WS: dict = {}


def workspace(nbytes: int, device: torch.device, tag: str = "") -> torch.Tensor:
    key = (device.index, _stream(), tag)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def make_desc(P=None, s=None, y=None, dy=None, dX=None, ds=None, layout=(0, 0, 0), flags=0) -> TensorDesc:
    ptr = lambda t: None if t is None else t.data_ptr()
    return TensorDesc(P=ptr(P), s=ptr(s), y=ptr(y), dy=ptr(dy), dX=ptr(dX), ds=ptr(ds),
                      outer=layout[0], C=layout[1], inner=layout[2], flags=flags, reserved=0)


def desc_array(descs: Sequence[TensorDesc]):
    return (TensorDesc * len(descs))(*descs)


def workspace_bytes(descs, phase: int) -> int:
    out = C.c_size_t()
    arr = descs if not isinstance(descs, (list, tuple)) else desc_array(descs)
    check(lib.lq_workspace_bytes(arr, len(arr), phase, C.byref(out)))
    return out.value


# --------------------------------------------------------------------- operators
def fq_fwd(P: torch.Tensor, s: torch.Tensor, layout=None, *, codes_dtype: Optional[torch.dtype] = None,
           round_mode: int = _lib.LQ_ROUND_FLOOR, qmin: float = NEG_INF, qmax: float = POS_INF,
           out: Optional[torch.Tensor] = None):
    """y = floor(P / s) * s (NQ:55-60).  Returns y, or (y, codes) when codes_dtype is given."""
    _require_cuda(P, s)
    layout = layout or tensor_layout(P, s.shape)
    y = torch.empty_like(P) if out is None else out          # same strides as P: element i of memory <-> element i
    if not same_layout(y, P):
        raise ValueError("out must be laid out in memory like P")
    codes, cd = None, _lib.LQ_CODES_NONE
    if codes_dtype is not None:
        cd = {torch.float32: _lib.LQ_CODES_F32, torch.int32: _lib.LQ_CODES_I32, torch.int8: _lib.LQ_CODES_I8}[codes_dtype]
        codes = torch.empty_like(P, dtype=codes_dtype)
    with torch.cuda.device(P.device):
        check(lib.lq_fq_fwd(P.data_ptr(), s.data_ptr(), y.data_ptr(),
                            None if codes is None else codes.data_ptr(), cd,
                            layout[0], layout[1], layout[2], round_mode, qmin, qmax, _stream()))
    return y if codes is None else (y, codes)


def fq_fwd_stats(P: torch.Tensor, s: torch.Tensor, layout=None, *, round_mode: int = _lib.LQ_ROUND_FLOOR,
                 qmin: float = NEG_INF, qmax: float = POS_INF, out: Optional[torch.Tensor] = None):
    """(y, stats): the forward and, from the same pass over P, stats[g] = {min P, max P, sum |P - P/s|} per group
    (float32 [C, 3]) -- what max |q|, the MaxBin term and the Difference term are functions of."""
    _require_cuda(P, s)
    layout = layout or tensor_layout(P, s.shape)
    y = torch.empty_like(P) if out is None else out
    if not same_layout(y, P):
        raise ValueError("out must be laid out in memory like P")
    stats = torch.empty((layout[1], 3), dtype=torch.float32, device=P.device)
    ws = workspace(workspace_bytes([make_desc(layout=layout)], _lib.LQ_PHASE_BWD), P.device)
    with torch.cuda.device(P.device):
        check(lib.lq_fq_fwd_stats(P.data_ptr(), s.data_ptr(), y.data_ptr(), None, _lib.LQ_CODES_NONE,
                                  layout[0], layout[1], layout[2], round_mode, qmin, qmax,
                                  stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return y, stats


def fq_bwd_threshold(dy: torch.Tensor, P: torch.Tensor, s: torch.Tensor, penalty_threshold: float,
                     layout=None, *, write_dx: bool = False,
                     round_mode: int = _lib.LQ_ROUND_FLOOR, qmin: float = NEG_INF, qmax: float = POS_INF):
    """(dX, ds) of NQ:62-118.  dX is dy itself unless write_dx asks for a separate copy."""
    _require_cuda(dy, P, s)
    layout = layout or tensor_layout(P, s.shape)
    dy = match_strides(dy, P)
    ds = torch.empty_like(s)
    dX = torch.empty_like(dy) if write_dx else None
    d = make_desc(layout=layout)
    need = workspace_bytes([d], _lib.LQ_PHASE_BWD)
    ws = workspace(need, P.device)
    with torch.cuda.device(P.device):
        check(lib.lq_fq_bwd_threshold(dy.data_ptr(), P.data_ptr(), s.data_ptr(), float(penalty_threshold),
                                      None if dX is None else dX.data_ptr(), ds.data_ptr(),
                                      layout[0], layout[1], layout[2], round_mode, qmin, qmax,
                                      ws.data_ptr(), ws.numel(), _stream()))
    return (dy if dX is None else dX), ds


def fq_bwd_threshold_stats(dy: torch.Tensor, P: torch.Tensor, s: torch.Tensor, penalty_threshold: float,
                           layout=None, *, write_dx: bool = False, stats: Optional[torch.Tensor] = None):
    """One SHARD of a tensor split along `outer`: (dX, stats) with stats[g] = {sum, n_below, max|q|}
    (float64 [C, 3]).  Gather the shards' stats and call fq_bwd_threshold_finish."""
    _require_cuda(dy, P, s)
    layout = layout or tensor_layout(P, s.shape)
    dy = match_strides(dy, P)
    if stats is None:
        stats = torch.empty((layout[1], 3), dtype=torch.float64, device=P.device)
    dX = torch.empty_like(dy) if write_dx else None
    need = workspace_bytes([make_desc(layout=layout)], _lib.LQ_PHASE_BWD)
    ws = workspace(need, P.device)
    with torch.cuda.device(P.device):
        check(lib.lq_fq_bwd_threshold_stats(dy.data_ptr(), P.data_ptr(), s.data_ptr(), float(penalty_threshold),
                                            None if dX is None else dX.data_ptr(), stats.data_ptr(),
                                            layout[0], layout[1], layout[2], _lib.LQ_ROUND_FLOOR, NEG_INF, POS_INF,
                                            ws.data_ptr(), ws.numel(), _stream()))
    return (dy if dX is None else dX), stats


def fq_bwd_threshold_finish(stats_all: torch.Tensor, penalty_threshold: float, n_per_group: float,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """ds from the gathered statistics of all shards (float64 [nshards, C, 3], shard order fixed)."""
    if not stats_all.is_cuda or stats_all.dtype != torch.float64 or not stats_all.is_contiguous() or stats_all.dim() != 3:
        raise ValueError("stats_all must be a contiguous float64 CUDA tensor [nshards, C, 3]")
    nshards, C = stats_all.shape[0], stats_all.shape[1]
    ds = torch.empty(C, dtype=torch.float32, device=stats_all.device) if out is None else out
    with torch.cuda.device(stats_all.device):
        check(lib.lq_fq_bwd_threshold_finish(stats_all.data_ptr(), nshards, C, float(penalty_threshold),
                                             float(n_per_group), ds.data_ptr(), _stream()))
    return ds


def fq_bwd_ste(dy: torch.Tensor, s: torch.Tensor):
    """(dX, ds) of CL:61-62 / PI:62-64: dX is dy, ds is zeros."""
    _require_cuda(dy, s)
    ds = torch.empty_like(s)
    with torch.cuda.device(dy.device):
        check(lib.lq_fq_bwd_ste(dy.data_ptr(), None, ds.data_ptr(), dy.numel(), s.numel(), _stream()))
    return dy, ds


def scale_constraint_(s: torch.Tensor, min_value: float = SCALE_INIT) -> torch.Tensor:
    """In-place MinValueConstraint (NQ:35-46)."""
    _require_cuda(s)
    with torch.cuda.device(s.device):
        check(lib.lq_scale_constraint(s.data_ptr(), s.numel(), float(min_value), _stream()))
    return s


# --------------------------------------------------------------------- multi-tensor
def mt_fq_fwd(Ps: Sequence[torch.Tensor], ss: Sequence[torch.Tensor], ys: Sequence[torch.Tensor],
              layouts=None, round_mode=_lib.LQ_ROUND_FLOOR, qmin=NEG_INF, qmax=POS_INF) -> None:
    _require_cuda(*Ps, *ss, *ys)
    layouts = layouts or [tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    if any(not same_layout(y, p) for y, p in zip(ys, Ps)):
        raise ValueError("every y must be laid out in memory like its P")
    arr = desc_array([make_desc(P=p, s=s, y=y, layout=l) for p, s, y, l in zip(Ps, ss, ys, layouts)])
    with torch.cuda.device(Ps[0].device):
        check(lib.lq_mt_fq_fwd(arr, len(arr), round_mode, qmin, qmax, _stream()))


def mt_fq_bwd(dys, Ps, ss, dss, penalty_threshold: float, layouts=None, dXs=None,
              scale_grad_mode=_lib.LQ_SGRAD_THRESHOLD, round_mode=_lib.LQ_ROUND_FLOOR,
              qmin=NEG_INF, qmax=POS_INF) -> None:
    _require_cuda(*dys, *Ps, *ss, *dss)
    layouts = layouts or [tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    if any(not same_layout(dy, p) for dy, p in zip(dys, Ps)):
        raise ValueError("every dy must be laid out in memory like its P (ops.match_strides)")
    dXs = dXs or [None] * len(Ps)
    arr = desc_array([make_desc(P=p, s=s, dy=dy, dX=dx, ds=ds, layout=l)
                      for p, s, dy, dx, ds, l in zip(Ps, ss, dys, dXs, dss, layouts)])
    need = workspace_bytes(arr, _lib.LQ_PHASE_BWD)
    ws = workspace(need, Ps[0].device)
    with torch.cuda.device(Ps[0].device):
        check(lib.lq_mt_fq_bwd(arr, len(arr), float(penalty_threshold), scale_grad_mode, round_mode, qmin, qmax,
                               ws.data_ptr(), ws.numel(), _stream()))


# --------------------------------------------------------------------- penalties
def penalty_fwd(kind: int, Ps, ss, layouts=None):
    """compute_*_penalty() over tensors (LF:75-116 / 161-195 / 240-275). Returns (scalar, ws)."""
    _require_cuda(*Ps, *ss)
    layouts = layouts or [tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    arr = desc_array([make_desc(P=p, s=s, layout=l) for p, s, l in zip(Ps, ss, layouts)])
    need = workspace_bytes(arr, _lib.LQ_PHASE_PENALTY_FWD)
    ws = torch.empty(max(need, 256), dtype=torch.uint8, device=Ps[0].device)   # kept for the backward
    out = torch.empty((), dtype=torch.float32, device=Ps[0].device)
    with torch.cuda.device(Ps[0].device):
        check(lib.lq_penalty_fwd(kind, arr, len(arr), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out, ws


def mt_fq_fwd_penalty(kind: int, Ps, ss, ys, layouts=None, ws: Optional[torch.Tensor] = None,
                      out: Optional[torch.Tensor] = None, round_mode=_lib.LQ_ROUND_FLOOR, qmin=NEG_INF, qmax=POS_INF):
    """The forward of every tensor and the MaxBin / Difference term over them in one pass (lq_mt_fq_fwd_penalty):
    ys are written as mt_fq_fwd writes them; returns (scalar, ws) exactly as penalty_fwd does, for penalty_bwd.
    ws / out may be passed in (static buffers of a captured training step)."""
    _require_cuda(*Ps, *ss, *ys)
    layouts = layouts or [tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    if any(not same_layout(y, p) for y, p in zip(ys, Ps)):
        raise ValueError("every y must be laid out in memory like its P")
    arr = desc_array([make_desc(P=p, s=s, y=y, layout=l) for p, s, y, l in zip(Ps, ss, ys, layouts)])
    need = max(workspace_bytes(arr, _lib.LQ_PHASE_PENALTY_FWD), 256)
    if ws is None:
        ws = torch.empty(need, dtype=torch.uint8, device=Ps[0].device)          # kept for the backward
    elif ws.numel() < need:
        raise ValueError(f"workspace of {need} bytes needed, got {ws.numel()}")
    if out is None:
        out = torch.empty((), dtype=torch.float32, device=Ps[0].device)
    with torch.cuda.device(Ps[0].device):
        check(lib.lq_mt_fq_fwd_penalty(kind, arr, len(arr), round_mode, qmin, qmax, out.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _stream()))
    return out, ws


def penalty_bwd(kind: int, Ps, ss, ws: torch.Tensor, upstream: float = 1.0,
                upstream_dev: Optional[torch.Tensor] = None, layouts=None, dXs=None, dss=None,
                accumulate: bool = False):
    """Gradients of upstream*penalty w.r.t. every P and s. Returns (dXs, dss)."""
    _require_cuda(*Ps, *ss)
    layouts = layouts or [tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
    need_dx = kind != _lib.LQ_PEN_INVERSE
    if dXs is None:
        dXs = [torch.empty_like(p) if need_dx else None for p in Ps]
    if dss is None:
        dss = [torch.empty_like(s) for s in ss]
    flags = (_lib.LQ_F_ACCUM_DX | _lib.LQ_F_ACCUM_DS) if accumulate else 0
    arr = desc_array([make_desc(P=p, s=s, dX=dx, ds=ds, layout=l, flags=flags)
                      for p, s, dx, ds, l in zip(Ps, ss, dXs, dss, layouts)])
    with torch.cuda.device(Ps[0].device):
        check(lib.lq_penalty_bwd(kind, arr, len(arr), float(upstream),
                                 None if upstream_dev is None else upstream_dev.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()))
    return dXs, dss


# --------------------------------------------------------------------- host-buffer path
def host_workspace_bytes(layout) -> int:
    out = C.c_size_t()
    check(lib.lq_host_workspace_bytes(layout[0], layout[1], layout[2], C.byref(out)))
    return out.value


_HOST_HANDLES: dict = {}


def _host_handle(device: torch.device):
    """lq_host_create once per device: the pipeline's streams and events are reused by every call."""
    h = _HOST_HANDLES.get(device.index)
    if h is None:
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(lib.lq_host_create(C.byref(h)))
        _HOST_HANDLES[device.index] = h
    return h


def host_fq_fwd_bwd(hP: torch.Tensor, hs: torch.Tensor, hdy: torch.Tensor, penalty_threshold: float,
                    hy: torch.Tensor, hds: torch.Tensor, layout, dev_ws: torch.Tensor,
                    hdX: Optional[torch.Tensor] = None) -> None:
    """Forward + threshold backward on HOST tensors (pinned recommended); synchronous."""
    for t in (hP, hs, hdy, hy, hds):
        if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("host_fq_fwd_bwd takes contiguous fp32 CPU tensors")
    if not dev_ws.is_cuda:
        raise RuntimeError("dev_ws must be a CUDA buffer (no CPU fallback)")
    with torch.cuda.device(dev_ws.device):
        # ordered after the current stream: dev_ws is a caching-allocator block whose previous use was enqueued there
        check(lib.lq_host_fq_fwd_bwd_on(_host_handle(dev_ws.device), hP.data_ptr(), hs.data_ptr(), hdy.data_ptr(),
                                        float(penalty_threshold), hy.data_ptr(),
                                        None if hdX is None else hdX.data_ptr(), hds.data_ptr(),
                                        layout[0], layout[1], layout[2], dev_ws.data_ptr(), dev_ws.numel(), _stream(), 1))


# --------------------------------------------------------------------- optimizer step
class OptimizerTable:
    """The descriptor table of one optimizer over a fixed set of tensors: built once (pointers of
    parameters, gradients and slots are stable across steps), passed to lq_opt_* every step."""

    def __init__(self, params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], *, adam: bool,
                 momentum: bool = True, l2: Optional[Sequence[float]] = None,
                 min_value: Optional[Sequence[Optional[float]]] = None, weight_decay: float = 0.0):
        _require_cuda(*params, *grads)
        n = len(params)
        l2 = list(l2) if l2 is not None else [0.0] * n
        min_value = list(min_value) if min_value is not None else [None] * n
        self.params, self.grads = list(params), list(grads)
        self.m = [torch.zeros_like(p) for p in params] if (adam or momentum) else [None] * n
        self.v = [torch.zeros_like(p) for p in params] if adam else [None] * n
        ptr = lambda t: None if t is None else t.data_ptr()
        self.descs = (_lib.OptDesc * n)(*[
            _lib.OptDesc(p=p.data_ptr(), g=g.data_ptr(), m=ptr(m), v=ptr(v), n=p.numel(), l2=float(l),
                         min_value=(NEG_INF if mv is None else float(mv)), weight_decay=float(weight_decay), reserved=0)
            for p, g, m, v, l, mv in zip(params, grads, self.m, self.v, l2, min_value)])
        self.n = n
        self.device = params[0].device if n else None


def opt_adam(table: OptimizerTable, lr: float, step, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
             grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None) -> None:
    """Keras-2.11 Adam over every tensor of the table, with l2 / weight decay / MinValueConstraint
    folded in.  `step` is the 1-based iteration: a Python int, or a CUDA int64 tensor that the
    kernel reads (for CUDA-graph replays)."""
    step_dev = None
    if isinstance(step, torch.Tensor):
        if not step.is_cuda or step.dtype != torch.int64 or step.numel() != 1:
            raise TypeError("a device step counter must be a CUDA int64 scalar tensor")
        step_dev, step = step.data_ptr(), 0
    with torch.cuda.device(table.device):
        check(lib.lq_opt_adam(table.descs, table.n, float(lr), float(beta_1), float(beta_2), float(epsilon), int(step),
                              step_dev, float(grad_scale),
                              None if grad_scale_dev is None else grad_scale_dev.data_ptr(), _stream()))


def opt_sgd(table: OptimizerTable, lr: float, momentum: float = 0.0, nesterov: bool = False,
            grad_scale: float = 1.0, grad_scale_dev: Optional[torch.Tensor] = None) -> None:
    with torch.cuda.device(table.device):
        check(lib.lq_opt_sgd(table.descs, table.n, float(lr), float(momentum), int(bool(nesterov)),
                             float(grad_scale), None if grad_scale_dev is None else grad_scale_dev.data_ptr(), _stream()))


# --------------------------------------------------------------------- integer-code statistics
def code_stats(P: torch.Tensor, s: torch.Tensor, layout=None, *, max_bins: int = 1 << 24,
               round_mode: int = _lib.LQ_ROUND_FLOOR, qmin: float = NEG_INF, qmax: float = POS_INF) -> dict:
    """What custom_callbacks.py:84-129 recomputes on the host every epoch, computed on the device:
    np.unique(floor(P / s), return_counts=True), max |q|, and the number of codes that do not fit
    int8 (utils/log_scripts.py:74-79 casts with .astype(np.int8), which wraps silently).
    Two passes over P (range, then histogram); only the non-empty bins come back to the host."""
    _require_cuda(P, s)
    layout = layout or tensor_layout(P, s.shape)
    dev = P.device
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    ws = workspace(256, dev, "codes")
    with torch.cuda.device(dev):
        check(lib.lq_code_range(P.data_ptr(), s.data_ptr(), *layout, round_mode, qmin, qmax, mm.data_ptr(),
                                ws.data_ptr(), ws.numel(), _stream()))
    lo_f, hi_f = (float(v) for v in mm.cpu())
    n = P.numel()
    if n == 0 or math.isnan(lo_f):
        return {"unique": torch.empty(0), "counts": torch.empty(0, dtype=torch.int64), "n_unique": 0,
                "min": math.nan, "max": math.nan, "max_abs": math.nan, "outside_window": n, "int8_overflow": 0}
    lim = float(1 << 30) - 1
    lo = int(max(min(lo_f, lim), -lim))
    hi = int(max(min(hi_f, lim), -lim))
    nbins = min(hi - lo + 1, max_bins)
    hist = torch.zeros(nbins, dtype=torch.int32, device=dev)
    outside = torch.zeros(1, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(lib.lq_code_histogram(P.data_ptr(), s.data_ptr(), *layout, round_mode, qmin, qmax, lo, nbins,
                                    hist.data_ptr(), outside.data_ptr(), _stream()))
    nz = torch.nonzero(hist, as_tuple=False).flatten()
    counts = hist[nz].to(torch.int64)
    values = (nz + lo).to(torch.float32)
    in8 = (values >= -128) & (values <= 127)
    n_out = int(outside)
    return {"unique": values.cpu(), "counts": counts.cpu(), "n_unique": int(nz.numel()), "min": lo_f, "max": hi_f,
            "max_abs": max(abs(lo_f), abs(hi_f)), "outside_window": n_out,
            "int8_overflow": int(counts[~in8].sum()) + n_out}


def codes_int8(P: torch.Tensor, s: torch.Tensor, layout=None) -> torch.Tensor:
    """floor(P / s).astype(int8) with numpy's wrap-around (utils/log_scripts.py:74-79), on the device."""
    _, codes = fq_fwd(P, s, layout, codes_dtype=torch.int8)
    return codes


# --------------------------------------------------------------------- bias gradient
def colsum(x2d: torch.Tensor) -> torch.Tensor:
    """Column sums of a contiguous fp32 [rows, cols] CUDA matrix (the bias gradient of an NHWC convolution)."""
    if not x2d.is_cuda or x2d.dtype != torch.float32 or x2d.dim() != 2 or not x2d.is_contiguous():
        raise ValueError("colsum takes a contiguous fp32 CUDA matrix")
    rows, cols = x2d.shape
    need = C.c_size_t()
    check(lib.lq_colsum_workspace_bytes(rows, cols, C.byref(need)))
    ws = workspace(need.value, x2d.device, "colsum")
    out = torch.empty(cols, dtype=torch.float32, device=x2d.device)
    with torch.cuda.device(x2d.device):
        check(lib.lq_colsum(x2d.data_ptr(), rows, cols, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out
