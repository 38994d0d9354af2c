"""Data parallelism for the QAT step: the batch is sharded, parameters are replicated, and once
per step a flat fp32 bucket of gradients is mean-allreduced.

Bucket layout (GradBucket):   [ dy of every quantised parameter | gradients of the plain
parameters (BatchNorm, unquantised Dense) | ds of every learned scale ]

The learned-scale gradient of the nested variant (NQ:62-118) is NOT linear in dy (threshold,
group-wide AND, tanh), so "average the per-rank ds" differs from the reference evaluated on the
global batch.  Two orders of operations are offered:

    exact  allreduce [dy | plain] first; every rank then runs the threshold backward on the
           AVERAGED dy: identical inputs on every rank => bit-identical ds, no ds traffic, and the
           result equals the single-GPU reference on the concatenated batch (up to the
           summation order of dy).  Default.
    local  run the threshold backward on the local dy, then allreduce [dy | plain | ds]:
           what BASELINE.json's wording describes ("learned-scale gradients are allreduced").

`exchange_gradients` is that ordering and nothing else, so that it can be exercised with the
gloo backend on CPU tensors (tests/test_dp_gloo.py); the product path plugs in NcclComm and the
CUDA backward.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence

import torch

from .. import _lib
from .._lib import check, lib


class Comm:
    world: int = 1
    rank: int = 0

    def allreduce_mean_(self, flat: torch.Tensor) -> None:
        raise NotImplementedError


class SingleComm(Comm):
    def allreduce_mean_(self, flat):
        return None


class NcclComm(Comm):
    """lq_dp_* (include/lq.h): our own NCCL communicator over NVLink 5 / NVSwitch, bootstrapped
    by broadcasting the 128-byte unique id through an existing torch.distributed group (any backend)."""

    def __init__(self, rank: Optional[int] = None, world: Optional[int] = None, id_bytes: Optional[bytes] = None):
        import torch.distributed as dist
        if rank is None:
            rank, world = dist.get_rank(), dist.get_world_size()
        self.rank, self.world = int(rank), int(world)
        if id_bytes is None:
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                check(lib.lq_dp_unique_id(buf))
            box = [bytes(buf.raw)]
            dist.broadcast_object_list(box, src=0)
            id_bytes = box[0]
        self._h = C.c_void_p()
        check(lib.lq_dp_init(id_bytes, self.rank, self.world, C.byref(self._h)))

    def allreduce_mean_(self, flat: torch.Tensor) -> None:
        if not flat.is_cuda or flat.dtype != torch.float32 or not flat.is_contiguous():
            raise ValueError("NcclComm.allreduce_mean_ takes a contiguous fp32 CUDA tensor")
        if flat.numel() == 0:
            return
        check(lib.lq_dp_allreduce(self._h, flat.data_ptr(), flat.numel(), 1.0 / self.world,
                                  torch.cuda.current_stream().cuda_stream))

    def allreduce_max_(self, flat: torch.Tensor) -> None:
        check(lib.lq_dp_allreduce_max(self._h, flat.data_ptr(), flat.numel(), torch.cuda.current_stream().cuda_stream))

    def allgather_f64(self, send: torch.Tensor, recv: torch.Tensor) -> None:
        """recv[r] = rank r's `send` (float64, CUDA): the exchange of the sharded threshold backward."""
        if send.dtype != torch.float64 or recv.dtype != torch.float64 or recv.numel() != send.numel() * self.world:
            raise ValueError("allgather_f64: float64 buffers, recv = world * send")
        check(lib.lq_dp_allgather_f64(self._h, send.data_ptr(), recv.data_ptr(), send.numel(),
                                      torch.cuda.current_stream().cuda_stream))

    def close(self):
        if self._h:
            check(lib.lq_dp_destroy(self._h))
            self._h = C.c_void_p()


class P2PFinish:
    """lq_p2p_* (include/lq.h): exchange + finish of the sharded threshold backward fused into one kernel
    over NVLink peer memory.  The 64-byte IPC handles are gathered through torch.distributed."""

    def __init__(self, max_groups: int, rank: Optional[int] = None, world: Optional[int] = None):
        import torch.distributed as dist
        if rank is None:
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.rank, self.world = int(rank), int(world)
        self._h = C.c_void_p()
        mine = C.create_string_buffer(64)
        # every rank reaches every collective below whatever happens locally: a failure on one rank
        # (IPC not permitted in this container, out of memory, ...) raises on ALL ranks instead of hanging
        err = None
        try:
            check(lib.lq_p2p_create(self.rank, self.world, int(max_groups), C.byref(self._h), mine))
        except _lib.LqError as e:
            err = str(e)
        if self.world > 1:
            got = [None] * self.world
            dist.all_gather_object(got, (err, bytes(mine.raw)))
            errs = [e for e, _ in got if e]
            if not errs:
                try:
                    check(lib.lq_p2p_connect(self._h, b"".join(h for _, h in got)))
                except _lib.LqError as e:
                    err = str(e)
                got2 = [None] * self.world
                dist.all_gather_object(got2, err)
                errs = [e for e in got2 if e]
            if errs:
                self.close()
                raise _lib.LqError(_lib.LQ_E_CUDA, "peer-memory exchange unavailable: " + errs[0])
        elif err:
            raise _lib.LqError(_lib.LQ_E_CUDA, err)

    def finish(self, local_stats: torch.Tensor, penalty_threshold: float, n_per_group: float, ds: torch.Tensor) -> torch.Tensor:
        if local_stats.dtype != torch.float64 or not local_stats.is_cuda or not local_stats.is_contiguous():
            raise ValueError("local_stats must be a contiguous float64 CUDA tensor [C, 3]")
        check(lib.lq_fq_bwd_threshold_finish_p2p(self._h, local_stats.data_ptr(), local_stats.shape[0],
                                                 float(penalty_threshold), float(n_per_group), ds.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
        return ds

    def timed_out(self) -> bool:
        flag = C.c_int(0)
        check(lib.lq_p2p_error(self._h, C.byref(flag)))
        return bool(flag.value)

    def close(self):
        if self._h:
            check(lib.lq_p2p_destroy(self._h))
            self._h = C.c_void_p()


class _DeviceSpan:
    """__cuda_array_interface__ over a raw device allocation owned by `owner` (kept alive by the tensor)."""

    def __init__(self, ptr: int, nbytes: int, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerComm(Comm):
    """lq_dpx_* (include/lq.h): the gradient exchange fused with the threshold backward in ONE kernel over
    NVLink peer memory.  The flat gradient bucket and the partial workspace live in an IPC arena that every
    peer maps; allocate() creates it (collective: every rank must call it with the same size) and returns
    it as a uint8 CUDA tensor.  No NCCL call is involved, so a step that uses it is captured as one CUDA graph."""

    def __init__(self, rank: Optional[int] = None, world: Optional[int] = None):
        import torch.distributed as dist
        if rank is None:
            rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
        self.rank, self.world = int(rank), int(world)
        self._h = C.c_void_p()
        self.arena: Optional[torch.Tensor] = None

    def allocate(self, nbytes: int) -> torch.Tensor:
        import torch.distributed as dist
        if self._h:
            raise RuntimeError("PeerComm.allocate may be called once")
        mine = C.create_string_buffer(64)
        base = C.c_void_p()
        err = None
        try:
            check(lib.lq_dpx_create(self.rank, self.world, int(nbytes), C.byref(self._h), mine, C.byref(base)))
        except _lib.LqError as e:
            err = str(e)
        if self.world > 1:          # every rank reaches every collective: a local failure raises everywhere
            got = [None] * self.world
            dist.all_gather_object(got, (err, bytes(mine.raw)))
            errs = [e for e, _ in got if e]
            if not errs:
                try:
                    check(lib.lq_dpx_connect(self._h, b"".join(h for _, h in got)))
                except _lib.LqError as e:
                    err = str(e)
                got2 = [None] * self.world
                dist.all_gather_object(got2, err)
                errs = [e for e in got2 if e]
            if errs:
                self.close()
                raise _lib.LqError(_lib.LQ_E_CUDA, "peer-memory exchange unavailable: " + errs[0])
        elif err:
            raise _lib.LqError(_lib.LQ_E_CUDA, err)
        nbytes = -(-int(nbytes) // 256) * 256
        self.arena = torch.as_tensor(_DeviceSpan(base.value, nbytes, self), device=torch.device("cuda", torch.cuda.current_device()))
        return self.arena

    def exchange(self, descs, n: int, plain: Optional[torch.Tensor], penalty_threshold: float, scale_grad_mode: int,
                 ws: Optional[torch.Tensor], channel: int = 0, max_blocks: int = 0) -> None:
        """mean-allreduce of every desc.dy and of `plain` (in place) + the scale gradients of the averaged dy.
        Exchanges in flight at the same time (different streams) take different channels."""
        check(lib.lq_dpx_exchange(self._h, descs, n,
                                  None if plain is None or plain.numel() == 0 else plain.data_ptr(),
                                  0 if plain is None else plain.numel(), float(penalty_threshold), int(scale_grad_mode),
                                  int(channel), int(max_blocks),
                                  None if ws is None else ws.data_ptr(), 0 if ws is None else ws.numel(),
                                  torch.cuda.current_stream().cuda_stream))

    def allreduce_mean_(self, flat: torch.Tensor, channel: int = 0) -> None:
        """Plain in-place mean all-reduce of a slice of the arena."""
        if flat.numel():
            self.exchange(None, 0, flat, 0.0, _lib.LQ_SGRAD_NONE, None, channel)

    def allreduce_dma_(self, flat: torch.Tensor, channel: int = 0, max_blocks: int = 0) -> None:
        """The same mean all-reduce on the copy engine (lq_dpx_allreduce): bulk asynchronous copies through
        shared memory, a few dozen blocks -- the variant to run beside other kernels."""
        if flat.numel():
            check(lib.lq_dpx_allreduce(self._h, flat.data_ptr(), flat.numel(), int(channel), int(max_blocks),
                                       torch.cuda.current_stream().cuda_stream))

    def timed_out(self) -> bool:
        flag = C.c_int(0)
        check(lib.lq_dpx_error(self._h, C.byref(flag)))
        return bool(flag.value)

    def close(self):
        if self._h:
            self.arena = None
            check(lib.lq_dpx_destroy(self._h))
            self._h = C.c_void_p()


class TorchDistComm(Comm):
    """The same interface over a torch.distributed process group (gloo on CPU in the tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self._group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce_mean_(self, flat):
        if flat.numel() == 0:
            return
        self._dist.all_reduce(flat, op=self._dist.ReduceOp.SUM, group=self._group)
        flat.mul_(1.0 / self.world)


class GradBucket:
    """One flat fp32 buffer with views for every gradient, in the order [dy | plain | ds]."""

    def __init__(self, quant_shapes: Sequence, plain_shapes: Sequence, scale_shapes: Sequence, device, align: int = 4,
                 alloc: Optional[Callable[[int], torch.Tensor]] = None):
        """Each entry is a shape, or a tensor whose memory layout the view must reproduce (a kernel stored
        as OHWI keeps that order inside the bucket: copies into it and the optimizer stay flat)."""
        def carve(items, off):
            spans = []
            for it in items:
                shp = tuple(int(d) for d in (it.shape if isinstance(it, torch.Tensor) else it))
                perm = None
                if isinstance(it, torch.Tensor) and not it.is_contiguous():
                    perm = tuple(sorted(range(it.dim()), key=lambda d: (-it.stride(d), d)))
                    if not it.permute(perm).is_contiguous():
                        raise ValueError("bucket views can only mirror dense tensors")
                n = 1
                for d in shp:
                    n *= d
                spans.append((off, n, shp, perm))
                off += -(-n // align) * align        # keep every view 16-byte aligned for the 128-bit paths
            return spans, off
        q, off = carve(quant_shapes, 0)
        self.plain_begin = off
        p, off = carve(plain_shapes, off)
        self.n_exact = off                            # [dy | plain] is what EXACT mode exchanges
        s, off = carve(scale_shapes, off)
        self.n_total = off
        # alloc(n_floats) supplies the storage instead (PeerComm: a slice of the peer-mapped arena)
        self.flat = torch.zeros(max(off, 1), dtype=torch.float32, device=device) if alloc is None else alloc(max(off, 1))
        if self.flat.dtype != torch.float32 or self.flat.numel() < max(off, 1) or not self.flat.is_contiguous():
            raise ValueError("bucket storage must be a contiguous fp32 tensor of at least n_total elements")
        def view(span):
            off, n, shp, perm = span
            base = self.flat[off:off + n]
            if perm is None:
                return base.view(shp)
            inv = [perm.index(d) for d in range(len(shp))]
            return base.view([shp[d] for d in perm]).permute(inv)
        self.dy = [view(x) for x in q]
        self.plain = [view(x) for x in p]
        self.ds = [view(x) for x in s]

    def exact_part(self) -> torch.Tensor:
        return self.flat[: self.n_exact]

    def all(self) -> torch.Tensor:
        return self.flat[: self.n_total]


def exchange_gradients(bucket: GradBucket, comm: Comm, scale_backward: Callable[[], None], mode: str = "exact") -> None:
    """Order the scale backward and the collective as `mode` says.  `scale_backward()` must read
    bucket.dy and write bucket.ds (the threshold rule, or zeros for the STE variants)."""
    if mode == "exact":
        comm.allreduce_mean_(bucket.exact_part())
        scale_backward()
    elif mode == "local":
        scale_backward()
        comm.allreduce_mean_(bucket.all())
    else:
        raise ValueError(f"mode must be 'exact' or 'local', got {mode!r}")
