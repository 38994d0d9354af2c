"""The QAT training step around the hot path, one launch per phase.

Per step, for a model with T quantised tensors (T = 40 for the IMAGENETTE ResNet, 22 for
paper_implementation, 4 for MNIST -- SURVEY Appendix C) the per-layer path of the drop-in layers
launches T forward kernels and 2T backward kernels, plus Keras-style per-variable optimizer and
constraint ops.  Here:

    QuantBank.forward()      1 launch   lq_mt_fq_fwd over all (P, s)         NQ:55-60
    model forward / loss / autograd backward  (cuDNN / cuBLAS through torch; not ours)
    gather dy + plain grads  torch._foreach_copy_ into the flat GradBucket
    exchange_gradients()     1 collective (lq_dp_allreduce) + QuantBank.backward():
                             2 launches lq_mt_fq_bwd (+ second stage)        NQ:62-118
    optimizer                1-3 launches lq_opt_adam / lq_opt_sgd with l2, 1/world scaling and
                             MinValueConstraint folded in                    NQ:35-46, experiment.py:763-771
"""
from __future__ import annotations

import os

from typing import Callable, List, Optional

import torch
from torch import nn

from .. import _lib, ops
from ..custom_components import _impl
from .dp import Comm, GradBucket, PeerComm, SingleComm, exchange_gradients


def quantised_pairs(model: nn.Module):
    """[(host layer, attribute name of the parameter, parameter, its CustomQuantizedScaleLayer)] in module order."""
    out = []
    for m in model.modules():
        if isinstance(m, _impl.DenseLayerBase) and m.built:
            out.append((m, "W", m.W, m.nested_q_w_layer))
            out.append((m, "b", m.b, m.nested_q_b_layer))
        elif isinstance(m, _impl.Conv2DLayerBase) and m.built:
            out.append((m, "kernel", m.kernel, m.nested_q_k_layer))
            if m._use_bias:
                out.append((m, "b", m.b, m.nested_q_b_layer))
    return out


class QuantBank:
    """Every quantised tensor of a (built) model, driven through the multi-tensor entry points."""

    def __init__(self, model: nn.Module):
        self.pairs = quantised_pairs(model)
        if not self.pairs:
            raise ValueError("model has no built quantised layers (run one forward first)")
        self.hosts = [h for h, _, _, _ in self.pairs]
        self.params = [p for _, _, p, _ in self.pairs]
        self.qlayers = [q for _, _, _, q in self.pairs]
        self.scales = [q.scale for q in self.qlayers]
        self.layouts = [ops.tensor_layout(p, s.shape) for p, s in zip(self.params, self.scales)]
        modes = {q._sgrad_mode for q in self.qlayers}
        lams = {float(q._threshold() or 0.0) for q in self.qlayers}
        if len(modes) != 1 or len(lams) != 1:
            raise ValueError("QuantBank needs one scale-gradient rule and one penalty threshold per model")
        self.sgrad_mode, self.lam = modes.pop(), lams.pop()
        dev = self.params[0].device
        total = sum(-(-p.numel() // 4) * 4 for p in self.params)
        self._yflat = torch.empty(total, dtype=torch.float32, device=dev)
        self.ys, off = [], 0
        for p in self.params:                       # y_i mirrors P_i's memory layout inside the flat buffer
            perm = ops.mem_perm(p)
            if perm is None:
                raise ValueError("quantised parameters must be dense in memory")
            inv = [perm.index(d) for d in range(p.dim())]
            y = self._yflat[off:off + p.numel()].view([p.shape[d] for d in perm]).permute(inv)
            self.ys.append(y.requires_grad_(True))
            off += -(-p.numel() // 4) * 4

    def fuse_loss(self, loss) -> bool:
        """Let the forward produce the MaxBin / Difference term of `loss` (an SCCE* object of custom_loss_functions) as a
        by-product of the same pass (lq_mt_fq_fwd_penalty, SURVEY 8b / 8d: "penalty forward fused into K1").  Possible when
        every tensor the loss walks (LF:85-88: kernel and bias of each layer, with their scales) is one of the bank's."""
        from ..custom_loss_terms.custom_components import custom_loss_functions as lf
        if getattr(loss, "_kind", None) not in (_lib.LQ_PEN_MAXBIN, _lib.LQ_PEN_DIFFERENCE):
            return False
        tensors = lf._layer_tensors(loss.layers)
        where = {id(p): i for i, p in enumerate(self.params)}
        idx = []
        for p, s in zip(tensors[0::2], tensors[1::2]):
            i = where.get(id(p))
            if i is None or self.scales[i] is not s:
                return False
            idx.append(i)
        if len(set(idx)) != len(idx):
            return False
        rest = [i for i in range(len(self.params)) if i not in set(idx)]
        self._fused = {"loss": loss, "idx": idx, "rest": rest, "ws": None,
                       "out": torch.empty((), dtype=torch.float32, device=self.params[0].device)}
        return True

    _fused = None

    def forward(self) -> None:
        """ONE launch: y_i = floor(P_i / s_i) * s_i for every tensor; layers then hand out y_i."""
        with torch.no_grad():
            f = self._fused
            if f is None:
                ops.mt_fq_fwd([p.data for p in self.params], [s.data for s in self.scales],
                              [y.data for y in self.ys], self.layouts)
            else:
                # the loss's tensors, in the loss's order: forward + loss term in one pass; whatever else is quantised after
                pick = lambda idx: ([self.params[i].data for i in idx], [self.scales[i].data for i in idx],
                                    [self.ys[i].data for i in idx], [self.layouts[i] for i in idx])
                Ps, ss, ys, layouts = pick(f["idx"])
                out, f["ws"] = ops.mt_fq_fwd_penalty(f["loss"]._kind, Ps, ss, ys, layouts, ws=f["ws"], out=f["out"])
                if f["rest"]:
                    ops.mt_fq_fwd(*pick(f["rest"]))
                f["loss"]._fused = {"out": out, "ws": f["ws"], "Ps": Ps, "ss": ss, "layouts": layouts}
        for q, y in zip(self.qlayers, self.ys):
            y.grad = None
            q._banked = y

    def release(self) -> None:
        for q in self.qlayers:
            q._banked = None
        if self._fused is not None:
            self._fused["loss"]._fused = None      # the fused term, too, is valid for this step's parameters only

    def backward(self, dys: List[torch.Tensor], dss: List[torch.Tensor]) -> None:
        """dss_i <- scale gradient from dys_i (the parameter gradient is dys_i itself, NQ:118)."""
        with torch.no_grad():
            ops.mt_fq_bwd(dys, [p.data for p in self.params], [s.data for s in self.scales], dss,
                          self.lam, self.layouts, scale_grad_mode=self.sgrad_mode)


class TrainStep:
    """forward -> loss -> backward -> gradient exchange -> fused optimizer, for a built model.

    loss_fn(y_true, y_pred) returns per-sample losses (Keras convention; the step takes the mean).
    optimizer: "adam" (Keras Adam: lr, 0.9, 0.999, 1e-7) or "sgd" (momentum / nesterov / weight_decay).
    """

    def __init__(self, model: nn.Module, loss_fn: Callable, *, optimizer: str = "adam", lr: float = 1e-4,
                 momentum: float = 0.0, nesterov: bool = False, weight_decay: float = 0.0,
                 comm: Optional[Comm] = None, ds_mode: str = "exact", overlap: bool = True, late_blocks: int = 16,
                 fuse_loss: bool = True):
        """overlap (PeerComm, exact mode): start exchanging the late layers' gradients on a side stream as soon
        as the backward has produced them, while it is still working on the early layers.  late_blocks: grid of
        that early-started exchange -- it waits on NVLink, not on arithmetic, so it is given few SMs and a long
        time (hidden behind the rest of the backward) instead of many SMs taken away from the convolutions."""
        self.late_blocks = int(late_blocks)
        self.model, self.loss_fn = model, loss_fn
        self.comm = comm or SingleComm()
        self.ds_mode = ds_mode
        self.bank = QuantBank(model)
        b = self.bank
        # a custom-loss object (SCCEMaxBin / SCCEDifference bound method): its term comes out of the forward pass itself
        self._loss_obj = getattr(loss_fn, "__self__", None)
        self.fused_loss = bool(fuse_loss and not os.environ.get("LQ_NO_FUSE_LOSS")      # (A/B switch for bench runs)
                               and self._loss_obj is not None and hasattr(self._loss_obj, "_kind")
                               and b.fuse_loss(self._loss_obj))
        quant_ids = {id(p) for p in b.params} | {id(s) for s in b.scales}
        self.plain = [p for p in model.parameters() if id(p) not in quant_ids and p.requires_grad]
        dev = b.params[0].device
        self._peer = isinstance(self.comm, PeerComm)
        alloc = None
        if self._peer:
            # The exchange runs in (at most) two pieces.  LATE = the tensors at the end of the model holding ~90 %
            # of the elements: the backward produces their gradients first, so their exchange can start while the
            # early layers' backward is still running.  EARLY = the rest, plus the plain gradients.
            n = len(b.params)
            sizes = [p.numel() for p in b.params]
            k = n
            while k > 0 and sum(sizes[k:]) < 0.9 * sum(sizes):
                k -= 1
            if not (overlap and ds_mode == "exact" and self.comm.world > 1 and 0 < k < n):
                k = 0                                  # one piece
            self._late, self._early = list(range(k, n)), list(range(0, k))
            ws_of = lambda idx: -(-max(ops.workspace_bytes([ops.make_desc(layout=b.layouts[i]) for i in idx],
                                                          _lib.LQ_PHASE_BWD), 256) // 256) * 256
            ws_late, ws_early = ws_of(self._late), ws_of(self._early)

            def alloc(n_floats):
                # the bucket and the partial workspaces of the fused exchange live in the peer-mapped arena
                nb = -(-n_floats * 4 // 256) * 256
                arena = self.comm.allocate(nb + ws_late + ws_early)
                self._xws = {"late": arena[nb:nb + ws_late], "early": arena[nb + ws_late:]}
                return arena[:n_floats * 4].view(torch.float32)
        self.bucket = GradBucket(list(b.params), [p.shape for p in self.plain], [s.shape for s in b.scales], dev,
                                 alloc=alloc)
        if self._peer:
            threshold = b.sgrad_mode == _lib.LQ_SGRAD_THRESHOLD
            descs = lambda idx: ops.desc_array([
                ops.make_desc(P=b.params[i].data, s=b.scales[i].data, dy=self.bucket.dy[i],
                              ds=(self.bucket.ds[i] if threshold else None), layout=b.layouts[i]) for i in idx])
            self._xdescs = {"late": descs(self._late), "early": descs(self._early)}
            self._xplain = self.bucket.flat[self.bucket.plain_begin:self.bucket.n_exact]
            self._side = torch.cuda.Stream(priority=-1)     # its few blocks are placed ahead of the backward's
            lo = self.bucket.dy[self._late[0]].data_ptr() - self.bucket.flat.data_ptr()
            self._late_range = self.bucket.flat[lo // 4:self.bucket.plain_begin]
            self._armed = False          # hooks act only inside _forward_backward
            self._late_done = False      # the LATE piece has been launched on the side stream in this step
            self._pending = 0
            self._direct_grads = None    # does the loss act on P / s directly (loss-term penalties)? known after step 1
            if self._early:
                for i in self._late:
                    b.ys[i].register_post_accumulate_grad_hook(self._on_late_grad)
        # optimizer table over [P | plain | s], gradients = the bucket views (pointers never change)
        l2 = []
        for host in b.hosts:
            reg = getattr(host, "regularizer", None)
            l2.append(float(getattr(reg, "l2", 0.0)) if reg is not None else 0.0)
        params = [p.data for p in b.params] + [p.data for p in self.plain] + [s.data for s in b.scales]
        grads = self.bucket.dy + self.bucket.plain + self.bucket.ds
        l2s = l2 + [0.0] * len(self.plain) + [0.0] * len(b.scales)
        mins = [None] * (len(b.params) + len(self.plain)) + [float(q.constraint.min_value) for q in b.qlayers]
        self.optimizer, self.lr = optimizer, lr
        self.momentum, self.nesterov = momentum, nesterov
        self.table = ops.OptimizerTable(params, grads, adam=(optimizer == "adam"), momentum=(momentum != 0.0),
                                        l2=l2s, min_value=mins, weight_decay=weight_decay)
        self.iterations = 0
        self._step_dev = torch.zeros((), dtype=torch.int64, device=dev)   # Adam's t, read by the kernel
        self._graph = None

    def _scale_backward(self):
        if self.bank.sgrad_mode == _lib.LQ_SGRAD_NONE:
            # CL:61-62 / PI:62-64: the operator's scale gradient is zeros -- one fill of the bucket's tail
            # instead of one per tensor
            self.bucket.flat[self.bucket.n_exact:self.bucket.n_total].zero_()
            return
        self.bank.backward(self.bucket.dy, self.bucket.ds)

    # The step in three pieces, so that the collective can stay outside a CUDA graph:
    def _forward_backward(self, x, y_true):
        """forward, loss, autograd backward, gradients gathered into the flat bucket."""
        b, bucket = self.bank, self.bucket
        b.forward()
        for p in self.plain:
            p.grad = None
        for t in b.params + b.scales:
            t.grad = None
        loss = self.loss_fn(y_true, self.model(x)).mean()
        split = self._peer and bool(self._early) and self._overlap_ok()
        if split:
            self._main = torch.cuda.current_stream()
            self._pending, self._late_done, self._armed = len(self._late), False, True
        try:
            loss.backward()
        finally:
            if self._peer:
                self._armed = False
        with torch.no_grad():
            rest = self._early if (split and self._late_done) else range(len(b.ys))
            torch._foreach_copy_([bucket.dy[i] for i in rest], [b.ys[i].grad for i in rest])
            if self.plain:
                torch._foreach_copy_(bucket.plain, [p.grad for p in self.plain])
            # loss terms that act on P / s directly (custom_loss_terms penalties) are identical on every rank:
            # adding them before or after the mean over ranks is the same thing.  They go in before it, except
            # when the LATE piece is already being exchanged (then _exchange adds them after the join).
            self._extra_p = [(v, p.grad) for v, p in zip(bucket.dy, b.params) if p.grad is not None]
            if self._peer and self._direct_grads is None:
                self._direct_grads = bool(self._extra_p) or any(s.grad is not None for s in b.scales)
            if self._extra_p and not (split and self._late_done):
                torch._foreach_add_([v for v, _ in self._extra_p], [g for _, g in self._extra_p])
                self._extra_p = []
        return loss.detach()

    # ------------------------------------------------------------------ overlapped exchange (PeerComm)
    def _overlap_ok(self) -> bool:
        """The early launch is taken from the second step on, and never when the threshold rule must see
        gradients that the loss adds to P directly (they are only known after the backward)."""
        if self._direct_grads is None:
            return False
        return not (self._direct_grads and self.bank.sgrad_mode == _lib.LQ_SGRAD_THRESHOLD)

    def _on_late_grad(self, _y) -> None:
        """post-accumulate hook of the LATE tensors (runs on the autograd thread during loss.backward()):
        when the last of them has its gradient, fork the side stream and exchange the piece there."""
        if not self._armed:
            return
        self._pending -= 1
        if self._pending != 0:
            return
        b, bucket = self.bank, self.bucket
        ev = torch.cuda.Event()
        ev.record(self._main)                        # everything the backward has enqueued so far
        self._side.wait_event(ev)
        with torch.no_grad(), torch.cuda.stream(self._side):
            torch._foreach_copy_([bucket.dy[i] for i in self._late], [b.ys[i].grad for i in self._late])
            # beside the convolutions: the copy-engine all-reduce (few blocks, bytes in flight in shared memory);
            # the scale backward of these tensors follows on the main stream after the join (_exchange)
            self.comm.allreduce_dma_(self._late_range, channel=1, max_blocks=self.late_blocks)
        self._late_done = True

    def _exchange_piece(self, which: str) -> None:
        b = self.bank
        d = self._xdescs[which]
        early = which == "early" or not self._early       # the piece that carries the plain gradients
        self.comm.exchange(d if len(d) else None, len(d), self._xplain if early else None, b.lam, b.sgrad_mode,
                           self._xws[which], channel=(0 if early else 1))

    def _update(self):
        """penalty scale gradients, then the fused optimizer + constraint."""
        b, bucket = self.bank, self.bucket
        with torch.no_grad():
            extra_s = [(v, s.grad) for v, s in zip(bucket.ds, b.scales) if s.grad is not None]
            if extra_s:
                torch._foreach_add_([v for v, _ in extra_s], [g.view(v.shape) for v, g in extra_s])
            self._step_dev.add_(1)
            if self.optimizer == "adam":
                ops.opt_adam(self.table, self.lr, self._step_dev)
            else:
                ops.opt_sgd(self.table, self.lr, self.momentum, self.nesterov)

    def _exchange(self, scale_backward=None):
        """gradient exchange + scale backward, ordered as ds_mode says (training/dp.py)."""
        scale_backward = scale_backward or self._scale_backward
        with torch.no_grad():
            if not self._peer:
                exchange_gradients(self.bucket, self.comm, scale_backward, self.ds_mode)
            elif self.ds_mode == "exact":
                # ONE kernel per piece: mean of [dy | plain] over the ranks' arenas + the threshold backward of the mean
                b = self.bank
                if self._early and not self._late_done:
                    self._exchange_piece("late")          # not launched early (first step, or called on its own)
                self._exchange_piece("early" if self._early else "late")
                if self._late_done:
                    torch.cuda.current_stream().wait_stream(self._side)
                    self._late_done = False
                    if b.sgrad_mode == _lib.LQ_SGRAD_THRESHOLD:      # NQ:62-118 on the averaged gradients of the late tensors
                        L = self._late
                        ops.mt_fq_bwd([self.bucket.dy[i] for i in L], [b.params[i].data for i in L],
                                      [b.scales[i].data for i in L], [self.bucket.ds[i] for i in L], b.lam,
                                      [b.layouts[i] for i in L], scale_grad_mode=b.sgrad_mode)
                if getattr(self, "_extra_p", None):
                    torch._foreach_add_([v for v, _ in self._extra_p], [g for _, g in self._extra_p])
                    self._extra_p = []
                if b.sgrad_mode == _lib.LQ_SGRAD_NONE:
                    self.bucket.flat[self.bucket.n_exact:self.bucket.n_total].zero_()
            elif self.ds_mode == "local":
                scale_backward()
                self.comm.allreduce_mean_(self.bucket.all())
            else:
                raise ValueError(f"mode must be 'exact' or 'local', got {self.ds_mode!r}")

    def __call__(self, x: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        try:
            loss = self._forward_backward(x, y_true)
        finally:
            # the banked y are valid for THIS step's parameters only: a later model(x) (validation, callbacks)
            # must quantise the updated parameters through the per-layer path
            self.bank.release()
        self._exchange()
        self._update()
        self.iterations += 1
        return loss

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, x: torch.Tensor, y_true: torch.Tensor, warmup: int = 3) -> None:
        """Capture the step in CUDA graphs: it issues several hundred small launches (cuDNN, BatchNorm,
        our multi-tensor kernels) and is otherwise bounded by the host's launch rate.  Everything the
        step changes between iterations lives in device memory (Adam's t included), so a replay is
        exact.  Single device: ONE graph for the whole step.  Data parallel: three graphs
        (forward+backward | scale backward | update) with the NCCL collective launched between them
        from the host -- capturing the collective itself deadlocked at replay on this stack
        (NCCL 2.28.9, 2 x B200; global and thread-local capture modes alike), so it stays outside.
        `x`, `y_true` fix the static input shapes; replay() copies new batches into them."""
        self._sx, self._sy = x.clone(), y_true.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                  # cuDNN autotuning and lazy allocations happen here
                self(self._sx, self._sy)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.comm.world == 1 or self._peer:      # no library collective inside: the whole step is one graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._sloss = self(self._sx, self._sy)
            self.iterations -= 1                     # recording does not execute
            self._graph = (g,)
            return
        ga, gs, gu = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(ga):
            self._sloss = self._forward_backward(self._sx, self._sy)
        self.bank.release()
        pool = ga.pool()
        with torch.cuda.graph(gs, pool=pool):
            self._scale_backward()
        with torch.cuda.graph(gu, pool=pool):
            self._update()
        self._graph = (ga, gs, gu)

    def replay(self, x: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
        """One captured step on a new batch (device or pinned-host tensors of the captured shapes)."""
        if self._graph is None:
            raise RuntimeError("call capture() first")
        self._sx.copy_(x, non_blocking=True)
        self._sy.copy_(y_true, non_blocking=True)
        return self._graph_replay_only()

    def _graph_replay_only(self) -> torch.Tensor:
        """Replay on whatever the static input buffers hold."""
        if len(self._graph) == 1:
            self._graph[0].replay()
        else:
            ga, gs, gu = self._graph
            ga.replay()
            self._exchange(gs.replay)
            gu.replay()
        self.iterations += 1
        return self._sloss
