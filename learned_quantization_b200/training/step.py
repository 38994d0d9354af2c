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

    def forward(self) -> None:
        """ONE launch: y_i = floor(P_i / s_i) * s_i for every tensor; layers then hand out y_i."""
        with torch.no_grad():
            ops.mt_fq_fwd([p.data for p in self.params], [s.data for s in self.scales],
                          [y.data for y in self.ys], self.layouts)
        for q, y in zip(self.qlayers, self.ys):
            y.grad = None
            q._banked = y

    def release(self) -> None:
        for q in self.qlayers:
            q._banked = None

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
                 comm: Optional[Comm] = None, ds_mode: str = "exact"):
        self.model, self.loss_fn = model, loss_fn
        self.comm = comm or SingleComm()
        self.ds_mode = ds_mode
        self.bank = QuantBank(model)
        b = self.bank
        quant_ids = {id(p) for p in b.params} | {id(s) for s in b.scales}
        self.plain = [p for p in model.parameters() if id(p) not in quant_ids and p.requires_grad]
        dev = b.params[0].device
        self._peer = isinstance(self.comm, PeerComm)
        self._xws = None
        alloc = None
        if self._peer:
            # the bucket and the partial workspace of the fused exchange live in the peer-mapped arena
            ws_bytes = ops.workspace_bytes([ops.make_desc(layout=l) for l in b.layouts], _lib.LQ_PHASE_BWD)

            def alloc(n_floats):
                nb = -(-n_floats * 4 // 256) * 256
                arena = self.comm.allocate(nb + max(ws_bytes, 256))
                self._xws = arena[nb:]
                return arena[:n_floats * 4].view(torch.float32)
        self.bucket = GradBucket(list(b.params), [p.shape for p in self.plain], [s.shape for s in b.scales], dev,
                                 alloc=alloc)
        if self._peer:
            threshold = b.sgrad_mode == _lib.LQ_SGRAD_THRESHOLD
            self._xdescs = ops.desc_array([
                ops.make_desc(P=p.data, s=s.data, dy=dy, ds=(ds if threshold else None), layout=l)
                for p, s, dy, ds, l in zip(b.params, b.scales, self.bucket.dy, self.bucket.ds, b.layouts)])
            self._xplain = self.bucket.flat[self.bucket.plain_begin:self.bucket.n_exact]
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
        loss.backward()
        with torch.no_grad():
            torch._foreach_copy_(bucket.dy, [y.grad for y in b.ys])
            if self.plain:
                torch._foreach_copy_(bucket.plain, [p.grad for p in self.plain])
            # loss terms that act on P / s directly (custom_loss_terms penalties): identical on every
            # rank, so adding them before the mean-allreduce leaves them unchanged
            extra_p = [(v, p.grad) for v, p in zip(bucket.dy, b.params) if p.grad is not None]
            if extra_p:
                torch._foreach_add_([v for v, _ in extra_p], [g for _, g in extra_p])
        return loss.detach()

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
                # ONE kernel: mean of [dy | plain] over the ranks' arenas + the threshold backward of the mean
                b = self.bank
                self.comm.exchange(self._xdescs, len(self._xdescs), self._xplain, b.lam, b.sgrad_mode, self._xws)
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
