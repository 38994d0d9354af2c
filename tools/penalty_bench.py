#!/usr/bin/env python
"""Loss-term reductions (LF:75-116 / 161-195 / 240-275) at microbench size and at model size.
One JSON line per case: forward (4 B/elem: read P) and backward (8 B/elem: read P, write dX) GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_quantization_b200 as lq  # noqa: E402
from learned_quantization_b200 import _lib  # noqa: E402

ops = lq.ops


def timed(fn, reps=10):
    """GPU time per call: a CUDA graph of `reps` calls (the eager call costs ~30 us of Python/ctypes)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    g = torch.Generator(device="cuda").manual_seed(3)
    cases = {"micro_64Mi_channelwise": [((256, 512, 512), (1, 512, 1))],
             "micro_64Mi_per_tensor": [((1 << 26,), (1,))],
             "micro_64Mi_interleaved": [((131072, 512), (1, 512))]}
    import tools.mt_bench as mb  # noqa: E402
    quick = os.environ.get("PEN_QUICK")          # A/B runs: the three 64 Mi layouts, forward passes of MaxBin / Difference
    cases["imagenette_inventory"] = [(k, ops.scale_shape(k, "channelwise")) for k in mb.KERNELS] + \
                                    [((k[-1],), (1,)) for k in mb.KERNELS]
    # the same inventory with the conv kernels stored O,H,W,I in memory (logical HWIO): what the training step holds
    cases["imagenette_inventory_ohwi"] = cases["imagenette_inventory"]
    if quick:
        cases.pop("imagenette_inventory")
        cases.pop("imagenette_inventory_ohwi")
    for name, shapes in cases.items():
        if name.endswith("_ohwi"):
            Ps = [torch.empty((s[3], s[0], s[1], s[2]) if len(s) == 4 else s, device="cuda").normal_(0, 0.05, generator=g)
                  for s, _ in shapes]
            Ps = [p.permute(1, 2, 3, 0) if p.dim() == 4 else p for p in Ps]
        else:
            Ps = [torch.empty(s, device="cuda").normal_(0, 0.05, generator=g) for s, _ in shapes]
        ss = [torch.empty(t, device="cuda").uniform_(5e-3, 2e-2, generator=g) for _, t in shapes]
        n = sum(p.numel() for p in Ps)
        dXs = [torch.empty_like(p) for p in Ps]
        dss = [torch.empty_like(s) for s in ss]
        kinds = (("maxbin", _lib.LQ_PEN_MAXBIN), ("difference", _lib.LQ_PEN_DIFFERENCE), ("inverse", _lib.LQ_PEN_INVERSE))
        for kname, kind in (kinds[:2] if quick else kinds):
            out, ws = ops.penalty_fwd(kind, Ps, ss)
            f_us = timed(lambda: ops.penalty_fwd(kind, Ps, ss))
            b_us = timed(lambda: ops.penalty_bwd(kind, Ps, ss, ws, 1.0, dXs=dXs, dss=dss))
            print(json.dumps({"case": name, "kind": kname, "elements": n, "fwd_us": f_us, "bwd_us": b_us,
                              "fwd_GBps": 4 * n / (f_us * 1e-6) / 1e9, "bwd_GBps": 8 * n / (b_us * 1e-6) / 1e9}), flush=True)


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    main()
