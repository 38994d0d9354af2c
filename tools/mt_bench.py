#!/usr/bin/env python
"""Model-sized tensors (IMAGENETTE inventory: 20 kernels + 20 biases,
11.17 M elements, L2-resident): per-tensor launches vs the multi-tensor entry points, eager and
as CUDA graphs.  One JSON line.
    python tools/mt_bench.py > gpurun_out/mt_bench.json
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_quantization_b200 as lq  # noqa: E402

ops = lq.ops
KERNELS = [(7, 7, 3, 64)] + [(3, 3, 64, 64)] * 4 + [(3, 3, 64, 128)] + [(3, 3, 128, 128)] * 3 + [(1, 1, 64, 128)] + \
          [(3, 3, 128, 256)] + [(3, 3, 256, 256)] * 3 + [(1, 1, 128, 256)] + [(3, 3, 256, 512)] + [(3, 3, 512, 512)] * 3 + \
          [(1, 1, 256, 512)]


def timeit(fn, reps=20, graph=False):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        run, div = g.replay, reps
    else:
        run, div = (lambda: [fn() for _ in range(reps)]), reps
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / div * 1e3)
    return sorted(ts)[2]


def main():
    torch.cuda.set_device(0)
    out = {}
    ohwi = bool(os.environ.get("MT_OHWI"))        # conv kernels stored O,H,W,I in memory (logical HWIO): the training step's layout
    for orient in ("channelwise", "scalar"):
        Ps, ss = [], []
        for shp in KERNELS:
            if ohwi:
                Ps.append((torch.randn((shp[3], shp[0], shp[1], shp[2]), device="cuda") * 0.05).permute(1, 2, 3, 0))
            else:
                Ps.append(torch.randn(shp, device="cuda") * 0.05)
            ss.append(torch.empty(ops.scale_shape(shp, orient), device="cuda").uniform_(5e-3, 2e-2))
            Ps.append(torch.randn(shp[-1], device="cuda") * 0.05)
            ss.append(torch.empty(1, device="cuda").uniform_(5e-3, 2e-2))
        ys = [torch.empty_like(p) for p in Ps]
        dys = [torch.randn_like(p) * 1e-3 for p in Ps]
        dss = [torch.empty_like(s) for s in ss]
        layouts = [ops.tensor_layout(p, s.shape) for p, s in zip(Ps, ss)]
        n = sum(p.numel() for p in Ps)
        lam = 1e-3

        def fwd_each():
            for p, s, y, l in zip(Ps, ss, ys, layouts):
                ops.fq_fwd(p, s, l, out=y)

        def bwd_each():
            for p, s, dy, l in zip(Ps, ss, dys, layouts):
                ops.fq_bwd_threshold(dy, p, s, lam, l)

        fwd_mt = lambda: ops.mt_fq_fwd(Ps, ss, ys, layouts)
        bwd_mt = lambda: ops.mt_fq_bwd(dys, Ps, ss, dss, lam, layouts)
        table = ops.OptimizerTable(Ps + ss, dys + dss, adam=True, min_value=[None] * len(Ps) + [1.19e-5] * len(ss))
        step_dev = torch.ones((), dtype=torch.int64, device="cuda")
        opt = lambda: ops.opt_adam(table, 1e-4, step_dev)
        r = {"tensors": len(Ps), "elements": n, "storage": "ohwi" if ohwi else "hwio"}
        for name, fn in (("fwd_per_tensor", fwd_each), ("fwd_mt", fwd_mt), ("bwd_per_tensor", bwd_each), ("bwd_mt", bwd_mt),
                         ("adam_fused", opt)):
            r[name + "_eager_us"] = timeit(fn, graph=False)
            r[name + "_graph_us"] = timeit(fn, graph=True)
        r["fwd_mt_GBps"] = 8 * n / (r["fwd_mt_graph_us"] * 1e-6) / 1e9
        r["bwd_mt_GBps"] = 8 * n / (r["bwd_mt_graph_us"] * 1e-6) / 1e9       # dX aliases dy: read dy, P
        r["adam_GBps"] = 28 * n / (r["adam_fused_graph_us"] * 1e-6) / 1e9    # r p,g,m,v  w p,m,v
        out[orient] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
