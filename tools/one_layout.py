#!/usr/bin/env python
"""A handful of fwd/bwd launches on ONE layout (for ncu captures of a specific kernel).
    python tools/one_layout.py 131072,512,1 [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_quantization_b200 as lq  # noqa: E402

ops = lq.ops
layout = tuple(int(v) for v in sys.argv[1].split(","))
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = layout[0] * layout[1] * layout[2]
g = torch.Generator(device="cuda").manual_seed(1)
P = torch.empty(n, device="cuda").normal_(0, 0.05, generator=g)
dy = torch.empty(n, device="cuda").normal_(0, 1e-3, generator=g)
s = torch.empty(layout[1], device="cuda").uniform_(5e-3, 2e-2, generator=g)
y = torch.empty_like(P)
for _ in range(reps):
    ops.fq_fwd(P, s, layout, out=y)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    ops.fq_bwd_threshold(dy, P, s, 1e-3, layout, write_dx=True)
    ev[i + 1].record()
torch.cuda.synchronize()
print("bwd us:", [round(ev[i].elapsed_time(ev[i + 1]) * 1e3, 1) for i in range(reps)])
