#!/usr/bin/env python
"""How much of the bench step is launch gaps: the same fwd+bwd pair timed (a) eagerly with per-kernel
events, (b) eagerly with two events around K steps, (c) as one CUDA graph of K steps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_quantization_b200 as lq  # noqa: E402

ops = lq.ops
L = (256, 512, 512)
n = L[0] * L[1] * L[2]
g = torch.Generator(device="cuda").manual_seed(1)
sets = [(torch.empty(n, device="cuda").normal_(0, 0.05, generator=g), torch.empty(n, device="cuda").normal_(0, 1e-3, generator=g)) for _ in range(2)]
s = torch.empty(L[1], device="cuda").uniform_(5e-3, 2e-2, generator=g)
y = torch.empty(n, device="cuda")
K = 30


def step(i):
    P, dy = sets[i % 2]
    ops.fq_fwd(P, s, L, out=y)
    ops.fq_bwd_threshold(dy, P, s, 1e-3, L, write_dx=True)


for i in range(5):
    step(i)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(K):
    step(i)
b.record()
torch.cuda.synchronize()
print("eager, 2 events around K steps: %.1f us/step" % (a.elapsed_time(b) / K * 1e3))
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for i in range(K):
        step(i)
gr.replay()
torch.cuda.synchronize()
a.record(); gr.replay(); b.record()
torch.cuda.synchronize()
print("one CUDA graph of K steps:      %.1f us/step" % (a.elapsed_time(b) / K * 1e3))
