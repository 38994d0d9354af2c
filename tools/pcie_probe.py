#!/usr/bin/env python
"""Pinned-memory PCIe bandwidth of the box: H2D alone, D2H alone, both directions at once."""
import torch

n = 64 << 20
h = torch.empty(n, dtype=torch.float32).pin_memory()
h2 = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
d2 = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


t = timed(lambda: d.copy_(h, non_blocking=True)); print(f"H2D alone : {4 * n / t / 1e9:.1f} GB/s")
t = timed(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H alone : {4 * n / t / 1e9:.1f} GB/s")


def both():
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


t = timed(both); print(f"both      : {4 * n / t / 1e9:.1f} GB/s each direction")
