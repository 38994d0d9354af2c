#!/usr/bin/env python
"""End-to-end host-buffer path only (lq_host_fq_fwd_bwd on pinned host arrays): ms per call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import learned_quantization_b200 as lq  # noqa: E402

ops = lq.ops
L = (256, 512, 512)
n = L[0] * L[1] * L[2]
rng = np.random.default_rng(0)
hP = torch.from_numpy(rng.standard_normal(n, dtype=np.float32) * np.float32(0.05)).pin_memory()
hdy = torch.from_numpy(rng.standard_normal(n, dtype=np.float32) * np.float32(1e-3)).pin_memory()
hs = torch.from_numpy(rng.uniform(5e-3, 2e-2, L[1]).astype(np.float32)).pin_memory()
hy = torch.empty(n).pin_memory()
hds = torch.empty(L[1]).pin_memory()
ws = torch.empty(ops.host_workspace_bytes(L), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ops.host_fq_fwd_bwd(hP, hs, hdy, 1e-3, hy, hds, L, ws)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ops.host_fq_fwd_bwd(hP, hs, hdy, 1e-3, hy, hds, L, ws)
dt = (time.perf_counter() - t0) / 10
print(f"{os.environ.get('TAG', '')}: {dt * 1e3:.3f} ms/call  {20 * n / dt / 1e9:.1f} GB/s  (H2D {8 * n / dt / 1e9:.1f} GB/s)")
