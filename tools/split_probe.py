"""Which kernel of the threshold backward pays for short rows at large sizes?  A few launches per layout, meant to
run under `ncu --metrics gpu__time_duration.sum` (per-kernel durations: element pass vs second stage)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_quantization_b200 as lq

ops = lq.ops
for layout in [(4096, 512, 512), (1, 512, 2097152), (1024, 512, 512), (1, 512, 524288), (2097152, 512, 1)]:
    n = layout[0] * layout[1] * layout[2]
    P = torch.empty(n, device="cuda").normal_(0, 0.05)
    dy = torch.empty(n, device="cuda").normal_(0, 1e-3)
    s = torch.empty(layout[1], device="cuda").uniform_(5e-3, 2e-2)
    for _ in range(3):
        ops.fq_bwd_threshold(dy, P, s, 1e-3, layout, write_dx=True)
    torch.cuda.synchronize()
    print(layout, "ok", flush=True)
    del P, dy
