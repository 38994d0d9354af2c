"""Loss-term passes on one 64 Mi tensor, a few launches each: meant for ncu.
    python tools/pen_probe.py 131072,512      interleaved (COL)
    python tools/pen_probe.py 256,512,512     channelwise on an HWIO-like tensor (ROW)
    python tools/pen_probe.py 67108864        per-tensor"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import learned_quantization_b200 as lq
from learned_quantization_b200 import _lib
ops = lq.ops
shape = tuple(int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "131072,512").split(","))
sshape = (1,) if len(shape) == 1 else ((1, shape[1]) if len(shape) == 2 else (1, shape[1], 1))
P = [torch.empty(shape, device="cuda").normal_(0, 0.05)]
s = [torch.empty(sshape, device="cuda").uniform_(5e-3, 2e-2)]
dX = [torch.empty_like(P[0])]; ds = [torch.empty_like(s[0])]
for kind in (_lib.LQ_PEN_MAXBIN, _lib.LQ_PEN_DIFFERENCE):
    for _ in range(2):
        out, ws = ops.penalty_fwd(kind, P, s)
        ops.penalty_bwd(kind, P, s, ws, 1.0, dXs=dX, dss=ds)
torch.cuda.synchronize()
print("ok")
