#!/usr/bin/env python
"""One-off exhaustive proof run (NOT collected by pytest; ~1-2 GPU-minutes):

    python tests/exhaustive_divcheck.py > profiles/r01_divcheck_exhaustive.txt

Checks div_fast (the hoisted-reciprocal division of the kernels) against __fdiv_rn for ALL
2^23 x 2^23 (divisor mantissa, dividend mantissa) pairs, at dividend exponents 0 and -1
relative to the divisor (quotients in [1,2) and [0.5,1): both rounding positions).  FMA
arithmetic and MUFU.RCP are invariant under power-of-two scaling inside the guard window
(no intermediate can leave the normal range there), so mantissa space is the whole space;
the window edges themselves are sampled by tests/test_gpu_parity.py.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learned_quantization_b200._lib import lib, check  # noqa: E402


def main():
    chunk = 1 << 20
    out = torch.zeros(3, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    t0 = time.time()
    for a_hi in (127 << 23, 126 << 23):
        for m0 in range(0, 1 << 23, chunk):
            s_bits = (np.arange(m0, m0 + chunk, dtype=np.uint32) | np.uint32(127 << 23)).view(np.int32)
            sb = torch.from_numpy(s_bits).cuda()
            check(lib.lq_debug_divcheck(sb.data_ptr(), sb.numel(), a_hi, out.data_ptr(), stream))
        torch.cuda.synchronize()
        bad, seen, first = (int(v) for v in out.cpu())
        print(f"a_exp_bits={a_hi:#x}: cumulative pairs checked={seen} mismatches={bad} first={first:#x} "
              f"elapsed={time.time() - t0:.1f}s", flush=True)
    bad, seen, first = (int(v) for v in out.cpu())
    print("RESULT", "PASS" if bad == 0 and seen == 2 * (1 << 46) else "FAIL", "pairs", seen, "mismatches", bad)
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
