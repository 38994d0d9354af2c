#!/bin/bash
# A/B of the L2 prefetch-ahead (LQ_NO_PREFETCH=1 switches it off): loss-term passes and the main kernels at 64 Mi
show_pen() { python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['case'][6:], d['kind'], round(d['fwd_us'],1), round(d['bwd_us'],1))"; }
show_sweep() { python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if d['elements']==1<<26 and d['scales']=='trained' and d['lambda']==1e-3: print(d['layout_name'], round(d['fwd_us'],1), round(d['bwd_us'],1), round(d['frac_of_measured_peak'],3))"; }
for v in on off; do
  echo "== prefetch $v"
  if [ $v = off ]; then export LQ_NO_PREFETCH=1; else unset LQ_NO_PREFETCH; fi
  PEN_QUICK=1 python tools/penalty_bench.py | show_pen
  python bench.py --sweep --only per_tensor,hwio_like_C512,interleaved_C512,rows_C512 --max-elems 67108864 --steps 20 2>/dev/null | show_sweep
done
