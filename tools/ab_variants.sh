#!/bin/bash
# A/B of library variants built by tools/build_variant.sh on the interleaved (COL) layouts:
#   tools/ab_variants.sh "layouts" variant1 variant2 ...      ("base" = the in-tree build)
layouts=$1; shift
show() { python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l)
    if d['elements']>=1<<24 and d['scales']=='trained' and d['lambda']==1e-3: print('  ', d['layout_name'], d['elements']>>20, 'Mi fwd', round(d['fwd_us'],1), 'bwd', round(d['bwd_us'],1), 'frac', round(d['frac_of_measured_peak'],3))"; }
for v in "$@"; do
  echo "== $v"
  if [ $v = base ]; then unset LQ_LIBRARY; else export LQ_LIBRARY=$PWD/build/variants/liblq_$v.so; fi
  python bench.py --sweep --only $layouts --max-elems ${MAXE:-67108864} --steps 20 2>/dev/null | show
done
