#!/bin/bash
# Our kernels inside one eager QAT step, warm caches (ncu launch list): tools/qat_kernels.sh qat|qat_custom_loss [tag]
c=${1:-qat}; tag=${2:-qk}
re='regex:^(col_kernel|row_kernel|mt_kernel|mt_finalize_kernel|mt_ste_kernel|opt_kernel|colsum_.*|stat_.*|pen_.*|upstream_kernel|col_tma_kernel|finalize_warp_kernel|fold_partials_kernel|dpx_.*)$'
ncu --metrics gpu__time_duration.sum -k "$re" --clock-control none --cache-control none --csv --log-file gpurun_out/${tag}_$c.csv python bench.py --only-qat --qat-config $c --qat-eager --qat-steps 2 --qat-warmup 3 > /dev/null 2>&1
python - <<EOF
import csv
rows=[r for r in csv.reader(open("gpurun_out/${tag}_$c.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
seq=[(r[ki][:64], r[gi], float(r[vi])/1e3) for r in rows[1:]]
per=len(seq)//5 if len(seq)%5==0 else None
print("== $c: launches", len(seq), "per step", per)
last=seq[-per:] if per else seq[-16:]
for s in last: print("   %-66s %-16s %7.1f us" % s)
print("   sum of our kernels in the step: %.1f us" % sum(s[2] for s in last))
EOF
