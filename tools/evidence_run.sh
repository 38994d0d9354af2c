#!/bin/bash
# One GPU call that refreshes the round's evidence (run through gpurun from the repo root):
#   tools/evidence_run.sh r02f
# writes gpurun_out/<tag>_*: pytest log, bench line, ncu full captures (ROW headline kernels, COL kernels), launch list,
# layout sweep, loss-term passes, model-sized tensors.
tag=${1:-r02f}
o=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $o/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $o/${tag}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-qat > $o/${tag}_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:row_kernel -c 6 -f -o $o/prof_${tag} python bench.py --steps 5 --warmup 3 --no-qat > $o/${tag}_ncu.log 2>&1; echo "ncu row rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-qat > /dev/null 2>&1; echo "launch list rc=$?"
timeout 120 python tools/one_layout.py 131072,512,1 5 > $o/${tag}_col_plain.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:col_ -c 4 -f -o $o/prof_${tag}_col python tools/one_layout.py 131072,512,1 2 > $o/${tag}_col_ncu.log 2>&1; echo "ncu col rc=$?"
timeout 600 python bench.py --sweep --max-elems 268435456 --steps 20 > $o/${tag}_sweep.jsonl 2> $o/${tag}_sweep.err; echo "sweep rc=$?"
timeout 200 python tools/penalty_bench.py > $o/${tag}_penalty.jsonl 2>&1; echo "penalty rc=$?"
timeout 200 python tools/mt_bench.py > $o/${tag}_mt_bench.json 2>/dev/null; echo "mt rc=$?"
for s in 131072,512 256,512,512 67108864; do
  timeout 200 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum --clock-control none --csv --log-file $o/${tag}_pen_launches_${s//,/x}.csv python tools/pen_probe.py $s > /dev/null 2>&1
done; echo "pen launch lists done"
