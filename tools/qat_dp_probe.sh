#!/bin/bash
# same box: the single-GPU QAT step, then the 2-GPU step with the exchange variants (bench.py --only-qat)
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --only-qat "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$*', 'ms', round(d['ms_per_step'],4), 'exch', d.get('exchange',{}).get('ms'))"; }
python bench.py --only-qat 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=1 ms', round(d['ms_per_step'],4))"
run --qat-no-overlap
run --qat-comm nccl
for b in 8 16 32 64 148; do run --qat-late-blocks $b; done
python bench.py --only-qat 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=1 again ms', round(d['ms_per_step'],4))"
