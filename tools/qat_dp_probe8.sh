#!/bin/bash
N=${1:-8}
show() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1', 'ms', round(d['ms_per_step'],4), 'exch', d.get('exchange',{}).get('ms'))"; }
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --only-qat --qat-steps 60 "$@" 2>/dev/null | show "N=$N $*"; }
one() { python bench.py --only-qat --qat-steps 60 2>/dev/null | show "N=1"; }
one
run --qat-skip-exchange
run
run --qat-no-overlap
run --qat-comm nccl
run --qat-late-blocks 16
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dpx_probe.py 2>/dev/null | grep floats
