#!/bin/bash
# env-only A/B of the library's tuning switches on the bench workload (GPU box): bash tools/envsweep.sh VAR=val ...
run() { env $1 timeout 200 python bench.py --steps 20 --warmup 3 --no-breakdown --no-cpu-baseline --no-library-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-40s %.3f ms/step  e2e %.1f' % ('$1', d['ms_per_step'], d['e2e']['value']))"; }
run "MCB_BASELINE=1"
for v in "$@"; do run "$v"; done
