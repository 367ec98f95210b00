#!/bin/bash
set -u
O=gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload infer --steps 5 --warmup 3 > $O/r02_bench_infer_d.json 2> $O/r02_bench_infer_d.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/r02_bench_infer_d.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages_ms'])"
