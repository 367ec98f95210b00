#!/bin/bash
set -u
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 400 python bench.py --workload infer --steps 5 --warmup 3 > $O/r02_bench_infer_c.json 2> $O/r02_bench_infer_c.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/r02_bench_infer_c.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages_ms'])"
