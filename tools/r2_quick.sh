#!/bin/bash
set -u
O=gpurun_out
timeout 300 python -m pytest tests/test_crf_watershed_gpu.py tests/test_instances_gpu.py -q 2>&1 | tail -4
timeout 400 python bench.py --workload infer --steps 5 --warmup 3 --no-cpu-baseline > $O/r02_bench_infer_b.json 2> $O/r02_bench_infer_b.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/r02_bench_infer_b.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['stages_ms'])"
