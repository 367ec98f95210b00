#!/bin/bash
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo rc=$?
python -c "
import json; d=json.loads(open('$O/r02_bench_n1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['phases_ms'], d['roofline']['frac'], d['library_baseline']['ours_over_library'], d['clocks'])"
