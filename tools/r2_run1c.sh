#!/bin/bash
set -u
O=gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -60
echo "== bench (train, default)"; MCB_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_train.json 2> $O/r2_bench_train.err; echo rc=$?; tail -c 2500 $O/r2_bench_train.json; tail -5 $O/r2_bench_train.err
echo "== bench (infer)"; MCB_BENCH_VERBOSE=1 timeout 500 python bench.py --workload infer --steps 5 --warmup 3 > $O/r2_bench_infer.json 2> $O/r2_bench_infer.err; echo rc=$?; cat $O/r2_bench_infer.json; tail -8 $O/r2_bench_infer.err
