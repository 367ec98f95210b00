#!/bin/bash
# 1-GPU evidence run of round 2 (every step bounded, every stderr kept):
#   tools/gpu.sh 2400 'bash tools/r2_run1.sh > gpurun_out/r2_run1.log 2>&1'
set -u
O=gpurun_out
echo "== conditioned-checkpoint parity numbers"; timeout 300 python tools/measure_config_parity.py conditioned > $O/r2_parity_conditioned.log 2>&1; tail -45 $O/r2_parity_conditioned.log
echo "== bench (train, default)"; MCB_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_train.json 2> $O/r2_bench_train.err; echo rc=$?; tail -c 1500 $O/r2_bench_train.json; tail -5 $O/r2_bench_train.err
echo "== bench (torch_cudnn arm)"; timeout 300 python bench.py --impl torch_cudnn --steps 10 --warmup 3 > $O/r2_bench_cudnn.json 2> $O/r2_bench_cudnn.err; echo rc=$?; cat $O/r2_bench_cudnn.json; tail -3 $O/r2_bench_cudnn.err
echo "== bench (infer)"; MCB_BENCH_VERBOSE=1 timeout 500 python bench.py --workload infer --steps 5 --warmup 3 > $O/r2_bench_infer.json 2> $O/r2_bench_infer.err; echo rc=$?; cat $O/r2_bench_infer.json; tail -8 $O/r2_bench_infer.err
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -60
echo "== ncu post-processing"; timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2_postproc_ncu.csv python tools/ncu_postproc.py > $O/r2_postproc_ncu.out 2>&1; echo rc=$?; tail -3 $O/r2_postproc_ncu.out
