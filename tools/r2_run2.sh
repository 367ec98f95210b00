#!/bin/bash
# 2-GPU evidence run (every step bounded, all output kept):
#   tools/gpu.sh --gpus 2 900 'bash tools/r2_run2.sh > gpurun_out/r2_run2.log 2>&1'
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout -k 5 $t "$@" > $O/r2_n2_$name.out 2> $O/r2_n2_$name.err; echo "rc=$?"; tail -c 600 $O/r2_n2_$name.out; echo; tail -4 $O/r2_n2_$name.err; }
run base 100 $TR --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown
run ov2 100 env MCB_OVERLAP_ALLREDUCE=2 $TR --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown
run tests 260 env MCB_TEST_SYNC_BN=1 python -m pytest tests/test_multi_gpu_gpu.py -x -q -s
run syncbn 100 env MCB_SYNC_BN=1 $TR --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown
