#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout -k 5 $t "$@" > $O/r2_n2_$name.out 2> $O/r2_n2_$name.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r2_n2_$name.out").read().strip().splitlines()[-1])
    print("   %.3f ms/step  %.1f tiles/s  in_sync=%s  bn=%s" % (d["ms_per_step"], d["value"], d["config"].get("replicas_in_sync"), d["config"].get("bn")))
except Exception as e:
    print("   no JSON:", e)
PY
grep -v "^$" $O/r2_n2_$name.err | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -4 | cut -c1-300; }
run sync2_fused 90 env MCB_SYNC_BN=2 MCB_SYNC_FUSED=1 $TR --master-port 29643 bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown
run sync2 90 env MCB_SYNC_BN=2 $TR --master-port 29644 bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown
echo "== tests (all three, incl. both SyncBN modes)"; MCB_TEST_SYNC_BN=1 timeout -k 5 300 python -m pytest tests/test_multi_gpu_gpu.py -q 2>&1 | grep -v "NCCL version" | tail -4 | cut -c1-300
