#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout -k 5 $t "$@" > $O/r02_n8_$name.json 2> $O/r02_n8_$name.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02_n8_$name.json").read().strip().splitlines()[-1])
    print("   %.3f ms/step  %.1f tiles/s  e2e %.1f  in_sync=%s  bn=%s clocks=%s" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["config"].get("replicas_in_sync"), d["config"].get("bn"), d.get("clocks")))
except Exception as e:
    print("   no JSON:", e)
PY
grep -v "^$" $O/r02_n8_$name.err | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -4 | cut -c1-300; }
run syncbn_nvlink 120 env MCB_SYNC_BN=2 $TR --master-port 29672 bench.py --gpus 8 --steps 20 --warmup 5 --no-breakdown
run base 120 $TR --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 5 --no-breakdown
