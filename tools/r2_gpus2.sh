#!/bin/bash
# 2-GPU evidence: multi-GPU tests (replica identity, SyncBN through NCCL and through the NVLink exchange) + bench arms
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run() { name=$1; t=$2; shift 2; echo "== $name"; timeout -k 5 $t "$@" > $O/r02_n2_$name.json 2> $O/r02_n2_$name.err; echo "rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02_n2_$name.json").read().strip().splitlines()[-1])
    print("   %.3f ms/step  %.1f tiles/s  in_sync=%s  bn=%s  ar=%s" % (d["ms_per_step"], d["value"], d["config"].get("replicas_in_sync"), d["config"].get("bn"), d["config"].get("grad_allreduce")))
except Exception as e:
    print("   no JSON:", e)
PY
grep -v "^$" $O/r02_n2_$name.err | grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" | tail -4 | cut -c1-300; }
echo "== tests"; timeout -k 5 300 python -m pytest tests/test_multi_gpu_gpu.py -q 2>&1 | grep -v "NCCL version" | tail -4 | cut -c1-300
run base 90 $TR --master-port 29661 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown
run syncbn_nvlink 90 env MCB_SYNC_BN=2 $TR --master-port 29662 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown
echo "== 1-GPU leftovers on this box: RLE edge cases, smoke"; timeout 200 python -m pytest tests/test_instances_gpu.py tests/test_input_gpu.py -q 2>&1 | tail -3; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
