#!/bin/bash
set -u
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== tests"; timeout -k 5 300 python -m pytest tests/test_multi_gpu_gpu.py -q 2>&1 | grep -v "NCCL version" | tail -2 | cut -c1-300
timeout -k 5 90 $TR --master-port 29691 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown > $O/r02_n2_base.json 2> $O/r02_n2_base.err; echo rc=$?
python -c "
import json; d=json.loads(open('$O/r02_n2_base.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('replicas_in_sync'))"
timeout -k 5 90 env MCB_SYNC_BN=2 $TR --master-port 29692 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown > $O/r02_n2_syncbn_nvlink.json 2> $O/r02_n2_sync.err; echo rc=$?
python -c "
import json; d=json.loads(open('$O/r02_n2_syncbn_nvlink.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['config'].get('replicas_in_sync'))"
