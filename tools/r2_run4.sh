#!/bin/bash
set -u
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus 4 --steps 20 --warmup 5 --no-breakdown > $O/r02_n4.json 2> $O/r02_n4.err; echo rc=$?
python -c "
import json; d=json.loads(open('$O/r02_n4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config'].get('replicas_in_sync'), d.get('clocks'))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29682 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 > $O/r02_n4_ref.json 2> $O/r02_n4_ref.err; echo rc=$?; cut -c1-300 $O/r02_n4_ref.json
