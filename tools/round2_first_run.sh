#!/bin/bash
# First gpurun of round 2: exercise the switches that were prepared but not run in round 1.
#   1 GPU :  /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/round2_first_run.sh one > gpurun_out/r2_first.log 2>&1'
#   2 GPUs:  /usr/local/graft/bin/gpurun --gpus 2 --timeout 600 -- 'bash tools/round2_first_run.sh two > gpurun_out/r2_first2.log 2>&1'
set -u
case "${1:-one}" in
one)
  echo "== resident-weights haloed path: gated parity tests"
  MCB_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -x -q -m gpu -k "haloed" 2>&1 | tail -3
  echo "== sweep of the thin layers with and without MCB_BRES"
  timeout 200 python - <<'PY'
import os, sys
sys.path.insert(0, ".")
sys.argv = ["sweep_gemm.py", "none"]
exec(open("tools/sweep_gemm.py").read())
SHAPES[:] = [(32, 32, 3, 320), (64, 64, 3, 80), (64, 64, 3, 160)]
envs = [{}, {"MCB_BRES": 1}]
run("fwd", fwd, envs); run("dgrad", dgrad, envs); run("dgradF", dgrad_fused, envs); run("dgradM", dgrad_mask, envs)
PY
  echo "== full step A/B"
  bash tools/envsweep.sh MCB_BRES=1
  ;;
two)
  echo "== 2-GPU parity tests (replica identity, SyncBN vs single process)"
  MCB_TEST_SYNC_BN=1 timeout 500 python -m pytest tests/test_multi_gpu_gpu.py -x -q -m gpu 2>&1 | tail -5
  for v in "MCB_X=0" "MCB_OVERLAP_ALLREDUCE=2" "MCB_OVERLAP_ALLREDUCE=1" "MCB_SYNC_BN=1"; do
    env $v timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-breakdown 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-28s %.3f ms/step %.1f tiles/s in_sync=%s' % ('$v', d['ms_per_step'], d['value'], d['config'].get('replicas_in_sync')))"
  done
  ;;
esac
