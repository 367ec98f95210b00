#!/bin/bash
set -u
O=gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -80
echo "== bench (train, default)"; MCB_BENCH_VERBOSE=1 timeout 400 python bench.py --steps 10 --warmup 3 > $O/r2_bench_train.json 2> $O/r2_bench_train.err; echo rc=$?; tail -c 2500 $O/r2_bench_train.json; tail -5 $O/r2_bench_train.err
echo "== layer3 tile sweep"
timeout 200 python - <<'PY'
import os, sys
sys.path.insert(0, ".")
sys.argv = ["sweep_gemm.py", "none"]
exec(open("tools/sweep_gemm.py").read())
SHAPES[:] = [(256, 256, 3, 20), (1024, 256, 1, 20), (256, 1024, 1, 20), (512, 512, 3, 10), (64, 256, 1, 80), (128, 512, 1, 40)]
envs = [{}, {"MCB_FORCE_BN": 64}, {"MCB_FORCE_BN": 128}, {"MCB_FORCE_BN": 256}]
run("fwd", fwd, envs); run("dgradM", dgrad_mask, envs)
PY
