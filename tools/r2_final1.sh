#!/bin/bash
# final 1-GPU evidence of round 2 -> gpurun_out/ (copied into profiles/ afterwards)
set -u
O=gpurun_out
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== bench (train, default)"; timeout 400 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo rc=$?; tail -c 600 $O/r02_bench_n1.json; tail -3 $O/r02_bench_n1.err
echo "== bench (infer)"; timeout 500 python bench.py --workload infer --steps 5 --warmup 3 > $O/r02_bench_infer.json 2> $O/r02_bench_infer.err; echo rc=$?; tail -c 900 $O/r02_bench_infer.json; tail -3 $O/r02_bench_infer.err
echo "== A/B tile rule"; bash tools/envsweep.sh MCB_BN256_MIN_WAVE_X10=10
echo "== ncu launch list of the train step"; REPLAYS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_launches_ncu.csv python tools/ncu_step.py > $O/r02_launches_ncu.out 2>&1; echo rc=$?; tail -3 $O/r02_launches_ncu.out; wc -l $O/r02_launches_ncu.csv
echo "== ncu post-processing"; timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_postproc_ncu.csv python tools/ncu_postproc.py > $O/r02_postproc_ncu.out 2>&1; echo rc=$?; tail -3 $O/r02_postproc_ncu.out
