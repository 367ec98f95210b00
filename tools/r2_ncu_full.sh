#!/bin/bash
set -u
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:crf_kernel -s 7 -c 1 -o $O/r02_crf_full python tools/ncu_crf.py > $O/r02_crf_full.out 2>&1; echo rc=$?; tail -2 $O/r02_crf_full.out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 2 -c 1 -o $O/r02_conv_l3_full python tools/ncu_one.py fwd 256 256 3 20 > $O/r02_conv_l3_full.out 2>&1; echo rc=$?; tail -2 $O/r02_conv_l3_full.out
ls -la $O/*.ncu-rep
