#!/bin/bash
# in-graph cost of each kernel family: step time with the family removed from the plan vs the full step
# usage (GPU box): bash tools/knockout.sh > gpurun_out/knockout.log
run() { MCB_KNOCKOUT="$1" timeout 200 python bench.py --steps 6 --warmup 3 --no-breakdown --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-40s %.3f ms/step' % ('$1' or 'full', d['ms_per_step']))"; }
run ""
for k in bn_apply bn_bwd_apply bn_bwd_reduce channel_sum conv_wgrad conv_dgrad conv_fwd stem_im2col final_conv \
         "convt_fwd,convt_dgrad,convt_wgrad" "bn_apply,bn_bwd_apply,bn_bwd_reduce,channel_sum"; do run "$k"; done
