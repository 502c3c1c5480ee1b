#!/bin/bash
# round-2 call 20: ring depth of the (only) fused GEMM, gate|up: 6 stages = 118 KB (alone on its SM) vs 5 = 100 KB (the down_proj
# GEMM's CTA becomes resident beside it and fills its ring during the finisher tail)
for s in "X=0" "VLO_WSF_STAGES=5" "VLO_WSF_STAGES=4" "X=1" "VLO_WSF_STAGES=5"; do
  env $s timeout 60 python tools/gpu_step_bench.py 2>&1 | tail -1 | cut -c1-200
done
