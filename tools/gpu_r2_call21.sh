#!/bin/bash
# round-2 call 21: deeper ring for the fused gate|up GEMM (8 / 10 stages: no co-resident ViT CTA during that GEMM)
for s in "X=0" "VLO_WSF_STAGES=8" "VLO_WSF_STAGES=10"; do
  env $s timeout 60 python tools/gpu_step_bench.py 2>&1 | tail -1 | cut -c1-200
done
for s in "X=0" "VLO_WSF_STAGES=8" "VLO_WSF_STAGES=10"; do
  env $s timeout 100 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('[$s] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1))"
done
