#!/bin/bash
# round-2 call 16: token-tile width of the co-resident (batch-1) ViT configuration
mkdir -p gpurun_out
for bn in 64 96 128; do
  echo "--- VLO_VIT_SMALL_BN=$bn"
  VLO_VIT_SMALL_BN=$bn timeout 60 python tools/gpu_vit_bench.py --batches 1,2 --iters 5 2>&1 | tail -2 | cut -c1-260
  VLO_VIT_SMALL_BN=$bn timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('   bench value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1), 'seq', round(d['run']['sequential_frames_per_s'],1))"
done
echo "--- parity of the batch-1 path with 96"
VLO_VIT_SMALL_BN=96 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=100 --timeout-method=thread --tb=line -k "vit or visual_embed" 2>&1 | tail -2
