#!/bin/bash
# round-2 call 18: which stream-K fix-ups are worth fusing into the GEMM finisher (VLO_FUSE bit mask: 1 q|k|v, 2 o_proj,
# 4 gate|up, 8 down_proj), decoder step alone + pipelined bench; then the non-default-kernel-path tests
mkdir -p gpurun_out
for m in 0 15 1 2 4 8 5 10; do
  VLO_FUSE=$m timeout 60 python tools/gpu_step_bench.py 2>&1 | tail -1 | cut -c1-200
done
for m in 0 4 5; do
  VLO_FUSE=$m timeout 100 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('[VLO_FUSE=$m] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1))"
done
timeout 400 python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout=300 --timeout-method=thread --tb=short -k non_default 2>&1 | tail -6
