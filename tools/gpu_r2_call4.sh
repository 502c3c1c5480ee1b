#!/bin/bash
mkdir -p gpurun_out
echo "=== attention debug (v3)"
timeout 150 python tools/gpu_attn_debug.py 2>&1 | tail -40
echo "=== attention debug (v2, for comparison)"
VLO_ATTN=2 timeout 150 python tools/gpu_attn_debug.py 2>&1 | tail -34
echo "=== failing tests, short tracebacks"
timeout 400 python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout=200 --timeout-method=thread --tb=short -k "scheduler or ragged" 2>&1 | tail -60
