#!/bin/bash
# round-2 call 2: key-sliced attention (v3) correctness + A/B vs v2, then the whole GPU test suite
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.jsonl gpurun_out/parity_observed.jsonl
echo "=== check v3"
timeout 300 python tools/gpu_attn_check.py > gpurun_out/attn_check_v3.log 2>&1; echo "rc=$?"; tail -16 gpurun_out/attn_check_v3.log
for cfg in "2" "3"; do
  echo "=== ab VLO_ATTN=$cfg"
  VLO_ATTN=$cfg timeout 300 python tools/gpu_attn_ab.py 2>&1 | tail -5
done
echo "=== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 2>&1 | tail -40
cat gpurun_out/parity_observed.jsonl
