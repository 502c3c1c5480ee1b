#!/bin/bash
# round-2 call 1: new attention kernel (v3) correctness + micro-loop A/B against v2
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for cfg in "3 128" "3 64"; do
  set -- $cfg
  echo "=== check VLO_ATTN=$1 BLK=$2"
  VLO_ATTN=$1 VLO_ATTN_BLK=$2 timeout 300 python tools/gpu_attn_check.py > gpurun_out/attn_check_$1_$2.log 2>&1
  echo "rc=$?"; tail -22 gpurun_out/attn_check_$1_$2.log
done
for cfg in "2 128" "3 128" "3 64"; do
  set -- $cfg
  echo "=== ab VLO_ATTN=$1 BLK=$2"
  VLO_ATTN=$1 VLO_ATTN_BLK=$2 timeout 300 python tools/gpu_attn_ab.py 2>&1 | tail -6
done
