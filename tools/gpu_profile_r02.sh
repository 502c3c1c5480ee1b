#!/bin/bash
# r02 ncu evidence (one GPU).  Numbers printed by bench.py under ncu are NOT bench values.
#   [a] launch list of ~2 frame steps (durations only)    -> gpurun_out/launches_r02.csv
#   [b] --set full of the decoder weight-streaming GEMMs   -> gpurun_out/prof_gemm_ws_r02.ncu-rep
#   [c] --set full of the KV-append attention + merge       -> gpurun_out/prof_attn_r02.ncu-rep
set -u
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras --encode-ahead 1"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vlo:: \
    -s 2500 -c 1400 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_ws_kernel<.int.1, .int.16|gemm_wsf_kernel" -s 500 -c 8 -o gpurun_out/prof_gemm_ws_r02 $B > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:attn_tc2_kernel|attn_tc_kernel|attn_merge_kernel" -s 130 -c 6 -o gpurun_out/prof_attn_r02 $B > gpurun_out/ncu_attn.log 2>&1
echo "attn rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02.csv
