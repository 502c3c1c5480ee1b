#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU).  Numbers printed by bench.py under ncu are NOT bench values.
set -u
TAG=${1:-r01}
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras"
# every launch of our kernels with its device time, ~2 frame steps of the timed region
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vlo:: \
    -s 2080 -c 1400 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# the decoder weight-streaming GEMM (qkv, o, gate|up, down of consecutive layers)
timeout 700 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_ws_kernel<.int.1, .int.16" -s 430 -c 8 -o gpurun_out/prof_gemm_ws_${TAG} $B > gpurun_out/ncu_gemm.log 2>&1
echo "gemm_ws rc=$?"
# the KV-append attention kernel
timeout 700 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:vlo::attn_tc_kernel -s 110 -c 3 -o gpurun_out/prof_attn_tc_${TAG} $B > gpurun_out/ncu_attn.log 2>&1
echo "attn_tc rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_${TAG}.csv
