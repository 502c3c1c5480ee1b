#!/bin/bash
# round-2 call 11: which PDL use of the 2-CTA GEMM dead-locks next to the other ViT kernels (VLO_GEMM2_PDL bit 0 = launched
# with the attribute, bit 1 = triggers its successor early); then perf + parity with the safe setting
mkdir -p gpurun_out
DBG=$PWD/videollm-online_b200/libvlo_b200_dbg.so
run() { local label=$1; shift; local t0=$(date +%s); env "$@" > gpurun_out/dbg.out 2>&1; local rc=$?; echo "[$label] rc=$rc $(( $(date +%s) - t0 )) s"; grep -v "^$" gpurun_out/dbg.out | sort | uniq -c | sort -rn | head -6 | cut -c1-260; return $rc; }
GOOD=""
for m in 0 2 1 3; do
  if run "vit B=3 gemm2_pdl=$m" VLO_LIB=$DBG VLO_GEMM2_PDL=$m timeout 40 python tools/gpu_vit_bench.py --batches 3 --iters 2 --no-classes; then GOOD="$GOOD $m"; fi
done
echo "modes that pass:$GOOD"
BEST=0
for m in $GOOD; do
  echo "--- perf mode $m"
  VLO_GEMM2_PDL=$m timeout 60 python tools/gpu_vit_bench.py --batches 3,4,8 --iters 5 2>&1 | tail -3 | cut -c1-330
done
echo "=== parity (full-size ViT, tiny ViT) with mode 0 and each passing mode"
for m in $GOOD; do
  VLO_GEMM2_PDL=$m timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=100 --timeout-method=thread --tb=line -k "vit or visual_embed or connector" 2>&1 | tail -2
done
