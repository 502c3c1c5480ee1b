#!/bin/bash
# round-2 call 8: 2-CTA ViT GEMM check; ViT alone per batch (old / new GEMMs, both attention kernels); fused-GEMM ring depth
# A/B; ncu evidence of the ViT kernels
mkdir -p gpurun_out
echo "=== [0] gemm2 check"
timeout 120 python tools/gpu_gemm2_check.py 2>&1 | tail -16; G2=${PIPESTATUS[0]}; echo "gemm2 rc=$G2"
echo "=== [1] ViT alone, single-CTA GEMMs (VLO_VIT_GEMM2=0), tcgen05 attention"
VLO_VIT_GEMM2=0 timeout 200 python tools/gpu_vit_bench.py 2>&1 | tail -5
echo "=== [1b] same, mma.sync attention"
VLO_VIT_GEMM2=0 VLO_VIT_ATTN=1 timeout 200 python tools/gpu_vit_bench.py --batches 1,8 2>&1 | tail -3
if [ "$G2" = "0" ]; then
echo "=== [1c] ViT alone, 2-CTA GEMMs"
timeout 200 python tools/gpu_vit_bench.py --batches 3,4,8 2>&1 | tail -4
echo "=== [1d] full-size ViT parity with the 2-CTA GEMMs (batch 3)"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=200 --timeout-method=thread --tb=short -k "full_size_vit" 2>&1 | tail -5
else
export VLO_VIT_GEMM2=0
fi
echo "=== [2] fused-GEMM ring depth"
for s in "VLO_WSF_STAGES=6" "VLO_WSF_STAGES=5" "VLO_WSF_STAGES=4" "VLO_FUSE=0" "VLO_FUSE=0 VLO_WS_STAGES=5" "VLO_FUSE=0 VLO_WS_STAGES=11"; do
  env $s VLO_ATTN=2 VLO_VIT_ATTN=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    r = d.get('roofline', {})
    print('[$s]', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'seq', round(d.get('run',{}).get('sequential_frames_per_s',0),1),
          'gemm_frac', round(r.get('frac',0),3), 'classes', {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()})
"
done
echo "=== [3] ncu: ViT kernels at batch 8 (set full), single-CTA GEMMs"
VLO_VIT_GEMM2=0 timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_ws_kernel<.int.0|vit_attn" -s 40 -c 8 -o gpurun_out/prof_vit_b8_old_r02 \
    python tools/gpu_vit_bench.py --batches 8 --iters 1 --no-classes > gpurun_out/ncu_vit_old.log 2>&1
echo "vit ncu (old) rc=$?"; tail -2 gpurun_out/ncu_vit_old.log
if [ "$G2" = "0" ]; then
timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm2_kernel|vit_attn" -s 40 -c 10 -o gpurun_out/prof_vit_b8_r02 \
    python tools/gpu_vit_bench.py --batches 8 --iters 1 --no-classes > gpurun_out/ncu_vit.log 2>&1
echo "vit ncu (gemm2) rc=$?"; tail -2 gpurun_out/ncu_vit.log
fi
ls -la gpurun_out/*.ncu-rep
