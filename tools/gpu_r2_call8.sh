#!/bin/bash
# round-2 call 8: ViT alone per batch (both attention kernels), fused-GEMM ring depth A/B, ncu evidence (ViT kernels, decoder)
mkdir -p gpurun_out
echo "=== [1] ViT alone, tcgen05 attention"
timeout 200 python tools/gpu_vit_bench.py 2>&1 | tail -5
echo "=== [1b] ViT alone, mma.sync attention"
VLO_VIT_ATTN=1 timeout 200 python tools/gpu_vit_bench.py --batches 1,8 2>&1 | tail -3
echo "=== [2] fused-GEMM ring depth"
for s in "VLO_WSF_STAGES=6" "VLO_WSF_STAGES=5" "VLO_WSF_STAGES=4" "VLO_FUSE=0" "VLO_FUSE=0 VLO_WS_STAGES=5"; do
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
echo "=== [3] ncu: ViT kernels at batch 8 (set full)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm_ws_kernel<.int.0|vit_attn" -s 40 -c 12 -o gpurun_out/prof_vit_b8_r02 \
    python tools/gpu_vit_bench.py --batches 8 --iters 1 --no-classes > gpurun_out/ncu_vit.log 2>&1
echo "vit ncu rc=$?"; tail -2 gpurun_out/ncu_vit.log
ls -la gpurun_out/*.ncu-rep
