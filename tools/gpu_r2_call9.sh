#!/bin/bash
# round-2 call 9: gemm2 with the TMA-store epilogue; why FUSE=0 + gemm2 stalls; full GPU suite
mkdir -p gpurun_out
echo "=== [0] gemm2 check"
timeout 120 python tools/gpu_gemm2_check.py 2>&1 | tail -16; G2=${PIPESTATUS[0]}; echo "gemm2 rc=$G2"
[ "$G2" = "0" ] || export VLO_VIT_GEMM2=0
echo "=== [1] ViT alone"
timeout 200 python tools/gpu_vit_bench.py --batches 3,4,8 2>&1 | tail -4
echo "=== [2] FUSE=0 diagnosis (stderr tail shown)"
run() { # label, env..., -- bench args
  local label=$1; shift
  local t0=$(date +%s)
  env "$@" timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/diag.json 2> gpurun_out/diag.err; local rc=$?
  echo "[$label] rc=$rc $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/diag.err | cut -c1-300
  python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/diag.json').read().strip().splitlines()[-1])
    print('   value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'li', round(d['e2e']['liveinfer']['value'], 1), 'seq', round(d['run']['sequential_frames_per_s'], 1),
          'classes', {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()})
except Exception as e:
    print('   no json:', e)
PY
}
run "fuse1 gemm2" VLO_ATTN=2 VLO_VIT_ATTN=1
run "fuse0 gemm2" VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1
run "fuse0 gemm2=0" VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1 VLO_VIT_GEMM2=0
run "fuse0 gemm2 ws11" VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1 VLO_WS_STAGES=11
echo "=== [3] full GPU suite"
timeout 900 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --tb=short 2>&1 | tail -8
