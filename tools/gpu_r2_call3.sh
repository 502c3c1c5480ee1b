#!/bin/bash
# round-2 call 3: each new kernel is checked in its own short-fused process first; a failing one is switched off for the rest
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.jsonl gpurun_out/parity_observed.jsonl
PT="python -m pytest -m gpu -q -x --timeout=90 --timeout-method=thread -p no:cacheprovider"
echo "=== [a] attention v3 check"
timeout 90 python tools/gpu_attn_check.py > gpurun_out/attn_check_v3.log 2>&1; rc=$?; echo "rc=$rc"; tail -15 gpurun_out/attn_check_v3.log
if [ $rc -ne 0 ]; then echo "!! attention v3 failed -> VLO_ATTN=2 for the rest"; export VLO_ATTN=2; fi
echo "=== [b] attention A/B"
for cfg in "2" "3"; do
  [ "$cfg" = "3" ] && [ "${VLO_ATTN:-3}" = "2" ] && continue
  VLO_ATTN=$cfg timeout 100 python tools/gpu_attn_ab.py 2>&1 | tail -4
done
echo "=== [c] fused stream-K GEMMs (tiny model parity, old ViT attention)"
VLO_VIT_ATTN=1 VLO_FUSE=1 timeout 200 $PT tests/test_gpu_parity.py -k "chunked or one_pass or greedy or batched or long_prompt or truncate" 2>&1 | tail -8
rc=${PIPESTATUS[0]}; echo "rc=$rc"
if [ $rc -ne 0 ]; then echo "!! fused GEMM failed -> VLO_FUSE=0 for the rest"; export VLO_FUSE=0; fi
echo "=== [d] tcgen05 ViT attention (tiny + full-size ViT parity, unfused decoder)"
VLO_FUSE=0 VLO_VIT_ATTN=2 timeout 300 $PT tests/test_gpu_parity.py -k "vit_tokens or visual_embed or connector_only" 2>&1 | tail -8
rc=${PIPESTATUS[0]}; echo "rc=$rc"
if [ $rc -ne 0 ]; then echo "!! ViT tcgen05 attention failed -> VLO_VIT_ATTN=1 for the rest"; export VLO_VIT_ATTN=1; fi
echo "=== [e] full GPU suite (env: ATTN=${VLO_ATTN:-3} FUSE=${VLO_FUSE:-1} VIT_ATTN=${VLO_VIT_ATTN:-2})"
timeout 1000 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --durations=8 2>&1 | tail -40
cat gpurun_out/parity_observed.jsonl 2>/dev/null
echo "=== [f] bench (quick extras)"
timeout 600 python bench.py --steps 20 --warmup 5 --quick-extras --no-cpu-baseline > gpurun_out/bench_call3.json 2> gpurun_out/bench_call3.err; echo "rc=$?"
tail -5 gpurun_out/bench_call3.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_call3.json').read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks') if k in d}
    keep['e2e'] = d['e2e']; keep['roofline_frac'] = d['roofline']['frac']; keep['roofline_us'] = d['roofline']['avg_us_per_launch']
    keep['attn'] = {k: d['roofline_attn'][k] for k in ('frac', 'avg_us_per_launch', 'main_kernel_only')}
    keep['step_frac'] = d['roofline_step']['frac']; keep['run'] = d.get('run'); keep['extras'] = d.get('extras')
    print(json.dumps(keep, indent=1))
except Exception as e:
    print('bench parse failed', e)
PY
