#!/bin/bash
# round-2 call 6 (re-entry): status of HEAD — attention check + A/B, full GPU suite, quick bench
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.jsonl gpurun_out/parity_observed.jsonl
echo "=== [1] attention check (v3)"
timeout 120 python tools/gpu_attn_check.py 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"
echo "=== [2] attention A/B v2, v3"
VLO_ATTN=2 timeout 100 python tools/gpu_attn_ab.py 2>&1 | tail -4
VLO_ATTN=3 timeout 100 python tools/gpu_attn_ab.py 2>&1 | tail -4
echo "=== [3] full GPU suite"
timeout 1000 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --tb=short --durations=6 2>&1 | tail -40
echo "--- observed"; cat gpurun_out/parity_observed.jsonl 2>/dev/null
echo "=== [4] bench (quick extras)"
timeout 500 python bench.py --steps 20 --warmup 5 --quick-extras --no-cpu-baseline > gpurun_out/bench_call6.json 2> gpurun_out/bench_call6.err; echo "rc=$?"
tail -3 gpurun_out/bench_call6.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_call6.json').read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks') if k in d}
    keep['e2e'] = d['e2e']; keep['roofline_frac'] = d['roofline']['frac']; keep['roofline_us'] = d['roofline']['avg_us_per_launch']
    keep['attn'] = {k: d['roofline_attn'][k] for k in ('frac', 'avg_us_per_launch', 'main_kernel_only')}
    keep['step_frac'] = d['roofline_step']['frac']; keep['run'] = d.get('run'); keep['extras'] = d.get('extras')
    keep['classes'] = {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()}
    print(json.dumps(keep, indent=1))
except Exception as e:
    print('bench parse failed', e)
PY
