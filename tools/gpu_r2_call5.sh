#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/attn_ab.jsonl gpurun_out/parity_observed.jsonl
echo "=== [1] attention debug (v3)"
timeout 150 python tools/gpu_attn_debug.py 2>&1 | tail -32
echo "=== [2] attention A/B v3"
VLO_ATTN=3 timeout 100 python tools/gpu_attn_ab.py 2>&1 | tail -4
echo "=== [3] attention trace (main kernel only, then with merge)"
SKIP_MERGE=1 timeout 100 python tools/gpu_attn_trace2.py 2>&1 | tail -60
SKIP_MERGE=0 timeout 100 python tools/gpu_attn_trace2.py 2>&1 | grep -E "loop:|globaltimer|grid dependency|partials|CTA end"
echo "=== [4] full GPU suite"
timeout 1000 python -m pytest tests -m gpu -q --timeout=300 --timeout-method=thread --tb=short 2>&1 | tail -40
cp gpurun_out/parity_observed.jsonl gpurun_out/parity_observed_new.jsonl 2>/dev/null
echo "=== [5] 32-layer + ragged parity on the round-1 kernels (VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1)"
VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1 timeout 400 python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout=300 --timeout-method=thread --tb=line -k "full_depth or ragged" 2>&1 | tail -8
echo "--- observed (new kernels)"; cat gpurun_out/parity_observed_new.jsonl 2>/dev/null
echo "--- observed (round-1 kernels, last run)"; tail -4 gpurun_out/parity_observed.jsonl 2>/dev/null
echo "=== [6] bench (quick extras), encode-ahead 4 and 1"
for d in 4 1; do
timeout 500 python bench.py --steps 20 --warmup 5 --quick-extras --no-cpu-baseline --encode-ahead $d > gpurun_out/bench_call5_d$d.json 2> gpurun_out/bench_call5_d$d.err; echo "rc=$?"
tail -3 gpurun_out/bench_call5_d$d.err
python - "$d" <<'PY'
import json, sys
d_ = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/bench_call5_d{d_}.json').read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks') if k in d}
    keep['e2e'] = d['e2e']; keep['roofline_frac'] = d['roofline']['frac']; keep['roofline_us'] = d['roofline']['avg_us_per_launch']
    keep['attn'] = {k: d['roofline_attn'][k] for k in ('frac', 'avg_us_per_launch', 'main_kernel_only')}
    keep['step_frac'] = d['roofline_step']['frac']; keep['run'] = d.get('run'); keep['extras'] = d.get('extras')
    keep['classes'] = {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()}
    print(json.dumps(keep, indent=1))
except Exception as e:
    print('bench parse failed', e)
PY
done
