#!/bin/bash
# end-of-round validation: smoke, full GPU suite, the default bench (full extras + CPU baseline), the reference arm
mkdir -p gpurun_out; rm -f gpurun_out/parity_observed.jsonl
echo "=== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== GPU suite"; timeout 600 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread --tb=short 2>&1 | tail -5
echo "=== bench (defaults)"
t0=$(date +%s); timeout 700 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$? $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/bench_final.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'clocks')})
print('e2e', d['e2e']['value'], 'liveinfer', d['e2e']['liveinfer']['value'], 'roofline', d['roofline']['frac'], 'attn', d['roofline_attn']['frac'], 'step', d['roofline_step']['frac'])
print('cpu_baseline', {k: v for k, v in d['cpu_baseline'].items() if k != 'sample'})
print({k: {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != 'note'} for k, v in d['extras'].items()})
PY
echo "=== reference arm (3 steps)"
t0=$(date +%s); timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print({k: d[k] for k in ('impl','value','ms_per_step','steps')}, {k: v for k, v in d['cpu_baseline'].items() if k not in ('sample','note')})"
echo "$(( $(date +%s) - t0 )) s"
