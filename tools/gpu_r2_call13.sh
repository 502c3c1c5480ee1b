#!/bin/bash
# round-2 call 13: new defaults (encode-ahead 1, unfused chain, attention by context length, pair-GEMM ViT serialised against
# decoder steps): bench + full suite, then the r02 ncu evidence
mkdir -p gpurun_out; rm -f gpurun_out/parity_observed.jsonl
summ() { python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    keep = {k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches') if k in d}
    keep['e2e'] = round(d['e2e']['value'], 1); keep['liveinfer'] = round(d['e2e']['liveinfer']['value'], 1)
    keep['gemm_frac'] = round(d['roofline']['frac'], 3); keep['attn'] = {'frac': round(d['roofline_attn']['frac'], 3), 'us': round(d['roofline_attn']['avg_us_per_launch'], 2), 'main_us': round(d['roofline_attn']['main_kernel_only']['avg_us_per_launch'], 2)}
    keep['step_frac'] = round(d['roofline_step']['frac'], 3); keep['seq_fps'] = round(d['run']['sequential_frames_per_s'], 1)
    ex = d.get('extras') or {}
    keep['extras'] = {k: {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk != 'note'} for k, v in ex.items()}
    keep['classes'] = {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()}
    print(json.dumps(keep, indent=1))
except Exception as e:
    print('bench parse failed', e)
PY
}
echo "=== [1] bench, defaults, quick extras"
timeout 420 python bench.py --steps 40 --warmup 5 --quick-extras --no-cpu-baseline > gpurun_out/bench_call13.json 2> gpurun_out/bench_call13.err; echo "rc=$?"
tail -2 gpurun_out/bench_call13.err | cut -c1-300; summ gpurun_out/bench_call13.json
echo "=== [2] bench, encode-ahead 4 (pair GEMMs, serialised), no extras"
timeout 120 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline --encode-ahead 4 > gpurun_out/bench_call13_d4.json 2> gpurun_out/bench_call13_d4.err; echo "rc=$?"
summ gpurun_out/bench_call13_d4.json | head -12
echo "=== [3] full GPU suite"
timeout 600 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread --tb=short 2>&1 | tail -6
echo "=== [4] ncu evidence r02"
bash tools/gpu_profile_r02.sh 2>&1 | tail -8
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:gemm2_kernel|vit_attn" -s 40 -c 10 -o gpurun_out/prof_vit_b8_r02 \
    python tools/gpu_vit_bench.py --batches 8 --iters 1 --no-classes > gpurun_out/ncu_vit.log 2>&1
echo "vit ncu rc=$?"; tail -1 gpurun_out/ncu_vit.log
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv 2>/dev/null
