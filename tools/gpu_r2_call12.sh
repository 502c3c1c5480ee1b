#!/bin/bash
# round-2 call 12: bench A/B with the 2-CTA ViT GEMMs (safe PDL mode) + ViT attention gen 3; full GPU suite
mkdir -p gpurun_out; rm -f gpurun_out/parity_observed.jsonl
ab() { # label, env...
  local label=$1; shift
  local t0=$(date +%s)
  env "$@" timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/ab.json 2> gpurun_out/ab.err; local rc=$?
  python - "$label" "$rc" "$(( $(date +%s) - t0 ))" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
    print(f"[{sys.argv[1]}] rc={sys.argv[2]} {sys.argv[3]}s", 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'li', round(d['e2e']['liveinfer']['value'], 1),
          'seq', round(d['run']['sequential_frames_per_s'], 1), 'gemm', round(d['roofline']['frac'], 3), 'attn', round(d['roofline_attn']['frac'], 3), 'step', round(d['roofline_step']['frac'], 3))
except Exception as e:
    print(f"[{sys.argv[1]}] rc={sys.argv[2]} {sys.argv[3]}s no json:", e); print(open('gpurun_out/ab.err').read()[-400:])
PY
}
echo "=== [1] A/B (encode-ahead 4 unless stated)"
ab "fuse1 attn3" X=1
ab "fuse0 attn2" VLO_FUSE=0 VLO_ATTN=2
ab "fuse0 attn3" VLO_FUSE=0
ab "fuse0 attn2 ws11" VLO_FUSE=0 VLO_ATTN=2 VLO_WS_STAGES=11
ab "fuse0 attn2 gemm2=0" VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_GEMM2=0
echo "--- encode-ahead 8 / 1"
env VLO_FUSE=0 VLO_ATTN=2 timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --encode-ahead 8 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('[fuse0 attn2 D=8] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1))"
env VLO_FUSE=0 VLO_ATTN=2 timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras --encode-ahead 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('[fuse0 attn2 D=1] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1))"
echo "=== [2] full GPU suite"
timeout 600 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread --tb=short 2>&1 | tail -8
