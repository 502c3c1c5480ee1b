#!/bin/bash
# round-2 call 7: (1) the two full-width parity tests on the new and on the round-1 kernels, observed statistics;
# (2) per-CTA timeline of the v3 attention; (3) same-box A/B of the round-2 kernel switches through bench.py
mkdir -p gpurun_out; rm -f gpurun_out/parity_observed.jsonl
PT="python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout=300 --timeout-method=thread --tb=line -k full_depth_or_ragged"
PT="python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout=300 --timeout-method=thread --tb=line"
echo "=== [1a] new kernels"
timeout 400 $PT -k "full_depth or ragged" 2>&1 | tail -6
echo "=== [1b] round-1 kernels (VLO_FUSE=0 VLO_ATTN=2)"
VLO_FUSE=0 VLO_ATTN=2 timeout 400 $PT -k "full_depth or ragged" 2>&1 | tail -6
echo "=== [1c] VLO_FUSE=0 VLO_ATTN=1 (mma.sync attention: independent implementation)"
VLO_FUSE=0 VLO_ATTN=1 timeout 400 $PT -k "ragged" 2>&1 | tail -6
echo "--- observed"; cat gpurun_out/parity_observed.jsonl
echo "=== [2] attention trace (v3): main kernel only, then with merge"
SKIP_MERGE=1 timeout 100 python tools/gpu_attn_trace2.py 2>&1 | tail -62
SKIP_MERGE=0 timeout 100 python tools/gpu_attn_trace2.py 2>&1 | grep -E "loop:|globaltimer|grid dependency|partials|CTA end|setup done"
echo "=== [3] A/B"
for s in "X=0" "VLO_FUSE=0" "VLO_ATTN=2" "VLO_VIT_ATTN=1" "VLO_FUSE=0 VLO_ATTN=2 VLO_VIT_ATTN=1" "X=1"; do
  env $s timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    ra = d.get('roofline_attn', {}); r = d.get('roofline', {})
    print('[$s]', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'seq', round(d.get('run',{}).get('sequential_frames_per_s',0),1),
          'gemm_frac', round(r.get('frac',0),3), 'attn_pair_us', round(ra.get('avg_us_per_launch',0),2), 'launches', d.get('gpu_launches'),
          'classes', {k: round(v['ms_per_step'], 3) for k, v in d.get('kernel_classes', {}).items()})
"
done
