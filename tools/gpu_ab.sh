#!/bin/bash
# A/B two builds of the library inside ONE gpurun call (box-to-box variance is ~10 %).
# usage: tools/gpu_ab.sh [bench args]; libs: libvlo_b200_prev.so (A) vs libvlo_b200.so (B)
set -u
ARGS=${*:---steps 40 --warmup 5 --no-cpu-baseline}
for rep in 1 2; do
  for lib in libvlo_b200_prev.so libvlo_b200.so; do
    VLO_LIB=$PWD/videollm-online_b200/$lib timeout 300 python bench.py $ARGS 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    try: d = json.loads(line)
    except Exception: continue
    ra = d.get('roofline_attn', {})
    print('$lib', 'value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'seq', d['config'].get('sequential_fps'),
          'attn_pair_us', ra.get('avg_us_per_launch'), 'frac', round(ra.get('frac',0),3), 'main', ra.get('main_kernel_only'), 'attn8', (d.get('extras') or {}).get('attn_kvappend_8streams', {}).get('us_per_launch_incl_merge'), 'ar', (d.get('extras') or {}).get('ar_decode', {}).get('tokens_per_s'), 'ms8', (d.get('extras') or {}).get('multistream8', {}).get('frames_per_s'))
"
  done
done
