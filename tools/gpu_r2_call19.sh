#!/bin/bash
# round-2 call 19: encode-ahead 1 vs 2 (both in the co-resident small-tile ViT regime)
for d in 1 2 1 2; do
  timeout 100 python bench.py --steps 40 --warmup 6 --no-cpu-baseline --no-extras --encode-ahead $d 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('[D=$d] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1))"
done
