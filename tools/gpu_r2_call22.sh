#!/bin/bash
# round-2 call 22: confirm the 8-stage fused gate|up default (parity subset + bench twice against 6 stages)
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=100 --timeout-method=thread --tb=line -k "chunked or greedy or batched or full_width or state_machine" 2>&1 | tail -2
for s in "X=0" "VLO_WSF_STAGES=6" "X=1"; do
  env $s timeout 100 python bench.py --steps 50 --warmup 6 --no-cpu-baseline --no-extras > gpurun_out/b22.json 2>/dev/null; python -c "
import json
d = json.loads(open('gpurun_out/b22.json').read().strip().splitlines()[-1])
print('[$s] value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'li', round(d['e2e']['liveinfer']['value'],1), 'step', round(d['roofline_step']['frac'],3))"
  [ "$s" = "X=1" ] && cp gpurun_out/b22.json gpurun_out/bench_final_noextras.json
done
