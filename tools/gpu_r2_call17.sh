#!/bin/bash
# round-2 call 17: event-driven issue order in the two-tile ViT attention: parity, then speed
mkdir -p gpurun_out
DBG=$PWD/videollm-online_b200/libvlo_b200_dbg.so
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=60 --timeout-method=thread --tb=short -k "vit or visual_embed or connector" 2>&1 | tail -4
timeout 80 python tools/gpu_vit_bench.py --batches 3,4,8 --iters 5 2>&1 | tail -3 | cut -c1-330
