"""Dev tool: the vision tower + connector alone (vlo_vit_encode) at batch 1 / 2 / 4 / 8 / ..., device-timed, with the
kernel-class breakdown of the engine's profile hooks.  One JSON line per batch.  Run under gpurun.
  python tools/gpu_vit_bench.py [--batches 1,4,8] [--iters 5]"""
import argparse, ctypes as C, json, pathlib, sys
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import llama3_8b_siglip_l, weights as W
from videollm_online_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--batches", default="1,2,4,8")
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--no-classes", action="store_true")
args = ap.parse_args()
batches = [int(b) for b in args.batches.split(",")]
dev = torch.device("cuda:0")
cfg = llama3_8b_siglip_l()
eng = Engine(cfg, dev, max_streams=1, max_kv_tokens=256, max_step_tokens=128, max_vit_batch=max(batches))
w = W.synthetic_engine_weights(cfg, dev, 256, seed=0)
eng.load_weights({k: v for k, v in w.items() if k.startswith("vit.") or k.startswith("conn.")})
del w
S = cfg.frame_resolution
frames = torch.randint(0, 256, (max(batches), 3, S, S), dtype=torch.uint8, device=dev)
C_, M_, P = cfg.vision_hidden_size, cfg.vision_intermediate_size, cfg.num_patches
flop_frame = 24 * P * (C_ * 3 * C_ + C_ * C_ + 2 * C_ * M_) * 2 + 24 * 4 * P * P * C_   # trunk GEMMs + attention
peak = 1468.8
mp = pathlib.Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
if mp.exists():
    peak = float(json.loads(mp.read_text()).get("bf16_tflops_sustained", peak))
for B in batches:
    for _ in range(2):
        eng.vit_encode(frames[:B])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.iters):
        eng.vit_encode(frames[:B])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.iters
    rec = {"batch": B, "ms_per_call": round(ms, 3), "ms_per_frame": round(ms / B, 3),
           "tflops": round(flop_frame * B / ms / 1e9, 1), "frac_of_sustained_peak": round(flop_frame * B / ms / 1e9 / peak, 3)}
    if not args.no_classes:
        eng.lib.vlo_profile_enable(1)
        eng.vit_encode(frames[:B])
        torch.cuda.synchronize()
        ncls = 6
        msv, n, by = (C.c_double * ncls)(), (C.c_longlong * ncls)(), (C.c_double * ncls)()
        eng.lib.vlo_profile_read(msv, n, by, ncls)
        eng.lib.vlo_profile_enable(0)
        names = ["gemm_weight_stream", "attn_kvappend", "attn_merge", "gemm_vit", "vit_attn", "other"]
        rec["classes_ms"] = {nm: round(msv[j], 3) for j, nm in enumerate(names) if n[j]}
        rec["classes_n"] = {nm: int(n[j]) for j, nm in enumerate(names) if n[j]}
    print(json.dumps(rec), flush=True)
