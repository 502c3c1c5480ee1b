"""Dev tool: timeline of the tcgen05 attention kernel inside a back-to-back (PDL-chained) launch loop, i.e. the
regime bench.py's roofline_attn micro-loop measures.  Run under gpurun.  VLO_ATTN_TRACE=1 stamps clock64 per CTA."""
import ctypes as C, os, pathlib, sys
os.environ["VLO_ATTN_TRACE"] = "1"
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import llama3_8b_siglip_l, weights as W
from videollm_online_b200.engine import Engine

dev = torch.device("cuda:0")
cfg = llama3_8b_siglip_l()
cfg.num_hidden_layers = 6                      # 6 x 50 MB of KV > L2
KV = int(os.environ.get("KV", "12100"))
cap = (KV + 256 + 127) // 128 * 128
eng = Engine(cfg, dev, max_streams=1, max_kv_tokens=cap, max_step_tokens=128, max_vit_batch=1)
w = W.synthetic_engine_weights(cfg, dev, cap, seed=0)
eng.load_weights({k: v for k, v in w.items() if not k.startswith("vit.")})
sid = eng.stream_open()
eng.kv_fill_synthetic(sid, KV, seed=3)
skip = int(os.environ.get("SKIP_MERGE", "1"))
arr = (C.c_int32 * 1)(sid)
ab = C.c_double(0)
for it in (1, 3):
    rc = eng.lib.vlo_bench_attn(eng._h, 1, arr, 11, it, skip, C.byref(ab), eng._stream()); assert rc == 0, eng.lib.vlo_last_error()
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.lib.vlo_bench_attn(eng._h, 1, arr, 11, 4, skip, C.byref(ab), eng._stream()); e1.record(); torch.cuda.synchronize()
print(f"loop: {e0.elapsed_time(e1) * 1e3 / (4 * cfg.num_hidden_layers):.2f} us per launch (skip_merge={skip}), algo {ab.value / 1e6:.1f} MB")
n_cta = 16 * 8
buf = (C.c_longlong * (192 * n_cta))()
assert eng.lib.vlo_debug_attn_trace(buf, 192 * n_cta) == 0
a = np.array(buf, dtype=np.int64).reshape(n_cta, 3, 64)
st0 = a[:, 0, 0]
def col(role, idx):
    m = a[:, role, idx] > 0
    return (a[:, role, idx] - st0)[m]
def stat(name, x):
    if len(x): print(f"{name:34s} mean {x.mean():8.0f}  min {x.min():8.0f}  max {x.max():8.0f}")
print("cycles from each CTA's own start (clock64 is SM-local)")
stat("setup done", col(0, 1)); stat("after pdl wait", col(0, 2)); stat("q staged (row 0)", col(2, 0))
for j in range(6):
    stat(f"blk{j} k_full seen by mma", col(1, 3 * j)); stat(f"blk{j} qk issued", col(1, 3 * j + 1)); stat(f"blk{j} pv issued", col(1, 3 * j + 2))
for k in range(3):
    stat(f"group A blk{2*k} s_full seen", col(2, 4 + 4 * k)); stat(f"group A blk{2*k} s read", col(2, 5 + 4 * k))
    stat(f"group A blk{2*k} p_empty ok", col(2, 6 + 4 * k)); stat(f"group A blk{2*k} P staged", col(2, 7 + 4 * k))
stat("o_done (both groups)", col(2, 1)); stat("partials written", col(2, 2))
for cta in (0, 77):
    print("CTA", cta, "producer K/V issue:", [int(x - st0[cta]) if x > 0 else -1 for x in a[cta, 0, 4:18]])
