"""Dev tool: per-CTA timeline of the v3 KV-append attention inside the back-to-back micro-loop (vlo_op_attn_bench),
no engine needed.  VLO_ATTN_TRACE=1 makes the kernel stamp clock64 per role and %globaltimer at CTA start / end."""
import ctypes as C, os, pathlib, sys
os.environ["VLO_ATTN_TRACE"] = "1"
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import _lib
lib = _lib.load()
dev = "cuda"
n_tok, kv_len = int(os.environ.get("NTOK", "11")), int(os.environ.get("KV", "12011"))
skip = int(os.environ.get("SKIP_MERGE", "1"))
H, Hk, D, L = 32, 8, 128, 16
stride = (kv_len + 127) // 128 * 128 + 128
k = torch.randn(L, Hk * stride, D, device=dev, dtype=torch.bfloat16)
v = torch.randn(L, Hk * stride, D, device=dev, dtype=torch.bfloat16)
q = torch.randn(n_tok, H, D, device=dev, dtype=torch.bfloat16)
out = torch.empty(n_tok, H * D, device=dev, dtype=torch.bfloat16)
ws = torch.empty(lib.vlo_op_attn_ws_bytes(n_tok, H, D, kv_len), device=dev, dtype=torch.uint8)
ab = C.c_double(0)
st = torch.cuda.current_stream().cuda_stream
def loop(iters):
    rc = lib.vlo_op_attn_bench(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), ws.data_ptr(), n_tok, H, Hk, D, kv_len, stride, L,
                               Hk * stride, iters, skip, C.byref(ab), st)
    assert rc == 0, lib.vlo_last_error()
loop(2); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); loop(4); e1.record(); torch.cuda.synchronize()
print(f"loop: {e0.elapsed_time(e1) * 1e3 / (4 * L):.2f} us per launch (skip_merge={skip}, tracing on), algo {ab.value / 1e6:.1f} MB")
n_cta = 144
buf = (C.c_longlong * (192 * n_cta))()
assert lib.vlo_debug_attn_trace(buf, 192 * n_cta) == 0
a = np.array(buf, dtype=np.int64).reshape(n_cta, 3, 64)
st0 = a[:, 0, 0]
clk = 1.9   # cycles per ns (approx.)
def col(role, idx):
    m = a[:, role, idx] > 0
    return ((a[:, role, idx] - st0)[m]) / clk / 1e3
def stat(name, x):
    if len(x): print(f"{name:34s} mean {x.mean():7.2f}  min {x.min():7.2f}  max {x.max():7.2f} us   (n={len(x)})")
print("microseconds from each CTA's own start")
stat("setup done (tmem, barriers)", col(0, 1)); stat("grid dependency resolved", col(0, 2))
nb = 7
for j in range(nb):
    stat(f"blk{j} K landed (mma saw k_full)", col(1, 3 * j)); stat(f"blk{j} S issued", col(1, 3 * j + 1))
    stat(f"blk{j} s_full seen by softmax", col(2, 4 + 4 * j)); stat(f"blk{j} S row read", col(2, 5 + 4 * j)); stat(f"blk{j} P stored", col(2, 7 + 4 * j))
    stat(f"blk{j} PV issued", col(1, 3 * j + 2))
stat("o_done seen", col(2, 1)); stat("partials written", col(2, 2))
end = (a[:, 0, 62] - st0) / clk / 1e3
stat("CTA end", end[a[:, 0, 62] > 0])
g0, g1 = a[:, 0, 60], a[:, 0, 61]
m = (g0 > 0) & (g1 > 0)
print(f"globaltimer: first CTA start -> last CTA end {(g1[m].max() - g0[m].min()) / 1e3:.2f} us; CTA starts spread {(g0[m].max() - g0[m].min()) / 1e3:.2f} us; "
      f"CTA ends spread {(g1[m].max() - g1[m].min()) / 1e3:.2f} us; mean CTA lifetime {(g1[m] - g0[m]).mean() / 1e3:.2f} us")
for cta in (0, 17, 143):
    print("CTA", cta, "producer K/V issue (us):", [round(float(x - st0[cta]) / clk / 1e3, 2) if x > 0 else -1 for x in a[cta, 0, 4:18]])
