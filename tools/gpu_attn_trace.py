"""Dev tool: per-CTA timeline of the tcgen05 attention kernel (VLO_ATTN_TRACE=1), run under gpurun."""
import ctypes as C, os, pathlib, sys
os.environ["VLO_ATTN_TRACE"] = "1"
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap
from videollm_online_b200 import _lib
lib = _lib.load()
p = lambda t: C.c_void_p(t.data_ptr())
dev = "cuda"
n_tok, H, Hk, D, kv_len = 11, 32, 8, 128, 12100
stride = (kv_len + 127) // 128 * 128
torch.manual_seed(0)
q = torch.randn(n_tok, H, D, device=dev).bfloat16()
k = torch.randn(Hk, stride, D, device=dev).bfloat16()
v = torch.randn(Hk, stride, D, device=dev).bfloat16()
ws = torch.empty(lib.vlo_op_attn_ws_bytes(n_tok, H, D, kv_len), device=dev, dtype=torch.uint8)
out = torch.empty(n_tok, H * D, device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for it in range(3):
    flush.zero_(); torch.cuda.synchronize()
    rc = lib.vlo_op_attn_kvappend(p(q), p(k), p(v), p(out), p(ws), n_tok, H, Hk, D, kv_len, stride, st)
    assert rc == 0, lib.vlo_last_error()
torch.cuda.synchronize()
n_cta = 16 * 8
buf = (C.c_longlong * (192 * n_cta))()
assert lib.vlo_debug_attn_trace(buf, 192 * n_cta) == 0
import numpy as np
a = np.array(buf, dtype=np.int64).reshape(n_cta, 3, 64)
# clock64 is an SM-local counter: only differences inside one CTA are meaningful
def rel(x, t0): return int(x - t0) if x > 0 else -1
for cta in (0, 5, 64, 127):
    pr, mm, sm = a[cta, 0], a[cta, 1], a[cta, 2]
    t0 = pr[0]
    print(f"--- CTA {cta}: setup_done {rel(pr[1], t0)} after_wait {rel(pr[2], t0)}")
    print("  producer K/V issue:", [rel(x, t0) for x in pr[4:4 + 16]])
    print("  mma (kfull, sempty, pv-issue) per block:", [rel(x, t0) for x in mm[:24]])
    print("  softmax q_ready:", rel(sm[0], t0), " per block (s_full, s_read, p_empty, p_done):", [rel(x, t0) for x in sm[4:4 + 32]])
    print("  epilogue: o_done", rel(sm[1], t0), "end", rel(sm[2], t0))
st0 = a[:, 0, 0]
def col(role, idx): return (a[:, role, idx] - st0)[a[:, role, idx] > 0]
def stat(name, x): print(f"{name:28s} mean {x.mean():8.0f}  min {x.min():8.0f}  max {x.max():8.0f}  (cycles from CTA start)")
stat("setup done", col(0, 1)); stat("after pdl wait", col(0, 2)); stat("q staged", col(2, 0))
for j in range(3):   # role-2 stamps are group A's (even blocks): index k = j // 2
    stat(f"blk{j} k_full seen by mma", col(1, 3 * j)); stat(f"blk{j} s_full seen by softmax", col(2, 4 + 4 * j))
    stat(f"blk{j} P staged", col(2, 7 + 4 * j)); stat(f"blk{j} pv issued", col(1, 3 * j + 2))
stat("o_done (last PV)", col(2, 1)); stat("partials written", col(2, 2))

