"""Dev tool: KV-append attention vs torch for many shapes / seeds, reporting WHERE the error sits."""
import ctypes as C, math, os, pathlib, sys
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa
from videollm_online_b200 import _lib
lib = _lib.load()
dev = "cuda"

def ref(q, k, v, kv_len):
    n_tok, H, D = q.shape
    G = H // k.shape[0]
    kk = k[:, :kv_len].float().repeat_interleave(G, 0)
    vv = v[:, :kv_len].float().repeat_interleave(G, 0)
    s = q.float().permute(1, 0, 2) @ kk.transpose(1, 2) / math.sqrt(D)
    pos = torch.arange(kv_len - n_tok, kv_len, device=dev)[:, None]
    s = s.masked_fill(~(torch.arange(kv_len, device=dev)[None, :] <= pos)[None], float("-inf"))
    return (torch.softmax(s, -1) @ vv).permute(1, 0, 2).reshape(n_tok, H, D)

shapes = [(17, 32, 8, 5000, 5056), (1, 32, 8, 13211, 13312), (1, 32, 8, 66011, 66176), (1, 32, 8, 5000, 5056), (5, 32, 8, 3000, 3072),
          (8, 32, 8, 9000, 9088), (16, 32, 8, 4000, 4096), (32, 32, 8, 6000, 6016), (11, 32, 8, 13211, 13312), (20, 32, 8, 12000, 12032)]
bad = 0
for (n_tok, H, Hk, kv_len, stride) in shapes:
    for seed in (0, kv_len, 7):
        torch.manual_seed(seed)
        D = 128
        q = torch.randn(n_tok, H, D, device=dev).bfloat16()
        k = torch.randn(Hk, stride, D, device=dev).bfloat16()
        v = torch.randn(Hk, stride, D, device=dev).bfloat16()
        ws = torch.empty(lib.vlo_op_attn_ws_bytes(n_tok, H, D, kv_len), device=dev, dtype=torch.uint8)
        out = torch.empty(n_tok, H * D, device=dev, dtype=torch.bfloat16)
        rc = lib.vlo_op_attn_kvappend(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), ws.data_ptr(), n_tok, H, Hk, D, kv_len, stride,
                                      torch.cuda.current_stream().cuda_stream)
        assert rc == 0, lib.vlo_last_error()
        torch.cuda.synchronize()
        o = out.float().view(n_tok, H, D)
        r = ref(q, k, v, kv_len)
        err = (o - r).abs()
        nan = ~torch.isfinite(o)
        e2 = torch.where(nan, torch.full_like(err, 1e9), err).amax(-1)     # [tok, head]
        worst = e2.max().item()
        ok = worst < 2e-2
        msg = f"n_tok={n_tok} kv={kv_len} seed={seed}: max_err={worst:.3e} {'OK' if ok else 'FAIL'}"
        if not ok:
            bad += 1
            idx = (e2 > 2e-2).nonzero()
            msg += f" bad(tok,head) n={idx.shape[0]} first={idx[:12].tolist()} nan_elems={int(nan.sum())}"
            t0, h0 = idx[0].tolist()
            msg += f" | row[{t0},{h0}] out[:4]={o[t0, h0, :4].tolist()} ref[:4]={r[t0, h0, :4].tolist()} out/ref={(o[t0,h0,:4]/r[t0,h0,:4]).tolist()}"
        print(msg, flush=True)
print("BAD" if bad else "ALL OK", bad)
