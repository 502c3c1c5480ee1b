"""Dev tool: the ViT trunk's GEMM sequence (q|k|v -> out_proj(+resid) -> fc1(gelu) -> fc2(+resid)) x layers through vlo_op_gemm2
back to back on one stream (no attention / LayerNorm in between): isolates gemm2 <-> gemm2 interplay under PDL."""
import pathlib, sys, time
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import _lib
lib = _lib.load()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
L = int(sys.argv[2]) if len(sys.argv) > 2 else 24
R, C_, M_ = 576 * B, 1024, 4096
torch.manual_seed(0)
xn = (torch.randn(R, C_, device=dev) * 0.5).half()
att = (torch.randn(R, C_, device=dev) * 0.5).half()
qkv = torch.zeros(R, 3 * C_, device=dev, dtype=torch.float16)
mlp = torch.zeros(R, M_, device=dev, dtype=torch.float16)
h = torch.zeros(R, C_, device=dev, dtype=torch.float32)
W = {n: (torch.randn(*s, device=dev) * 0.03).half() for n, s in (("qkv", (3 * C_, C_)), ("out", (C_, C_)), ("fc1", (M_, C_)), ("fc2", (C_, M_)))}
Bs = {n: (torch.randn(w.shape[0], device=dev) * 0.1).half().float() for n, w in W.items()}
st = torch.cuda.current_stream().cuda_stream
mt = (R + 255) // 256
def bn(n_out): return 256 if (n_out % 256 == 0 and mt * (n_out // 256) >= 60) else 128
def g(x, name, out, act, epi):
    w = W[name]
    rc = lib.vlo_op_gemm2(x.data_ptr(), R, w.data_ptr(), w.shape[0], w.shape[1], out.data_ptr(), w.shape[0], Bs[name].data_ptr(), act, epi, bn(w.shape[0]), st)
    assert rc == 0, lib.vlo_last_error()
t0 = time.time()
for l in range(L):
    g(xn, "qkv", qkv, 0, 0)
    g(att, "out", h, 0, 1)
    g(xn, "fc1", mlp, 1, 0)
    g(mlp, "fc2", h, 0, 1)
torch.cuda.synchronize()
ref = (att.float() @ W["out"].float().t() + Bs["out"]).half().float() + (mlp.float() @ W["fc2"].float().t() + Bs["fc2"]).half().float()
err = (h / L - ref).abs().max().item()
print(f"chain B={B} L={L}: ok in {time.time() - t0:.2f} s, resid err {err:.4f}", flush=True)
