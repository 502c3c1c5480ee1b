"""Dev tool: exercise vlo_op_gemm on the GPU against torch.matmul (run under gpurun)."""
import ctypes as C, pathlib, sys, json, time
import torch

lib = C.CDLL(str(pathlib.Path(__file__).resolve().parents[1] / "videollm-online_b200" / "libvlo_b200.so"))
lib.vlo_last_error.restype = C.c_char_p
P, I, LL = C.c_void_p, C.c_int, C.c_longlong
lib.vlo_op_gemm.argtypes = [I, I, I, I, P, I, P, I, I, P, I, P, P, I, I, LL, I, P]
lib.vlo_op_gemm.restype = I

def gemm(fmt, swap, epi, act, a, b, out, ld_out, bias=None, pos=None, pos_rows=0, splits=1, split_stride=0, bn=0):
    rc = lib.vlo_op_gemm(fmt, swap, epi, act, a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1],
                         out.data_ptr(), ld_out, bias.data_ptr() if bias is not None else None,
                         pos.data_ptr() if pos is not None else None, pos_rows, splits, split_stride, bn,
                         torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        raise RuntimeError(lib.vlo_last_error().decode())

res = []
dev = "cuda"
torch.manual_seed(0)
def report(name, got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(scale, 1e-6)
    res.append(dict(name=name, err=err, scale=scale, ok=bool(ok)))
    print(f"{name:60s} err={err:.4e} scale={scale:.3e} {'OK' if ok else 'FAIL'}", flush=True)

# 1. swap-AB bf16 STORE16, various token counts / K / N
for (T, N, K) in [(11, 256, 128), (16, 384, 512), (1, 128, 64), (33, 1000, 1024), (88, 4096, 4096), (130, 512, 256)]:
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.randn(T, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.zeros(T, N, device=dev, dtype=torch.bfloat16)
    gemm(1, 1, 1, 0, w, x, out, N, bias=bias)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t() + bias).bfloat16()
    report(f"swap bf16 store16 T={T} N={N} K={K}", out, ref, 1e-2)

# 2. split-K partial
for (T, N, K, S) in [(11, 512, 1024, 4), (11, 4096, 14336, 7), (40, 640, 4096, 3)]:
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    x = torch.randn(T, K, device=dev).bfloat16()
    ws = torch.full((S, T, N), float('nan'), device=dev)
    gemm(1, 1, 0, 0, w, x, ws, N, splits=S, split_stride=T * N)
    torch.cuda.synchronize()
    ref = x.float() @ w.float().t()
    report(f"swap bf16 partial T={T} N={N} K={K} S={S}", ws.sum(0), ref, 2e-3)

# 3. non-swap fp16 (ViT): store16 + gelu, resid32, patch32
for (M, N, K, bn) in [(576, 1024, 1024, 128), (576, 3072, 1024, 128), (300, 200, 768, 64), (1152, 4096, 1024, 128)]:
    a = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * 0.05).half()
    bias = torch.randn(N, device=dev)
    ld = (N + 7) // 8 * 8
    out = torch.zeros(M, ld, device=dev, dtype=torch.float16)
    gemm(0, 0, 1, 1, a, w, out, ld, bias=bias, bn=bn)
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu((a.float() @ w.float().t() + bias).half().float(), approximate='tanh').half()
    report(f"noswap fp16 store16+gelu_tanh M={M} N={N} K={K} bn={bn}", out[:, :N], ref, 1e-2)
    h = torch.randn(M, ld, device=dev)
    h0 = h.clone()
    gemm(0, 0, 2, 0, a, w, h, ld, bias=bias, bn=bn)
    torch.cuda.synchronize()
    ref = h0[:, :N] + (a.float() @ w.float().t() + bias).half().float()
    report(f"noswap fp16 resid32 M={M} N={N} K={K} bn={bn}", h[:, :N], ref, 1e-2)
    pos_rows = 36
    pos = torch.randn(pos_rows, ld, device=dev)
    o32 = torch.zeros(M, ld, device=dev)
    gemm(0, 0, 3, 0, a, w, o32, ld, bias=bias, pos=pos, pos_rows=pos_rows, bn=bn)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float().t() + bias).half().float() + pos[torch.arange(M, device=dev) % pos_rows][:, :N]
    report(f"noswap fp16 patch32 M={M} N={N} K={K} bn={bn}", o32[:, :N], ref, 1e-2)

# 4. gelu erf python form, bf16 swap
T, N, K = 20, 512, 256
w = (torch.randn(N, K, device=dev) * 0.1).bfloat16(); x = torch.randn(T, K, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
out = torch.zeros(T, N, device=dev, dtype=torch.bfloat16)
gemm(1, 1, 1, 2, w, x, out, N, bias=bias)
torch.cuda.synchronize()
y = (x.float() @ w.float().t() + bias).bfloat16()
ref = y * 0.5 * (1.0 + torch.erf(y / 1.4142135623730951))
report("swap bf16 store16 gelu_erf_py", out, ref, 1e-2)

# 5. bandwidth: gate/up-sized weight streaming, T=11
T, N, K = 11, 28672, 4096
w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); x = torch.randn(T, K, device=dev).bfloat16()
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
for S in (1, 2, 3):
    ws = torch.empty(S, T, N, device=dev)
    for _ in range(3):
        gemm(1, 1, 0, 0, w, x, ws, N, splits=S, split_stride=T * N)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm(1, 1, 0, 0, w, x, ws, N, splits=S, split_stride=T * N); e1.record()
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    t = sorted(times)[len(times) // 2]
    gbs = N * K * 2 / t / 1e6
    print(f"bandwidth gate_up T=11 splits={S}: {t*1e3:.1f} us  {gbs:.0f} GB/s", flush=True)
    res.append(dict(name=f"bw_gateup_S{S}", us=t * 1e3, gbs=gbs, ok=True))
# down proj
T, N, K = 11, 4096, 14336
w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); x = torch.randn(T, K, device=dev).bfloat16()
for S in (4, 5, 8, 9):
    ws = torch.empty(S, T, N, device=dev)
    for _ in range(3):
        gemm(1, 1, 0, 0, w, x, ws, N, splits=S, split_stride=T * N)
    times = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gemm(1, 1, 0, 0, w, x, ws, N, splits=S, split_stride=T * N); e1.record()
        torch.cuda.synchronize(); times.append(e0.elapsed_time(e1))
    t = sorted(times)[len(times) // 2]
    gbs = N * K * 2 / t / 1e6
    print(f"bandwidth down T=11 splits={S}: {t*1e3:.1f} us  {gbs:.0f} GB/s", flush=True)
    res.append(dict(name=f"bw_down_S{S}", us=t * 1e3, gbs=gbs, ok=True))
# ViT compute: fc1 at batch 8
M, N, K = 576 * 8, 4096, 1024
a = torch.randn(M, K, device=dev).half(); w = (torch.randn(N, K, device=dev) * 0.05).half(); bias = torch.randn(N, device=dev)
out = torch.zeros(M, N, device=dev, dtype=torch.float16)
for _ in range(3):
    gemm(0, 0, 1, 1, a, w, out, N, bias=bias, bn=128)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gemm(0, 0, 1, 1, a, w, out, N, bias=bias, bn=128)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 20
print(f"vit fc1 B=8: {t*1e3:.1f} us {2*M*N*K/t/1e9:.0f} TFLOP/s", flush=True)
res.append(dict(name="vit_fc1_b8", us=t * 1e3, tflops=2 * M * N * K / t / 1e9, ok=True))

pathlib.Path("gpurun_out").mkdir(exist_ok=True)
json.dump(res, open("gpurun_out/gemm_check.json", "w"), indent=1)
bad = [r for r in res if not r["ok"]]
print("FAILED:" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)
