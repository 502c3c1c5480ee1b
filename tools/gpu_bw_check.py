"""Dev tool: bandwidth of the persistent stream-K GEMM on the decoder shapes (run under gpurun)."""
import ctypes as C, pathlib, sys, json
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap
from videollm_online_b200 import _lib
lib = _lib.load()
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
res = {}
for name, (N, K) in {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336), "lm_head": (128256, 4096)}.items():
    for T in (11, 88):
        w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); x = torch.randn(T, K, device=dev).bfloat16()
        mode = 1 if name == "lm_head" else 0
        planes = C.c_int(1)
        lib.vlo_op_gemm_ws(1, mode, p(w), N, p(x), T, K, None, N, T * N, None, 0, 0, 0, C.byref(planes), st())
        out = torch.zeros(planes.value, T, N, device=dev) if mode == 0 else torch.zeros(T, N, device=dev, dtype=torch.bfloat16)
        def run():
            rc = lib.vlo_op_gemm_ws(1, mode, p(w), N, p(x), T, K, p(out), N, T * N, None, 0, 0, 0, C.byref(planes), st())
            assert rc == 0, lib.vlo_last_error()
        for _ in range(3): run()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        print(f"{name:8s} T={T:3d} planes={planes.value}: {t*1e3:7.1f} us  {N*K*2/t/1e6:6.0f} GB/s", flush=True)
        res[f"{name}_T{T}"] = dict(us=t * 1e3, gbs=N * K * 2 / t / 1e6)
json.dump(res, open("gpurun_out/bw_check.json", "w"), indent=1)
