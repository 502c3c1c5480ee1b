"""Dev tool: micro-loop timing of the KV-append attention (vlo_op_attn_bench) for the kernel generation selected by
VLO_ATTN / VLO_ATTN_BLK in the environment; one JSON line per shape.  Run under gpurun, once per env setting."""
import ctypes as C, json, os, pathlib, sys
import torch

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import _lib

lib = _lib.load()
dev = "cuda"
peak = 6486.8
mp = pathlib.Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
if mp.exists():
    peak = float(json.loads(mp.read_text())["hbm_gbs"])
tag = os.environ.get("VLO_ATTN", "3") + "/" + os.environ.get("VLO_ATTN_BLK", "128")
H, Hk, D, L = 32, 8, 128, 32
shapes = [(11, 12011), (1, 12011), (11, 6011), (11, 66011)] if len(sys.argv) < 2 else [tuple(map(int, a.split(","))) for a in sys.argv[1:]]
torch.manual_seed(0)
out_lines = []
for n_tok, kv_len in shapes:
    stride = (kv_len + 127) // 128 * 128 + 128
    layers = L if kv_len < 40000 else 8
    k = torch.randn(layers, Hk * stride, D, device=dev, dtype=torch.bfloat16)
    v = torch.randn(layers, Hk * stride, D, device=dev, dtype=torch.bfloat16)
    q = torch.randn(n_tok, H, D, device=dev, dtype=torch.bfloat16)
    out = torch.empty(n_tok, H * D, device=dev, dtype=torch.bfloat16)
    ws = torch.empty(lib.vlo_op_attn_ws_bytes(n_tok, H, D, kv_len), device=dev, dtype=torch.uint8)
    ab = C.c_double(0)
    st = torch.cuda.current_stream().cuda_stream

    def loop(iters, skip):
        rc = lib.vlo_op_attn_bench(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), ws.data_ptr(), n_tok, H, Hk, D, kv_len,
                                   stride, layers, Hk * stride, iters, skip, C.byref(ab), st)
        assert rc == 0, lib.vlo_last_error()

    res = {"kernel": tag, "n_tok": n_tok, "kv_len": kv_len}
    for skip in (0, 1):
        loop(2, skip)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); loop(4, skip); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / (4 * layers))
        key = "main_only" if skip else "incl_merge"
        res[key + "_us"] = round(best * 1e3, 2)
        res[key + "_frac"] = round(ab.value / 1e9 / (best / 1e3) / peak, 3)
    res["algo_mb"] = round(ab.value / 1e6, 2)
    print(json.dumps(res), flush=True)
    out_lines.append(res)
    del k, v
pathlib.Path("gpurun_out").mkdir(exist_ok=True)
with open("gpurun_out/attn_ab.jsonl", "a") as f:
    for r in out_lines:
        f.write(json.dumps(r) + "\n")
