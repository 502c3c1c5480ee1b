"""Dev tool: the 2-CTA ViT GEMM (vlo_op_gemm2) against torch on the ViT shapes: correctness (fp32 reference of the same
fp16 operands) and device time next to torch.matmul (cuBLAS) as the library yardstick.  Run under gpurun with a timeout."""
import json, pathlib, sys
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import vlo_bootstrap  # noqa: F401
from videollm_online_b200 import _lib
lib = _lib.load()
dev = "cuda"
torch.manual_seed(0)
st = torch.cuda.current_stream().cuda_stream
ok_all = True


def run(rows_x, rows_w, k, epi, act, bn, check=True, time_it=True):
    global ok_all
    x = (torch.randn(rows_x, k, device=dev) * 0.5).half()
    w = (torch.randn(rows_w, k, device=dev) * 0.03).half()
    bias = (torch.randn(rows_w, device=dev) * 0.1).half().float()
    if epi == 0:
        out = torch.zeros(rows_x, rows_w, device=dev, dtype=torch.float16)
    else:
        out = torch.randn(rows_x, rows_w, device=dev, dtype=torch.float32)
        h0 = out.clone()

    def call():
        rc = lib.vlo_op_gemm2(x.data_ptr(), rows_x, w.data_ptr(), rows_w, k, out.data_ptr(), rows_w, bias.data_ptr(), act, epi, bn, st)
        assert rc == 0, lib.vlo_last_error()
    call()
    torch.cuda.synchronize()
    rec = {"rows_x": rows_x, "rows_w": rows_w, "k": k, "epi": epi, "act": act, "bn": bn}
    if check:
        y = (x.float() @ w.float().t() + bias).half().float()
        if epi == 0:
            if act == 1:
                y = torch.nn.functional.gelu(y, approximate="tanh").half().float()
            err = (out.float() - y).abs().max().item()
            tol = 2e-2 if act == 0 else 3e-2
        else:
            err = (out - (h0 + y)).abs().max().item()
            tol = 2e-2
        rec["max_err"] = err
        rec["ok"] = bool(err < tol)
        ok_all = ok_all and rec["ok"]
    if time_it:
        if epi == 1:
            out.zero_()
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            call()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        rec["us"] = round(us, 2)
        rec["tflops"] = round(2.0 * rows_x * rows_w * k / us / 1e6, 1)
        o2 = torch.empty(rows_x, rows_w, device=dev, dtype=torch.float16)
        for _ in range(3):
            torch.matmul(x, w.t(), out=o2)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            torch.matmul(x, w.t(), out=o2)
        e1.record()
        torch.cuda.synchronize()
        rec["cublas_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 2)
    print(json.dumps(rec), flush=True)


# small / ragged shapes first (a protocol bug shows up here in milliseconds)
run(256, 256, 64, 0, 0, 256, time_it=False)
run(256, 128, 128, 0, 0, 128, time_it=False)
run(200, 256, 256, 0, 0, 256, time_it=False)       # ragged rows inside the peer half
run(700, 512, 1024, 0, 1, 256, time_it=False)      # peer half fully out of range on the last token tile; several tiles per pair
run(1728, 1024, 1024, 1, 0, 128, time_it=False)    # B = 3: residual epilogue
run(2880, 1024, 1024, 1, 0, 128, time_it=False)    # B = 5: the peer CTA's half of the last token tile starts past the matrix
run(2880, 3072, 1024, 0, 0, 256, time_it=False)
for B in (4, 8):
    R = 576 * B
    run(R, 3072, 1024, 0, 0, 256)     # q|k|v
    run(R, 1024, 1024, 1, 0, 128)     # out_proj (+ residual)
    run(R, 4096, 1024, 0, 1, 256)     # fc1 + gelu
    run(R, 1024, 4096, 1, 0, 128)     # fc2 (+ residual)
print("ALL OK" if ok_all else "FAILED")
sys.exit(0 if ok_all else 1)
