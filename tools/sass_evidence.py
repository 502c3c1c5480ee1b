"""Counts the SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, mbarrier -> SYNCS, mma.sync -> HMMA, ldmatrix -> LDSM).  Runs on the build
host (no GPU): python tools/sass_evidence.py > profiles/<round>_sass_evidence.md"""
import collections, pathlib, re, subprocess, sys

so = pathlib.Path(__file__).resolve().parents[1] / "videollm-online_b200" / "libvlo_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
keys = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "UCGABAR_ARV", "ACQBULK", "HMMA", "LDSM", "MUFU"]
names = {"UTCHMMA": "tcgen05.mma (incl. .2CTA = cta_group::2)", "UTMALDG": "TMA tensor load", "UTMASTG": "TMA tensor store",
         "UTMAREDG": "TMA tensor reduce-add", "UCGABAR_ARV": "barrier.cluster.arrive", "LDTM": "tcgen05.ld", "STTM": "tcgen05.st",
         "UTCBAR": "tcgen05.commit", "SYNCS": "mbarrier ops", "UTCATOMSWS": "TMEM alloc/dealloc",
         "ACQBULK": "griddepcontrol.wait (PDL)", "HMMA": "mma.sync", "LDSM": "ldmatrix", "MUFU": "SFU (exp2, rsqrt, ...)"}
rows = []
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    mangled = f.split("\n", 1)[0].strip()
    ops = collections.Counter(re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", f))
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*$", "", dem).replace("void ", "").replace("vlo::", "")
    rows.append((short, ops, sum(ops.values())))
want = sys.argv[1:] or ["gemm_ws_kernel<1, 16, 6>", "gemm_ws_kernel<1, 96, 7>", "gemm_wsf_kernel<16, 6, 2>", "gemm_ws_kernel<0, 64, 3>",
                        "gemm2_kernel<256, 5, 0>", "gemm2_kernel<128, 7, 1>", "gemm_tn_kernel<0, 128, false, 3>",
                        "attn_tc_kernel", "attn_tc2_kernel<128>", "attn_kvappend_kernel", "attn_merge_kernel",
                        "vit_attn_tc2_kernel", "vit_attn_tc_kernel", "vit_attn_kernel", "resid_rmsnorm_kernel",
                        "qkv_rope_append_kernel", "swiglu_kernel", "decision_kernel"]
print("# SASS evidence (static instruction counts, `cuobjdump -sass libvlo_b200.so`, sm_100a)\n")
print("| mnemonic | PTX / meaning |\n|---|---|")
for k in keys:
    print(f"| `{k}` | {names[k]} |")
print("\n| kernel | SASS instructions | " + " | ".join(keys) + " |")
print("|---|---:|" + "---:|" * len(keys))
for w in want:
    for short, ops, n in rows:
        if short == w:
            print(f"| `{short}` | {n} | " + " | ".join(str(ops.get(k, 0)) for k in keys) + " |")
print("\nEvery GEMM (`gemm_ws_kernel`, `gemm_wsf_kernel`, `gemm2_kernel`, `gemm_tn_kernel`), the decoder attention (`attn_tc_kernel` below"
      " 24k keys, `attn_tc2_kernel` above) and the ViT attention of the tensor-bound batches (`vit_attn_tc2_kernel`, >= 3 frames) issue"
      " tcgen05.mma from TMA-fed shared memory with TMEM accumulators; `gemm2_kernel` is the 2-CTA form (UTCHMMA.2CTA, UTMALDG.2CTA,"
      " multicast commits, TMA store / reduce epilogue).  `attn_kvappend_kernel` (mma.sync) is the A/B baseline (`VLO_ATTN=1`);"
      " `vit_attn_kernel` (mma.sync) serves 1-2 frame batches, where a launch is latency-bound and measured faster (DESIGN.md section 3).")
