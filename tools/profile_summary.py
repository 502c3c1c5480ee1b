"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv profiles/rNN_launches.md "note"
  python tools/profile_summary.py metrics  gpurun_out/prof.ncu-rep profiles/rNN_name_metrics.csv
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max", "smsp__inst_executed.sum"]


def launches(src, dst, note=""):
    lines = [l for l in open(src) if l.startswith('"')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("vlo::", "")
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        a = agg.setdefault((name, row.get("Grid Size", "")), [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    n_attn = sum(a[0] for k, a in agg.items() if "attn_tc" in k[0] or "attn_kvappend" in k[0])
    steps = max(1.0, n_attn / 32.0)
    out = [f"# ncu launch list summary ({src})", "", note, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.",
           f"captured {tot:.0f} us over ~{steps:.1f} frame steps ({tot / steps:.0f} us/step serialised)", "",
           "| us / step | share | launches / step | avg us | kernel | grid |", "|---:|---:|---:|---:|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {a[1] / steps:.1f} | {100 * a[1] / tot:.1f}% | {a[0] / steps:.1f} | {a[1] / a[0]:.2f} | `{k[0][:70]}` | {k[1]} |")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out[:20]))


def metrics(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for row in r[2:]:
            w.writerow([row[i] for i in idx])
    print(open(dst).read()[:1500])


def traffic(dst, *pairs):
    """pairs: key=report.ncu-rep[:name-substring] -> profiles/ncu_traffic.json with mean dram bytes per launch"""
    import json
    out = {}
    for pr in pairs:
        key, rest = pr.split("=", 1)
        rep, _, sub = rest.partition(":")
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(raw.splitlines()))
        hdr, units = r[0], r[1]
        ir, iw, it, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum"), hdr.index("Kernel Name")
        def to_b(v, u):
            v = float(v)
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rows = [x for x in r[2:] if sub in x[ik]]
        tot = [to_b(x[ir], units[ir]) + to_b(x[iw], units[iw]) for x in rows]
        dur = [float(x[it]) for x in rows]
        out[key] = {"dram_bytes_per_launch": sum(tot) / max(1, len(tot)), "launches": len(tot), "avg_duration_" + units[it]: sum(dur) / max(1, len(dur)),
                    "source": rep.split("/")[-1] + " (ncu --set full --clock-control none, cold cache)"}
    try:   # stamp the capture with the commit it was taken at (bench.py reports it as traffic_commit)
        out["commit"] = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:
        pass
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], *sys.argv[3:])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        metrics(sys.argv[2], sys.argv[3])
