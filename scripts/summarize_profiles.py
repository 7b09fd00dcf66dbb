"""Digest gpurun_out/{launches,prof_*}_<round>.* into profiles/ (tracked).  Runs in the build container.
python scripts/summarize_profiles.py r01"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
os.makedirs(OUT, exist_ok=True)

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "gpc__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum",
]


def launches():
    path = os.path.join(GO, "launches_%s.csv" % R)
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    md = ["# Launch list of 2 bench steps (config 2, N=1) -- `ncu --metrics gpu__time_duration.sum --clock-control none`",
          "", "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.", "",
          "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, a in agg.items():
        md.append("| `%s` | %d | %.3f | %.1f | %.1f %% |" % (k[:90], a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    open(os.path.join(OUT, "%s_launches.md" % R), "w").write("\n".join(md) + "\n")
    # keep the raw csv too (small)
    open(os.path.join(OUT, "%s_launches.csv" % R), "w").write("".join(lines))
    return {k: {"launches": a[0], "total_ns": a[1]} for k, a in agg.items()}


def ncu_raw(name):
    rep = os.path.join(GO, "%s_%s.ncu-rep" % (name, R))
    if not os.path.exists(rep):
        return []
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
                try:
                    d[h] = (float(row[i].replace(",", "")), units[i])
                except ValueError:
                    pass
        res.append(d)
    return res


def to_bytes(v):
    val, unit = v
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return val * mult.get(unit, 1)


summary = {"round": R, "launches": launches()}
md = ["# ncu --set full captures (%s)" % R, ""]
for name, label in (("prof_mindist", "mindist_tc_kernel (symmetric form, config 2)"), ("prof_embed", "embed_tma_kernel (config 2, one launch per layer)")):
    for d in ncu_raw(name):
        md.append("## %s -- `%s`" % (label, d["kernel"][:80]))
        md.append("")
        md.append("| metric | value | unit |")
        md.append("|---|---:|---|")
        for k, v in d.items():
            if k == "kernel":
                continue
            if "issue_stalled" in k and v[0] < 0.2:
                continue
            md.append("| %s | %.4g | %s |" % (k, v[0], v[1]))
        md.append("")
        traffic = None
        if "dram__bytes_read.sum" in d:
            traffic = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
        summary.setdefault(name, []).append({"kernel": d["kernel"][:80], "dram_bytes_per_launch": traffic,
                                             "time_ms": d.get("gpu__time_duration.sum", (None,))[0]})
if "prof_mindist" in summary:
    summary["mindist_tc_dram_bytes_per_launch"] = {"config2": summary["prof_mindist"][0]["dram_bytes_per_launch"]}
open(os.path.join(OUT, "%s_ncu_full.md" % R), "w").write("\n".join(md) + "\n")
json.dump(summary, open(os.path.join(OUT, "ncu_summary.json"), "w"), indent=1)
bj = os.path.join(GO, "bench_%s.json" % R)
if os.path.exists(bj):
    for line in open(bj):
        if line.startswith("{"):
            json.dump(json.loads(line), open(os.path.join(OUT, "%s_bench.json" % R), "w"), indent=1)
print("profiles written to", OUT)
