"""Digest the ncu captures of scripts/r02_call10.sh (gpurun_out/r02_call10/) into profiles/ (tracked):
profiles/r02_ncu_full.md, profiles/r02_launches.{md,csv}, profiles/ncu_summary.json (read by bench.py for roofline.traffic).
Runs in the build container:  python scripts/summarize_profiles_r02.py [gpurun_out/r02_call10]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "r02_call10")
OUT = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "DRAM busy"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active"),
    ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcgen05 issue pipe"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "wait", "barrier", "short_scoreboard", "branch_resolving", "not_selected", "math_pipe_throttle",
               "no_instruction", "mio_throttle", "lg_throttle", "sleeping", "membar", "dispatch_stall", "selected"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = {"kernel": row[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            d[h] = (row[i], units[i])
        res.append(d)
    return res


def fnum(d, k):
    try:
        return float(d[k][0].replace(",", ""))
    except Exception:
        return None


def section(title, what, rep, notes):
    path = os.path.join(SRC, rep)
    if not os.path.exists(path):
        return ["## %s" % title, "", "(capture missing: %s)" % rep, ""], None
    d = raw(path)[0]
    md = ["## %s" % title, "", "`%s` -- %s" % (d["kernel"], what), "", "| metric | value |", "|---|---|"]
    for k, name in KEYS:
        if k in d and d[k][0] != "":
            md.append("| %s (`%s`) | %s %s |" % (name, k, d[k][0], d[k][1]))
    st = [(fnum(d, STALLS % s) or 0.0, s) for s in STALL_NAMES if (STALLS % s) in d]
    st.sort(reverse=True)
    md.append("| warp stall cycles per issued instruction | %s |" % ", ".join("%s %.2f" % (s, v) for v, s in st[:7]))
    md += [""] + notes + [""]
    return md, d


def launches():
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += float(row["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    md = ["# Launch list of 2 bench steps (config 2, N=1, Z-free default) -- `ncu --metrics gpu__time_duration.sum --clock-control none`",
          "", "Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.  (Command: `scripts/r02_call20.sh`, final build of round 2.)", "",
          "| kernel | launches | total ms | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, a in agg.items():
        md.append("| `%s` | %d | %.3f | %.1f | %.1f %% |" % (k[:100], a[0], a[1] / 1e6, a[1] / a[0] / 1e3, 100 * a[1] / tot))
    open(os.path.join(OUT, "r02_launches.md"), "w").write("\n".join(md) + "\n")
    open(os.path.join(OUT, "r02_launches.csv"), "w").write("".join(lines))
    return {k: {"launches": a[0], "total_ns": a[1]} for k, a in agg.items()}


md = ["# ncu --set full, round 2 (one B200, `--clock-control none`, `scripts/r02_call10.sh`)", "",
      "One launch of each dominant kernel, captured inside `bench.py` after 3 warm-up steps.  Times under the profiler are "
      "cold-cache and serialised -- the timing claims are `bench.py`'s CUDA events; these captures explain them.", ""]
summary = {"round": "r02", "mindist_tc_dram_bytes_per_launch": {}}
for title, what, rep, key, notes in [
    ("Symmetric distance kernel (config 2)", "tcgen05 `cta_group::2` GEMM, per-bank-image row-min + column-min in the epilogue; 100 images, every unordered pair once",
     "r02_mindist_sym.ncu-rep", "config2",
     ["Executed 24.92 TFLOP per launch.  The kernel is tensor-bound and power-capped (SM clock above): DRAM traffic per launch is the "
      "`roofline.traffic` of the bench line (compulsory: 642 MB of operands + 62 MB of minima)."]),
    ("All-pairs distance kernel (config 3, supervised)", "same kernel without the unit list: 100 query images x 200 bank images",
     "r02_mindist_allpairs.ncu-rep", "config3",
     ["Executed = algorithmic = 100.7 TFLOP per launch (VERDICT r01 item 6: this form was 12 % behind the symmetric one in round 1; "
      "it is now at the same rate -- bench lines `profiles/r02_call10/bench_config3_steps_5.json`, `r02_bench_n1_config2_nosymmetry.json`)."]),
    ("Exact refine kernel (config 2, precision f16r)", "fp32 re-evaluation of the selected (query row, arg-min bank row) pairs",
     "r02_refine.ncu-rep", None,
     ["7.76 M pairs x one gathered 8 KB fp16 bank row: the kernel is bound by L2 -> SM traffic (see the L2 -> SM bytes row)."]),
    ("Lean fused embed kernel, FIRST version (call 10; config 2, Z-free default)", "statistics + both layers + fp16 operands + norms in one persistent launch",
     "r02_embed_lean.ncu-rep", None,
     ["Algorithmic bytes per launch: 482 MB of maps + 642 MB of operands + 0.3 MB of norms = 1 124 MB.  DRAM traffic above is "
      "read + written; the excess over 1 124 MB is maps evicted from L2 between their statistics read and their embed reads.  "
      "The source page of this capture showed the single producer thread as the bottleneck (188 instructions per ring slot, the "
      "consumers waiting for data 23 % of their samples), the per-item re-reduction of the statistics behind two CTA barriers (11 %) "
      "and the release fences (6 %) -- what calls 11-14 removed."]),
    ("Lean fused embed kernel, FINAL version (call 14)", "hoisted + hardware-suspended producer, one bulk copy per ring slot (column walk), "
     "last-arriver statistics with flag-free 64-bit slots, evict_last map loads, streaming operand stores, FHFMA norms",
     "r02_embed_lean_v5.ncu-rep", None,
     ["1 124 MB algorithmic in 254 us under the profiler (0.250-0.256 ms by CUDA events, isolated) = 0.67-0.685 of the 6 553 GB/s copy peak; "
      "real DRAM traffic 1 248 MB = 4.9 TB/s = 0.75 of the peak.  No stall reason dominates any more (wait 26 %, long scoreboard 18 %, "
      "selected 16 %); the kernel time scales with the SM clock (0.36-0.38 ms inside the bench step at 1.25-1.29 GHz)."]),
]:
    sec, d = section(title, what, rep, notes)
    md += sec
    if d is not None:
        rd, wr = fnum(d, "dram__bytes_read.sum"), fnum(d, "dram__bytes_write.sum")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = (rd or 0) * scale.get(d["dram__bytes_read.sum"][1], 1.0) + (wr or 0) * scale.get(d["dram__bytes_write.sum"][1], 1.0)
        summary.setdefault("kernels", {})[rep.replace(".ncu-rep", "")] = {
            "kernel": d["kernel"], "dram_bytes_per_launch": tot,
            "time_us": fnum(d, "gpu__time_duration.sum") * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1.0)}
        if key:
            summary["mindist_tc_dram_bytes_per_launch"][key] = tot
open(os.path.join(OUT, "r02_ncu_full.md"), "w").write("\n".join(md) + "\n")
summary["launches"] = launches()
if summary["launches"] is None:      # the launch list of call 10 is not in this directory: keep the digested one
    try:
        summary["launches"] = json.load(open(os.path.join(OUT, "ncu_summary.json"))).get("launches")
    except Exception:
        pass
json.dump(summary, open(os.path.join(OUT, "ncu_summary.json"), "w"), indent=1)
print(open(os.path.join(OUT, "r02_ncu_full.md")).read())
