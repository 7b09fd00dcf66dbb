#!/bin/bash
# Round 2, GPU call 20 (one B200): two-kernel multi-CTA unit builder -- tests, launch lists (f16 and f16r), ncu of the
# refine kernel, bench lines.
OUT=gpurun_out/r02_call20
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/pytest_gpu.log
for prec in auto f16r; do
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_$prec.csv python bench.py --workload config2 --precision $prec --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --cuda-profiler > $OUT/launches_$prec.log 2>&1; echo "launch list $prec rc=$?"
python - $OUT/launches_$prec.csv <<'PY'
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    a = agg.setdefault(row["Kernel Name"][:70], [0, 0.0]); a[0] += 1; a[1] += float(row["Metric Value"].replace(",", ""))
for k, a in agg.items(): print("  %-72s %3d  %9.1f us avg" % (k, a[0], a[1] / a[0] / 1e3))
PY
done
for spec in "config2 --no-cpu-baseline" "config2 --precision f16r --no-cpu-baseline" "config1 --no-cpu-baseline"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 500 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  launches/step %.1f  stages %s  roofline %.3f (%.0f TF/s) clocks %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"]))
except Exception as e:
    print("  no result:", e)
PY
done
