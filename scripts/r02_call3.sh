#!/bin/bash
# Round 2, GPU call 3 (one B200): refine kernel v2 (4 warps per row, batched loads, dot form), per-category batched launch.
OUT=gpurun_out/r02_call3
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x --ignore=tests/test_gpu_baseline_sizes.py > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 $OUT/pytest_gpu.log
timeout 200 python scripts/precision_table.py 100 > $OUT/precision_table.log 2>&1; echo "precision table rc=$?"; cat $OUT/precision_table.log
timeout 1500 python -m pytest tests/test_gpu_baseline_sizes.py -q --durations=5 > $OUT/pytest_baseline_sizes.log 2>&1; echo "baseline-size tests rc=$?"; tail -12 $OUT/pytest_baseline_sizes.log
for spec in "config2 --precision f16r" "config4pc --cpu-sample 1 --steps 5" "config5 --cpu-sample 1 --steps 5" "config3 --steps 5"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  launches/step %.1f  stages %s  roofline frac %.3f (%.0f TF/s)" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"]))
    print("  parity", d.get("parity"))
except Exception as e:
    print("  no result:", e)
PY
done
ls -la $OUT
