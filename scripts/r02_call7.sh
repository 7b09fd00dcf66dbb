#!/bin/bash
# Round 2, GPU call 7 (one B200), after the container was re-created: the driver's own GPU test command, fused embed v2
# timing, the default bench line + reference arm, ncu (full) of the fused embed kernel, launch list of two bench steps.
OUT=gpurun_out/r02_call7
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -14 $OUT/pytest_gpu.log
timeout 300 python scripts/tune_embed_fused.py 100 > $OUT/tune_embed_fused.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_embed_fused.log
for spec in "config2" "config2 --keep-z"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  launches/step %.1f  stages %s  roofline frac %.3f (%.0f TF/s) clocks %s cpu %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"], d["clocks"], d.get("cpu_baseline")))
except Exception as e:
    print("  no result:", e)
PY
done
( time timeout 400 python bench.py --impl reference ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm rc=$?"; cat $OUT/bench_reference.json | cut -c1-600; tail -4 $OUT/bench_reference.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_fused -s 2 -c 1 -o $OUT/r02_embed_fused \
  python bench.py --workload config2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_embed.log 2>&1; echo "ncu embed rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 60 --csv --log-file $OUT/launches.csv python bench.py --workload config2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/launches.log 2>&1; echo "launch list rc=$?"
ls -la $OUT
