#!/bin/bash
# Round 2, GPU call 17 (one B200): the records of the final build -- the driver's GPU test command, smoke(), the launch list of
# two bench steps, the driver's bench commands (ours and --impl reference), one bench line per BASELINE workload.
OUT=gpurun_out/r02_call17
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -9 $OUT/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke.log
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --workload config2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --cuda-profiler > $OUT/launches.log 2>&1; echo "launch list rc=$?"; grep -c "mindist\|embed" $OUT/launches.csv
( time timeout 600 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench rc=$?"; tail -2 $OUT/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference bench rc=$?"; tail -2 $OUT/bench_reference.err; cat $OUT/bench_reference.json | cut -c1-600
for spec in "config2 --precision f16r --no-cpu-baseline" "config2 --keep-z --no-cpu-baseline" "config3 --steps 5" "config1" "config4pc --cpu-sample 1 --steps 5" "config5 --cpu-sample 1 --steps 5"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 500 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
done
python - $OUT <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        if d.get("impl") == "reference":
            print(f.split("/")[-1], "reference arm:", d["value"], "images/s", d["ms_per_step"], "ms/step", d.get("cpu_baseline", {}).get("sample"))
            continue
        print("%s  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  launches/step %.1f  stages %s  roofline %.3f (%.0f TF/s)  cpu %s  gpu-torch-loop %s" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"), d["gpu_launches"] / d["steps"],
            {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
            d["roofline"]["frac"], d["roofline"]["achieved"], (d.get("cpu_baseline") or {}).get("value"), d.get("reference_loop_on_gpu")))
    except Exception as e:
        print(f, "no result:", e)
PY
ls -la $OUT
