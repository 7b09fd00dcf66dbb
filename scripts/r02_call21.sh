#!/bin/bash
# Round 2, GPU call 21 (one B200): final records at HEAD -- the driver's GPU test command, smoke(), the driver's bench commands.
OUT=gpurun_out/r02_call21
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -9 $OUT/pytest_gpu.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/smoke.log
( time timeout 600 python bench.py ) > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench rc=$?"; tail -2 $OUT/bench_default.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference bench rc=$?"
python - $OUT <<'PY'
import glob, json, sys
for f in sorted(glob.glob(sys.argv[1] + "/bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        if d.get("impl") == "reference":
            print(f.split("/")[-1], "reference arm:", d["value"], "images/s", d["ms_per_step"], "ms/step")
            continue
        print("%s  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  launches/step %.1f  stages %s  roofline %.3f (%.0f TF/s)  cpu %s  gpu-torch-loop %s clocks %s" % (
            f.split("/")[-1], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"), d["gpu_launches"] / d["steps"],
            {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
            d["roofline"]["frac"], d["roofline"]["achieved"], (d.get("cpu_baseline") or {}).get("value"), (d.get("reference_loop_on_gpu") or {}).get("value"), d["clocks"]))
    except Exception as e:
        print(f, "no result:", e)
PY
