#!/bin/bash
# Round 2, GPU call 10 (one B200): the driver's GPU test command, then the records of the round: ncu --set full of the four
# dominant kernels (symmetric GEMM, all-pairs GEMM, exact refine, lean fused embed), the launch list of two bench steps, and
# one bench line per BASELINE workload.
OUT=gpurun_out/r02_call10
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -9 $OUT/pytest_gpu.log
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:mindist_tc -s 3 -c 1 -o $OUT/r02_mindist_sym python bench.py --workload config2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_sym.log 2>&1; echo "ncu sym rc=$?"
timeout 300 $NCU -k regex:mindist_tc -s 3 -c 1 -o $OUT/r02_mindist_allpairs python bench.py --workload config3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_allpairs.log 2>&1; echo "ncu all-pairs rc=$?"
timeout 300 $NCU -k regex:refine -s 3 -c 1 -o $OUT/r02_refine python bench.py --workload config2 --precision f16r --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_refine.log 2>&1; echo "ncu refine rc=$?"
timeout 300 $NCU -k regex:embed_fused -s 3 -c 1 -o $OUT/r02_embed_lean python bench.py --workload config2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_embed.log 2>&1; echo "ncu embed rc=$?"
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --workload config2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --cuda-profiler > $OUT/launches.log 2>&1; echo "launch list rc=$?"
for spec in "config2" "config2 --precision f16r" "config3 --steps 5" "config1" "config4pc --cpu-sample 1 --steps 5" "config5 --cpu-sample 1 --steps 5"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 400 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  launches/step %.1f  stages %s  roofline frac %.3f (%.0f TF/s)  cpu %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"], (d.get("cpu_baseline") or {}).get("value")))
except Exception as e:
    print("  no result:", e)
PY
done
ls -la $OUT
