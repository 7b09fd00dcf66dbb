#!/bin/bash
# Round 2, GPU call 1 (one B200): BASELINE-size parity tests against the oracle, the bench workloads, one ncu capture of
# the all-pairs kernel.  gpurun --timeout 1500 -- 'bash scripts/r02_call1.sh'
OUT=gpurun_out/r02_call1
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
timeout 1200 python -m pytest tests/test_gpu_baseline_sizes.py -q -x --durations=10 > $OUT/pytest_baseline_sizes.log 2>&1; echo "baseline-size tests rc=$?"; tail -15 $OUT/pytest_baseline_sizes.log
timeout 400 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_baseline_sizes.py > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 $OUT/pytest_gpu.log
for spec in "config2" "config2 --keep-z" "config2 --no-symmetry" "config3" "config1" "config4pc --cpu-sample 1" "config5 --cpu-sample 1 --steps 5"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --workload $spec --steps 10 --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  stages %s  roofline frac %.3f (%.0f TF/s)  cpu %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"),
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"], (d.get("cpu_baseline") or {}).get("value")))
    print("  parity", d.get("parity"))
except Exception as e:
    print("  no result:", e)
PY
done
# all-pairs (supervised) kernel: full ncu capture of one launch
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mindist_tc -s 3 -c 1 -o $OUT/r02_mindist_allpairs \
  python bench.py --workload config3 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_allpairs.log 2>&1; echo "ncu all-pairs rc=$?"
ls -la $OUT
