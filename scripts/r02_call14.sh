#!/bin/bash
# Round 2, GPU call 14 (one B200): lean fused embed walking DOWN patch columns (one bulk copy per ring slot) -- embed tests, knob sweep, all GPU tests, bench, ncu.
OUT=gpurun_out/r02_call14
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "embed" > $OUT/pytest_embed.log 2>&1; echo "embed tests rc=$?"; tail -8 $OUT/pytest_embed.log
timeout 400 python scripts/tune_embed_fused.py 100 > $OUT/tune_embed_fused.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_embed_fused.log
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -3 $OUT/pytest_gpu.log
for spec in "config2"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --workload $spec --warmup 3 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"; tail -3 $OUT/bench_$name.err
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  launches/step %.1f  stages %s  roofline frac %.3f (%.0f TF/s) clocks %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"], d["clocks"]))
except Exception as e:
    print("  no result:", e)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_fused -s 2 -c 1 -o $OUT/r02_embed_lean5 \
  python bench.py --workload config2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_embed.log 2>&1; echo "ncu embed rc=$?"
ls -la $OUT
