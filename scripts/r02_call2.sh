#!/bin/bash
# Round 2, GPU call 2 (one B200): refined precision mode (arg-min + exact re-evaluation) -- kernel tests, BASELINE-size
# parity against the oracle, precision / timing table, benches with precision=auto; ncu of the embed kernel (Z-free).
OUT=gpurun_out/r02_call2
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "refine or arg" > $OUT/pytest_refine.log 2>&1; echo "refine kernel tests rc=$?"; tail -12 $OUT/pytest_refine.log
timeout 200 python scripts/precision_table.py 100 > $OUT/precision_table.log 2>&1; echo "precision table rc=$?"; cat $OUT/precision_table.log
timeout 200 python scripts/precision_table.py 100 1.45 > $OUT/precision_table_norm46.log 2>&1; echo "precision table (norm 46) rc=$?"; cat $OUT/precision_table_norm46.log
timeout 1500 python -m pytest tests/test_gpu_baseline_sizes.py -q --durations=10 > $OUT/pytest_baseline_sizes.log 2>&1; echo "baseline-size tests rc=$?"; tail -25 $OUT/pytest_baseline_sizes.log
timeout 400 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_baseline_sizes.py > $OUT/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -5 $OUT/pytest_gpu.log
for spec in "config5 --cpu-sample 1 --steps 5" "config5 --cpu-sample 1 --steps 5 --precision f16" "config2 --precision f16r"; do
  name=$(echo $spec | tr ' ' '_' | tr -d '-')
  timeout 300 python bench.py --workload $spec --warmup 3 > $OUT/bench_$name.json 2> $OUT/bench_$name.err; echo "bench $spec rc=$?"
  python - "$OUT/bench_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  images/s %.0f  ms/step %.3f  e2e %.0f  parity_ok %s  stages %s  roofline frac %.3f (%.0f TF/s)" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value", 0), d.get("parity_ok"),
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items() if k != "comm_ms_per_step_rank0"},
        d["roofline"]["frac"], d["roofline"]["achieved"]))
    print("  parity", d.get("parity"))
except Exception as e:
    print("  no result:", e)
PY
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:embed_tma -s 6 -c 2 -o $OUT/r02_embed_zfree \
  python bench.py --workload config2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_embed.log 2>&1; echo "ncu embed rc=$?"
ls -la $OUT
