#!/bin/bash
# Round 2, multi-GPU call after the embed rework (call 16):  gpurun --gpus N -- 'bash scripts/r02_call16.sh N'
#  1. the driver's own scaling command (bench.py under torchrun, default workload);  2. scripts/r02_multi.py (checks against one
#  GPU, every workload's bench line, timelines) in one process group.
N=${1:-2}
OUT=gpurun_out/r02_multi_v7
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw --format=csv > $OUT/gpus_n$N.txt 2>&1
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 30 --warmup 3 ) > $OUT/driver_style_bench_n$N.json 2> $OUT/driver_style_bench_n$N.err; echo "driver-style bench rc=$?"
python - "$OUT/driver_style_bench_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("  images/s %.0f  ms/step %.3f  e2e %s  parity_ok %s  launches/step %.1f  stages %s  roofline %.3f  clocks %s" % (
        d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("parity_ok"), d["gpu_launches"] / d["steps"],
        {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["stages"].items()}, d["roofline"]["frac"], d["clocks"]))
except Exception as e:
    print("  no result:", e)
PY
tail -3 $OUT/driver_style_bench_n$N.err
timeout ${2:-600} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 scripts/r02_multi.py $OUT > $OUT/stdout_n$N.log 2> $OUT/stderr_n$N.log; echo "multi rc=$?"
tail -40 $OUT/log_n$N.txt
grep -v "^$" $OUT/stderr_n$N.log | tail -5
ls $OUT
