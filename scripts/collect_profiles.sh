#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of one bench step + full ncu captures of the two
# dominant kernels.  Outputs land in gpurun_out/ and are summarised by scripts/summarize_profiles.py.
set -x
R=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${R}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cuda-profiler \
    > gpurun_out/ncu_launches_${R}.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mindist_tc -c 1 \
    -o gpurun_out/prof_mindist_${R} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --cuda-profiler \
    > gpurun_out/ncu_mindist_${R}.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:embed_tma -c 2 \
    -o gpurun_out/prof_embed_${R} python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --cuda-profiler \
    > gpurun_out/ncu_embed_${R}.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
cat gpurun_out/bench_${R}.json
