#!/bin/bash
# Round 2, multi-GPU call: everything in one process group (scripts/r02_multi.py).  gpurun --gpus N -- 'bash scripts/r02_call4.sh N'
N=${1:-2}
OUT=gpurun_out/r02_multi
mkdir -p $OUT
timeout ${2:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 scripts/r02_multi.py $OUT > $OUT/stdout_n$N.log 2> $OUT/stderr_n$N.log; echo "multi rc=$?"
cat $OUT/log_n$N.txt | tail -60
tail -5 $OUT/stderr_n$N.log
ls $OUT
