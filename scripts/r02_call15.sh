#!/bin/bash
# Round 2, GPU call 15 (one B200): lean fused embed at 4 CTAs per SM (5-slot ring, 96 registers) against 3 -- embed tests, sweep.
OUT=gpurun_out/r02_call15
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "embed" > $OUT/pytest_embed.log 2>&1; echo "embed tests rc=$?"; tail -3 $OUT/pytest_embed.log
timeout 400 python scripts/tune_embed_fused.py 100 > $OUT/tune_embed_fused.log 2>&1; echo "tune rc=$?"; cat $OUT/tune_embed_fused.log
