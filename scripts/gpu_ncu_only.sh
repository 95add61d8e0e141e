#!/bin/bash
# One gpurun call: one full ncu capture of the fused search kernel (report kept) and nothing else.
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sims_device -s 3 -c 1 -o $OUT/prof_sims \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-steady $EXTRA > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $OUT/ncu_full.log
tail -3 $OUT/ncu_full.log
ls -la $OUT
