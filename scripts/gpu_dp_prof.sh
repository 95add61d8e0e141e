#!/bin/bash
# source-level capture of k_default_policy only
TAG=${1:-dpp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section SourceCounters --import-source on \
    --clock-control none -k regex:k_default_policy -c 1 -o $OUT/prof_default_policy python scripts/dp_profile.py > $OUT/ncu_dp.log 2>&1; echo "ncu rc=$?" >> $OUT/ncu_dp.log
tail -3 $OUT/ncu_dp.log
