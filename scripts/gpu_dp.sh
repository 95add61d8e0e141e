#!/bin/bash
# DefaultPolicy: GPU parity tests, throughput, a bounded ncu capture
TAG=${1:-dp}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_default_policy.py tests/test_parity_gpu.py -m gpu -k "default_policy or playout or win_check" -x -q > $OUT/pytest_dp.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_dp.log
tail -4 $OUT/pytest_dp.log
timeout 600 python scripts/micro.py --no-cpu 2>&1 | tee $OUT/micro.jsonl | cut -c1-330
for v in alphazero-connect6_b200/lib/variants/libm_*.so; do
  [ -e "$v" ] || continue
  echo "micro $(basename $v .so)"; C6_LIB=$PWD/$v timeout 600 python scripts/micro.py --no-cpu 2>&1 | tee $OUT/micro_$(basename $v .so).jsonl | tail -2 | cut -c1-330
done
python scripts/dp_profile.py
timeout 400 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis \
    --clock-control none -k regex:k_default_policy -c 1 -o $OUT/prof_default_policy python scripts/dp_profile.py > $OUT/ncu_dp.log 2>&1; echo "ncu rc=$?" >> $OUT/ncu_dp.log
tail -3 $OUT/ncu_dp.log
