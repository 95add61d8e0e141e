#!/bin/bash
# r02p: parity suite on the allocation cache, single-game timings, DefaultPolicy capture, one default bench line
TAG=${1:-r02p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 600 python scripts/config1.py > $OUT/config1.jsonl 2> $OUT/config1.err; head -4 $OUT/config1.jsonl
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; cat $OUT/bench.json | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_default_policy" -c 1 -o $OUT/prof_default_policy \
    python scripts/secondary_kernels.py > $OUT/ncu_default_policy.log 2>&1; echo "ncu default policy rc=$?" >> $OUT/ncu_default_policy.log
tail -2 $OUT/ncu_default_policy.log
