#!/bin/bash
# One gpurun call: GPU parity tests, a bench line and one full ncu capture of the fused search kernel (report kept).
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
timeout 300 python bench.py --epsilon 0 --no-e2e --no-cpu-baseline --no-steady --steps 6 > $OUT/bench_eps0.json 2> $OUT/bench_eps0.err
cat $OUT/bench_eps0.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sims_device -s 3 -c 1 -o $OUT/prof_sims \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-steady > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $OUT/ncu_full.log
tail -3 $OUT/ncu_full.log
ls -la $OUT
