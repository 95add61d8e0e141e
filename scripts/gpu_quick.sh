#!/bin/bash
# One short gpurun call: GPU parity tests + smoke + the default bench line (no profiler).
# usage: bash scripts/gpu_quick.sh [tag]
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log
timeout 900 python bench.py $BENCH_ARGS > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
timeout 600 python bench.py --epsilon 0 --no-e2e --no-cpu-baseline --no-steady > $OUT/bench_eps0.json 2> $OUT/bench_eps0.err
cat $OUT/bench_eps0.json
