#!/bin/bash
OUT=gpurun_out/${1:-e2e7}
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -15 $OUT/pytest_gpu.log
run() { name=$1; shift
  env "$@" timeout 120 python scripts/e2e_probe.py > $OUT/$name.json 2> $OUT/$name.err || tail -3 $OUT/$name.err
  echo "$name $(cat $OUT/$name.json | cut -c1-150)"
}
for rep in 1 2; do
for Z in 1 0; do
for LA in 1 0; do
for P in 4 8 16; do
for T in 3 4; do
run z${Z}_la${LA}_p${P}_t${T}_r$rep C6_ZEROCOPY=$Z C6_LAUNCHER=$LA C6_PIPELINE=$P C6_SAMPLE_THREADS=$T
done; done; done; done; done
