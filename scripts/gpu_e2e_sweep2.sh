#!/bin/bash
OUT=gpurun_out/${1:-e2e}
mkdir -p $OUT
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d['e2e']; print(sys.argv[2], 'e2e %.4g'%e['value'], 'ms/wave %.3f'%e['ms_per_wave'], e['host_ms_per_wave'])
except Exception as ex: print(sys.argv[2], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-300:])" $OUT/$name.json $name
}
for rep in 1 2; do
for G in 0 1; do
for T in 4 8 16; do
for S in 0 20000; do
run p16_g${G}_t${T}_s${S}_r$rep C6_PIPELINE=16 C6_GRAPHS=$G C6_SAMPLE_THREADS=$T C6_SAMPLE_SPIN=$S
done; done; done; done
