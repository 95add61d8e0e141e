#!/bin/bash
# A/B of build variants and knobs on one box: prints one short line per configuration.
TAG=${1:-tune}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-steady $EXTRA > $OUT/$name.json 2> $OUT/$name.err
  python - "$OUT/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "sims/s %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], d["clocks"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run base C6_DUMMY=1
for v in alphazero-connect6_b200/lib/variants/*.so; do
  n=$(basename $v .so)
  run $n C6_LIB=$PWD/$v
  run ${n}_fetch32 C6_LIB=$PWD/$v C6_L2_FETCH=32
done
EXTRA="--epsilon 0"
for v in alphazero-connect6_b200/lib/variants/*.so; do
  n=$(basename $v .so)
  run ${n}_eps0 C6_LIB=$PWD/$v
done
