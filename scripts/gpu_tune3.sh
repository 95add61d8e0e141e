#!/bin/bash
# One gpurun call: GPU parity suite, the default bench line, then an A/B of experiment builds on the same box
# (alphazero-connect6_b200/lib/variants/lib_*.so from `make -C alphazero-connect6_b200 exp NAME=.. DEFS=..` for the search kernel,
# libm_*.so from `make expm` for the micro-benchmark kernels).
TAG=${1:-tune3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt; lscpu | head -25 >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-steady $EXTRA > $OUT/$name.json 2> $OUT/$name.err
  python - "$OUT/$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "sims/s %.4g"%d["value"], "ms/step %.2f"%d["ms_per_step"], "kernel ms %.2f"%d["roofline"]["kernel_ms_per_launch"], "frac %.4f"%d["roofline"]["frac"])
except Exception as e: print(sys.argv[2], "FAILED", e)
PY
}
run base C6_DUMMY=1
for v in alphazero-connect6_b200/lib/variants/lib_*.so; do
  [ -e "$v" ] || continue
  n=$(basename $v .so)
  run $n C6_LIB=$PWD/$v
done
EXTRA="--epsilon 0" run base_eps0 C6_DUMMY=1
run base_again C6_DUMMY=1
echo "micro base" ; timeout 300 python scripts/micro.py --no-cpu --quick 2>&1 | head -2 | tee $OUT/micro_base.jsonl
for v in alphazero-connect6_b200/lib/variants/libm_*.so; do
  [ -e "$v" ] || continue
  echo "micro $(basename $v .so)"; C6_LIB=$PWD/$v timeout 300 python scripts/micro.py --no-cpu --quick 2>&1 | head -2 | tee $OUT/micro_$(basename $v .so).jsonl
done
