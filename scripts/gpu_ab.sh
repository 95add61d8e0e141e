#!/bin/bash
# A/B only (no test suite): base build vs alphazero-connect6_b200/lib/variants/lib_*.so, bench.py device leg.
TAG=${1:-ab}
OUT=gpurun_out/$TAG
mkdir -p $OUT
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
  run $(basename $v .so) C6_LIB=$PWD/$v
done
run l2_32 C6_L2_FETCH=32
run l2_128 C6_L2_FETCH=128
run base_again C6_DUMMY=1
# DRAM traffic of one launch of the search kernel per build (two-metric ncu pass: cheap)
traffic() { name=$1; shift
  env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_sims_device -s 3 -c 1 --csv --log-file $OUT/traffic_$name.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-steady > /dev/null 2>&1
  echo "traffic $name: $(grep -E 'dram__bytes|hit_rate' $OUT/traffic_$name.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tr -d '"' | tr '\n' ';')"
}
traffic base C6_DUMMY=1
for v in alphazero-connect6_b200/lib/variants/lib_*.so; do
  [ -e "$v" ] || continue
  traffic $(basename $v .so) C6_LIB=$PWD/$v
done
