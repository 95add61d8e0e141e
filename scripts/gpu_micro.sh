#!/bin/bash
# One gpurun call for the secondary kernels: GPU parity suite, configs[2] micro-benchmarks (this build and the libm_* variants
# from `make -C alphazero-connect6_b200 expm NAME=.. DEFS=..`), configs[0] single-game timings, and a full ncu capture of
# k_default_policy / k_win_check / k_playouts.
TAG=${1:-micro}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -4 $OUT/pytest_gpu.log
echo "micro base"; timeout 600 python scripts/micro.py --no-cpu 2>&1 | tee $OUT/micro.jsonl
for v in alphazero-connect6_b200/lib/variants/libm_*.so; do
  [ -e "$v" ] || continue
  echo "micro $(basename $v .so)"; C6_LIB=$PWD/$v timeout 300 python scripts/micro.py --no-cpu --quick 2>&1 | head -2 | tee $OUT/micro_$(basename $v .so).jsonl
done
echo "micro base again"; timeout 600 python scripts/micro.py --no-cpu --quick 2>&1 | tee $OUT/micro_again.jsonl
timeout 600 python scripts/config1.py > $OUT/config1.jsonl 2> $OUT/config1.err; cat $OUT/config1.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_win_check|k_playouts" -c 5 -o $OUT/prof_secondary \
    python scripts/secondary_kernels.py > $OUT/ncu_secondary.log 2>&1; echo "ncu secondary rc=$?" >> $OUT/ncu_secondary.log
tail -2 $OUT/ncu_secondary.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_default_policy" -c 1 -o $OUT/prof_default_policy \
    python scripts/secondary_kernels.py > $OUT/ncu_default_policy.log 2>&1; echo "ncu default policy rc=$?" >> $OUT/ncu_default_policy.log
tail -2 $OUT/ncu_default_policy.log
