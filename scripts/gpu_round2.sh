#!/bin/bash
# One gpurun call (round 2): GPU parity tests, smoke, bench lines (both arms), ncu launch list, full captures of the fused search
# kernel and of the secondary kernels, configs[0] / configs[2] scripts.
# usage (from the repo root on the GPU box): bash scripts/gpu_round2.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/clocks.csv &
SMI=$!
nproc > $OUT/nproc.txt; lscpu | head -25 >> $OUT/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
cat $OUT/bench.json
timeout 600 python bench.py --epsilon 0 --no-e2e --no-cpu-baseline --no-steady > $OUT/bench_eps0.json 2> $OUT/bench_eps0.err
cat $OUT/bench_eps0.json
timeout 600 python bench.py --virtual-loss 4 --no-e2e --no-cpu-baseline --no-steady > $OUT/bench_vl4.json 2> $OUT/bench_vl4.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err
cat $OUT/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-steady > $OUT/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sims_device -s 3 -c 1 -o $OUT/prof_sims \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-steady > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?" >> $OUT/ncu_full.log
tail -3 $OUT/ncu_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_win_check|k_playouts|k_expand_records" -c 5 -o $OUT/prof_secondary \
    python scripts/secondary_kernels.py > $OUT/ncu_secondary.log 2>&1; echo "ncu secondary rc=$?" >> $OUT/ncu_secondary.log
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step|k_gather" -s 40 -c 4 -o $OUT/prof_callback \
    python scripts/secondary_kernels.py > $OUT/ncu_callback.log 2>&1; echo "ncu callback rc=$?" >> $OUT/ncu_callback.log
timeout 600 python scripts/micro.py > $OUT/micro.jsonl 2> $OUT/micro.err; cat $OUT/micro.jsonl
timeout 600 python scripts/config1.py > $OUT/config1.jsonl 2> $OUT/config1.err; cat $OUT/config1.jsonl
timeout 400 ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section MemoryWorkloadAnalysis \
    --clock-control none -k regex:k_default_policy -c 1 -o $OUT/prof_default_policy python scripts/dp_profile.py > $OUT/ncu_dp.log 2>&1; echo "ncu dp rc=$?" >> $OUT/ncu_dp.log
kill $SMI
ls -la $OUT
