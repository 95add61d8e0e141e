#!/bin/bash
# Multi-GPU check on one box: bench.py under torchrun exactly as the driver launches it, the reference arm
# the same way, and the GPU tests that need more than one device.
# usage: bash scripts/gpu_multi.sh N [tag]
N=${1:-2}
TAG=${2:-multi$N}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 6 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
tail -3 $OUT/bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?" >> $OUT/bench_reference.err
timeout 600 python -m pytest tests -m gpu -x -q -k "shard" > $OUT/pytest_shard.log 2>&1; tail -3 $OUT/pytest_shard.log
python scripts/show_bench.py $OUT/bench.json $OUT/bench_reference.json
