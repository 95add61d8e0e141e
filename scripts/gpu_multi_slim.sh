#!/bin/bash
# bench.py under torchrun exactly as the driver launches it (both arms), nothing else.  usage: gpurun --gpus N -- 'bash scripts/gpu_multi_slim.sh <tag> <N>'
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?" >> $OUT/bench_${N}gpu.err
cat $OUT/bench_${N}gpu.json
timeout 600 $TR --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_${N}gpu.json 2> $OUT/bench_reference_${N}gpu.err
cat $OUT/bench_reference_${N}gpu.json
