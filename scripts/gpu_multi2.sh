#!/bin/bash
# One multi-GPU gpurun call (round 2): the bench as the driver launches it (value + config4 + both e2e legs), host-side variants
# of the callback path, and BASELINE configs[4] under torchrun (fp32 as written, bf16 as deployed).
# usage: gpurun --gpus N -- 'bash scripts/gpu_multi2.sh <tag> <N>'
TAG=${1:-r02m}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt; lscpu | head -30 >> $OUT/nproc.txt; numactl -H >> $OUT/nproc.txt 2>&1; nvidia-smi topo -m >> $OUT/nproc.txt 2>&1
timeout 600 python -m pytest tests/test_replay.py "tests/test_parity_gpu.py::test_host_mirror_random_evaluator_and_async_agent" -m gpu -q > $OUT/pytest_tail.log 2>&1; tail -3 $OUT/pytest_tail.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err; echo "rc=$?" >> $OUT/bench_${N}gpu.err
cat $OUT/bench_${N}gpu.json
timeout 600 $TR --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $OUT/bench_reference_${N}gpu.json 2> $OUT/bench_reference_${N}gpu.err
cat $OUT/bench_reference_${N}gpu.json
e2e() { name=$1; shift
  env "$@" timeout 600 $TR --master-port 29513 bench.py --gpus $N --steps 3 --no-steady --no-config4 > $OUT/e2e_$name.json 2> $OUT/e2e_$name.err
  python - "$OUT/e2e_$name.json" "$name" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d["e2e"]; r=d["e2e_reading_evaluator"]
    print(sys.argv[2], "e2e %.4g sims/s, %.3f ms/wave, host %s | reading %.4g sims/s | %s"%(e["value"], e["ms_per_wave"], e["host_ms_per_wave"], r["value"], d["host"]))
except Exception as ex: print(sys.argv[2], "FAILED", ex)
PY
}
e2e spin200 C6_SPIN_US=200
e2e nobind C6_NUMA_BIND=0
e2e pool2 C6_SAMPLE_THREADS=2
e2e pipe8 C6_PIPELINE=8
timeout 900 $TR --master-port 29514 scripts/config5.py --dtype fp32 --waves 2 > $OUT/config5_fp32_${N}gpu.jsonl 2> $OUT/config5_fp32.err; cat $OUT/config5_fp32_${N}gpu.jsonl
timeout 600 $TR --master-port 29515 scripts/config5.py --dtype bf16 --waves 4 > $OUT/config5_bf16_${N}gpu.jsonl 2> $OUT/config5_bf16.err; cat $OUT/config5_bf16_${N}gpu.jsonl
ls -la $OUT
