#!/bin/bash
# e2e leg with the sample evaluator's fill variants (C6_SAMPLE_ISA x C6_SAMPLE_NT), then the GPU parity suite
TAG=${1:-e2e_isa}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for isa in avx512 avx2; do for nt in 1 0; do
  C6_SAMPLE_ISA=$isa C6_SAMPLE_NT=$nt timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-steady > $OUT/e2e_${isa}_nt$nt.json 2> $OUT/e2e_${isa}_nt$nt.err
  python - "$OUT/e2e_${isa}_nt$nt.json" "$isa nt=$nt" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d["e2e"]; r=d["e2e_reading_evaluator"]
    print(sys.argv[2], "e2e %.4g sims/s, %.3f ms/wave, host %s, floor %.3f | reading %.4g"%(e["value"], e["ms_per_wave"], e["host_ms_per_wave"], e["host_link"]["floor_ms_per_wave"], r["value"]))
except Exception as ex: print(sys.argv[2], "FAILED", ex)
PY
done; done
C6_SAMPLE_ISA=avx512 C6_SAMPLE_NT=1 timeout 300 python bench.py --steps 3 --no-cpu-baseline --no-steady > $OUT/e2e_again.json 2>/dev/null
python -c "import json; d=json.load(open('$OUT/e2e_again.json')); print('avx512 nt=1 again', d['e2e']['value'], d['e2e']['ms_per_wave'])"
timeout 800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -2 $OUT/pytest_gpu.log
