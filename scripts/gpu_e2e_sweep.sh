#!/bin/bash
# e2e (C-ABI host-callback path) sweep: pipeline depth, graphs, sample-evaluator pool settings; PCIe probe.
OUT=gpurun_out/${1:-e2e}
mkdir -p $OUT
python - <<'PY'
import sys, ctypes as C
sys.path.insert(0, "tests"); import c6lib
L = c6lib.product()
a, b, c = C.c_double(), C.c_double(), C.c_double()
for mb in (2, 8, 32):
    L.c6_probe_pcie(mb << 20, 20, 0, C.byref(a), C.byref(b), C.byref(c))
    print(f"pcie {mb} MiB: h2d {a.value:.1f} GB/s, d2h {b.value:.1f} GB/s, both {c.value:.1f} GB/s per direction")
PY
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  python -c "
import json,sys
try:
    d=json.load(open(sys.argv[1])); e=d['e2e']; print(sys.argv[2], 'e2e %.4g'%e['value'], 'ms/wave %.3f'%e['ms_per_wave'], e['host_ms_per_wave'])
except Exception as ex: print(sys.argv[2], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-300:])" $OUT/$name.json $name
}
run p8_graphs C6_PIPELINE=8
run p8_nographs C6_PIPELINE=8 C6_GRAPHS=0
run p16_graphs C6_PIPELINE=16
run p16_spin0 C6_PIPELINE=16 C6_SAMPLE_SPIN=0
run p16_spin20000 C6_PIPELINE=16 C6_SAMPLE_SPIN=20000
run p16_t8 C6_PIPELINE=16 C6_SAMPLE_THREADS=8
run p16_t8_spin20000 C6_PIPELINE=16 C6_SAMPLE_THREADS=8 C6_SAMPLE_SPIN=20000
run p16_t4 C6_PIPELINE=16 C6_SAMPLE_THREADS=4
run p12 C6_PIPELINE=12
run p6 C6_PIPELINE=6
