#!/usr/bin/env python
"""One small launch of every secondary kernel of the path, for `ncu --set full -k regex:...` (scripts/gpu_round2.sh):
k_win_check / k_random_boards / k_playouts (configs[2]), k_step + k_gather (callback path, 1,024 games, a few waves),
k_default_policy (4,096 games), k_expand_records (replay feed)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-connect6_b200", "python"))
from pyconnect6 import _capi as capi  # noqa: E402

L = capi.product()
ms = C.c_float()
assert L.c6_bench_win_check(1 << 24, 0xC6, 2, 0, C.byref(ms), None, None, 0) == 0, L.c6_last_error()
tot = C.c_uint64()
assert L.c6_bench_playouts(0, 1 << 20, 0xC6, 1, 0, C.byref(ms), None, None, None, 0, C.byref(tot)) == 0, L.c6_last_error()
cfg = capi.make_config(num_simulation=800, epsilon=0.25, seed=0xC6, evaluator=capi.EVAL_CALLBACK, slots=4096)
e = L.c6_engine_create(C.byref(cfg))
assert e, L.c6_last_error()
assert L.c6_engine_start(e, 0, 1 << 20) == 0
h2d, d2h = C.c_double(), C.c_double()
assert L.c6_engine_timed_waves(e, C.cast(L.c6_sample_random_callback, capi.CALLBACK), 40, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
L.c6_engine_destroy(e)
n = 4096
rec = (capi.Record * n)()
st = capi.DpStats()
assert L.c6_default_policy_play(50, 0xC6, 0, n, 0, 0, rec, None, None, C.byref(st)) == 0, L.c6_last_error()
rows = sum(rec[g].n_moves for g in range(n))
FP = C.POINTER(C.c_float)
bufs = [(C.c_float * rows)() for _ in range(3)] + [(C.c_float * (rows * 225))()]
nr = C.c_longlong()
assert L.c6_expand_records_host(rec, n, 0, bufs[0], bufs[1], bufs[3], bufs[2], C.byref(nr)) == 0, L.c6_last_error()
print("secondary kernels launched;", rows, "replay rows")
