#!/usr/bin/env python
"""End-to-end leg of bench.py alone (C-ABI host Callback path, sample host RandomEvaluator): one JSON line.
  python scripts/e2e_probe.py [--games 4096] [--waves 96] [--warm 16]      (knobs via C6_* environment variables)"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import c6lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=4096)
ap.add_argument("--waves", type=int, default=96)
ap.add_argument("--warm", type=int, default=16)
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--device", type=int, default=0)
a = ap.parse_args()
L = c6lib.product()
cfg = c6lib.make_config(num_simulation=a.sims, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6, device=a.device, slots=a.games,
                        log2_records=18, evaluator=c6lib.EVAL_CALLBACK)
e = L.c6_engine_create(C.byref(cfg))
assert e, L.c6_last_error()
assert L.c6_engine_start(e, 0, 50_000_000) == 0
cb = C.cast(L.c6_sample_random_callback, c6lib.CALLBACK)
ms, h2d, d2h = C.c_float(), C.c_double(), C.c_double()
assert L.c6_engine_timed_waves(e, cb, a.warm, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
s0, s1 = c6lib.Stats(), c6lib.Stats()
L.c6_engine_stats(e, C.byref(s0))
hw0, hc0, hw1, hc1 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
L.c6_engine_host_times(e, C.byref(hw0), C.byref(hc0))
assert L.c6_engine_timed_waves(e, cb, a.waves, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
L.c6_engine_stats(e, C.byref(s1))
L.c6_engine_host_times(e, C.byref(hw1), C.byref(hc1))
sims = s1.sims - s0.sims
print(json.dumps({"e2e_sims_per_s": sims / (ms.value * 1e-3), "ms_per_wave": ms.value / a.waves,
                  "in_callback_ms": 1e3 * (hc1.value - hc0.value) / a.waves, "caller_wait_ms": 1e3 * (hw1.value - hw0.value) / a.waves,
                  "h2d_MB": h2d.value / a.waves / 1e6, "d2h_MB": d2h.value / a.waves / 1e6,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("C6_")}}))
L.c6_engine_destroy(e)
