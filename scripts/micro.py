#!/usr/bin/env python
"""BASELINE.json configs[2]: win-detection / random-playout micro-benchmarks on one B200.
16,777,216 random boards (bit-packed 64 B each) -> 1 B winner: 65 algorithmic bytes per board;
16,777,216 random playouts from the empty board: 72 B written per playout.  Prints one JSON line each."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import c6lib  # noqa: E402

L = c6lib.product()
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
n = 1 << 24
ms = C.c_float()
assert L.c6_bench_win_check(n, 0xC6, 20, 0, C.byref(ms), None, None, 0) == 0, L.c6_last_error()
gbs = n * 65 / (ms.value * 1e-3) / 1e9
print(json.dumps({"bench": "win_check", "boards": n, "ms": ms.value, "boards_per_s": n / (ms.value * 1e-3), "algorithmic_GBps": gbs,
                  "peak_GBps": peak, "frac": gbs / peak}))
tot = C.c_uint64()
assert L.c6_bench_playouts(0, n, 0xC6, 3, 0, C.byref(ms), None, None, None, 0, C.byref(tot)) == 0, L.c6_last_error()
print(json.dumps({"bench": "random_playouts", "playouts": n, "ms": ms.value, "playouts_per_s": n / (ms.value * 1e-3),
                  "stones_per_s": tot.value / (ms.value * 1e-3), "avg_stones": tot.value / n,
                  "written_GBps": n * 72 / (ms.value * 1e-3) / 1e9}))

# DefaultPolicy (pure MCTS, 50 searches per stone = DefaultPolicy::new()) self-play: whole games on the GPU, one thread per
# game, against the oracle restatement on all host cores (one OS thread per game like AsyncAgent) on a bounded sample.
for n_games in (() if "--quick" in sys.argv else (65536, 262144)):
    rec = (c6lib.Record * n_games)()
    st = c6lib.DpStats()
    assert L.c6_default_policy_play(50, 0xC6, 0, n_games, 0, 0, rec, None, None, C.byref(st)) == 0, L.c6_last_error()
    stones = sum(rec[g].n_moves for g in range(0, n_games, 64)) * 64
    secs = st.kernel_ms * 1e-3
    print(json.dumps({"bench": "default_policy_self_play", "games": n_games, "num_iter": 50, "kernel_ms": st.kernel_ms,
                      "games_per_s": n_games / secs, "searches_per_s": st.searches / secs,
                      "rollout_stones_per_s": st.rollout_stones / secs, "avg_stones_per_game": stones / n_games,
                      "avg_rollout_stones": st.rollout_stones / max(st.rollouts, 1)}))
if "--no-cpu" not in sys.argv:
    O = c6lib.oracle()
    cores = O.oracle_hardware_concurrency()
    outs = (c6lib.OracleDpOut * (2 * cores))()
    secs = O.oracle_default_policy_many(50, 0xC6, 0, 2 * cores, 0, cores, outs)
    print(json.dumps({"bench": "default_policy_self_play_cpu_oracle", "games": 2 * cores, "cores": cores, "seconds": secs,
                      "games_per_s": 2 * cores / secs, "searches_per_s": sum(o.searches for o in outs) / secs,
                      "rollout_stones_per_s": sum(o.rollout_stones for o in outs) / secs}))
