#!/usr/bin/env python
"""BASELINE.json configs[0]: ONE game of the AlphaZero policy with a RandomEvaluator (the reference's `cargo run -p
sample`: `light_weight()` = 2 simulations per stone, Connect6/sample/rust/src/main.rs:6-15; and the reference default,
800).  Wall time per game: the engine with one resident game (device evaluator, and through the host Callback ABI with
the sample host evaluator) next to the oracle restatement on one host core.  One JSON line per case.
A single game is one warp: this is the latency end of the design, the throughput end is bench.py."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import c6lib  # noqa: E402

L, O = c6lib.product(), c6lib.oracle()
for sims in (2, 800):
    cfg = c6lib.make_config(num_simulation=sims, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6, device=0, slots=1,
                            log2_records=18, evaluator=c6lib.EVAL_PHILOX)
    rec, st = (c6lib.Record * 1)(), c6lib.Stats()
    L.c6_self_play_device(C.byref(cfg), 0, 1, rec, C.byref(st))  # warm-up (context, allocation)
    t0 = time.perf_counter()
    assert L.c6_self_play_device(C.byref(cfg), 0, 1, rec, C.byref(st)) == 0, L.c6_last_error()
    dt = time.perf_counter() - t0
    print(json.dumps({"case": f"engine, device RandomEvaluator, {sims} simulations/stone", "seconds_per_game": dt, "stones": rec[0].n_moves,
                      "simulations": st.sims, "simulations_per_s": st.sims / dt}), flush=True)

    if sims == 2:  # where the wall time of so short a job goes: driver round trips, not the search
        tt = [time.perf_counter()]
        e = L.c6_engine_create(C.byref(cfg)); tt.append(time.perf_counter())
        assert e, L.c6_last_error()
        assert L.c6_engine_start(e, 0, 1) == 0; tt.append(time.perf_counter())
        assert L.c6_engine_run_device(e) == 0, L.c6_last_error(); tt.append(time.perf_counter())
        L.c6_engine_records(e, rec, 1); L.c6_engine_stats(e, C.byref(st)); tt.append(time.perf_counter())
        L.c6_engine_destroy(e); tt.append(time.perf_counter())
        print(json.dumps({"case": "engine, device RandomEvaluator, 2 simulations/stone: breakdown", "ms": dict(zip(
            ["create", "start", "run", "records+stats", "destroy"], [round(1e3 * (b - a), 3) for a, b in zip(tt, tt[1:])]))}), flush=True)

    # through the host Callback ABI: the reference's sequential schedule (K = 1), then the opt-in multi-leaf mode — K leaves per
    # evaluator call, so a single game needs ~K times fewer round trips per stone (not the reference's search: SURVEY 7 "virtual loss")
    for K in ((1,) if sims == 2 else (1, 2, 4, 8)):
        cfg2 = c6lib.make_config(num_simulation=sims, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6, device=0, slots=1,
                                 log2_records=18, evaluator=c6lib.EVAL_CALLBACK, virtual_loss=K)
        e = L.c6_engine_create(C.byref(cfg2))
        assert e, L.c6_last_error()
        assert L.c6_engine_start(e, 0, 1) == 0
        cb = C.cast(L.c6_sample_random_callback, c6lib.CALLBACK)
        t0 = time.perf_counter()
        assert L.c6_engine_run_callback(e, cb) == 0, L.c6_last_error()
        dt = time.perf_counter() - t0
        L.c6_engine_stats(e, C.byref(st))
        L.c6_engine_records(e, rec, 1)
        L.c6_engine_destroy(e)
        print(json.dumps({"case": f"engine, host Callback ABI (sample evaluator), {sims} simulations/stone, virtual_loss {K}", "seconds_per_game": dt,
                          "stones": rec[0].n_moves, "simulations": st.sims, "simulations_per_s": st.sims / dt}), flush=True)
    if sims == 800:
        for K in (2, 4, 8):
            cfg3 = c6lib.make_config(num_simulation=sims, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6, device=0, slots=1,
                                     log2_records=18, evaluator=c6lib.EVAL_PHILOX, virtual_loss=K)
            t0 = time.perf_counter()
            assert L.c6_self_play_device(C.byref(cfg3), 0, 1, rec, C.byref(st)) == 0, L.c6_last_error()
            dt = time.perf_counter() - t0
            print(json.dumps({"case": f"engine, device RandomEvaluator, {sims} simulations/stone, virtual_loss {K}", "seconds_per_game": dt,
                              "stones": rec[0].n_moves, "simulations": st.sims, "simulations_per_s": st.sims / dt}), flush=True)

    mm = 0 if sims == 2 else 12  # the 800-simulation oracle game is timed on its first 12 stones and scaled by simulations
    p = c6lib.OracleParams(sims, 0.25, 0.03, 1.0, 0xC6, 1, 2, 1, 1, mm)
    out = c6lib.OracleGameOut()
    t0 = time.perf_counter()
    O.oracle_self_play(None, C.byref(p), 0, C.byref(out), None, None, None)
    dt = time.perf_counter() - t0
    print(json.dumps({"case": f"oracle (reference cost model), one host core, {sims} simulations/stone" + (f", first {mm} stones" if mm else ""),
                      "seconds": dt, "stones": out.n_moves, "simulations": out.sims, "simulations_per_s": out.sims / dt}), flush=True)
