#!/usr/bin/env python
"""bench.py — MCTS simulations/sec of batched Connect6 AlphaZero self-play (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload: 4,096 concurrent self-play games per GPU, AlphaZero policy, RandomEvaluator, 800
simulations per stone, reference default hyper-parameters (epsilon .25, alpha .03, c_puct 1).
A step is one stone for every resident game (= 4,096 x 800 simulations incl. move choice and node
collection).  `value` = simulations/s with everything resident in HBM (device RandomEvaluator fused
into expansion); `e2e` = the same search driven through the reference C-ABI evaluator contract
(host Callback, host buffers, D2H + H2D inside the timed region).  Games shard by game id across
GPUs with no collective on the data path (weak scaling).

The product is reached through its own ctypes binding (alphazero-connect6_b200/python/pyconnect6/_capi.py over the C-ABI);
the oracle (oracle/binding.py) is touched only by the cpu_baseline leg and by --impl reference.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "alphazero-connect6_b200", "python"))

GAMES_PER_GPU = 4096
SIMS = 800
EPS, ALPHA, CPUCT, SEED = 0.25, 0.03, 1.0, 0xC6
METRIC = "mcts_simulations_per_sec"
UNIT = "simulations/s"


def oracle_binding():
    """oracle/binding.py — only the cpu_baseline and --impl reference legs come here."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import binding
    return binding


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop = index, [], False
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


def ncu_traffic(sims_per_launch):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    workload (profiles/traffic.json, written by scripts/summarize_profile.py); None if the capture is of another shape."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if int(t["sims_per_launch"]) == int(sims_per_launch):
            return float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        pass
    return None, None


def algorithmic_bytes(st0, st1, materialised_batch):
    """16*sum C_k + 28*sum C_leaf + 16*(D+1) per simulation (+14,432 B per
    evaluated leaf only when an evaluator batch is materialised in global memory)."""
    d = {k: getattr(st1, k) - getattr(st0, k) for k in ("sims", "sum_ck", "sum_cleaf", "sum_depth", "evals")}
    b = 16 * d["sum_ck"] + 28 * d["sum_cleaf"] + 16 * (d["sum_depth"] + d["sims"])
    if materialised_batch:
        b += 14432 * d["evals"]
    return b, d


def run_reference(args, rank, world):
    """--impl reference: the restated reference (oracle, reference cost model) on the host cores."""
    if rank != 0:
        return None
    ob = oracle_binding()
    O = ob.oracle()
    cores = O.oracle_hardware_concurrency()
    moves = 2  # stones per game per step: bounded sample
    p = ob.OracleParams(SIMS, EPS, ALPHA, CPUCT, SEED, 1, 2, 1, 1, moves)
    outs = (ob.OracleGameOut * cores)()
    for w in range(max(args.warmup, 0)):
        if w >= 1:
            break  # one warm-up pass is enough for a CPU job; keeps the run within minutes
        O.oracle_self_play_many(None, C.byref(p), 10_000_000, cores, cores, outs)
    sims, secs = 0, 0.0
    for k in range(args.steps):
        secs += O.oracle_self_play_many(None, C.byref(p), 20_000_000 + k * cores, cores, cores, outs)
        sims += sum(o.sims for o in outs)
    v = sims / secs
    sample = f"{args.steps} x ({cores} games x first {moves} stones x {SIMS} sims), one OS thread per game"
    return json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, GAMES_PER_GPU),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def workload_config(n_gpus, games):
    return {"workload": f"{games} concurrent self-play games per GPU, AlphaZero policy + RandomEvaluator, 800 simulations/stone "
                        "(BASELINE.json configs[1]); reference default hyper-parameters epsilon=0.25 alpha=0.03 c_puct=1, "
                        "exact PUCT ties broken uniformly (Philox) as the reference does",
            "games_per_gpu": games, "num_simulation": SIMS, "epsilon": EPS, "dirichlet_alpha": ALPHA,
            "step": f"one stone for every resident game ({games} x 800 simulations + move choice + node collection), one kernel launch",
            "sharding": f"game id -> rank, {n_gpus} rank(s), no data-path collective",
            "l2": "working set >> L2 (node pool of 4096 games is tens of GB); no flush needed"}


class StdoutToStderr:
    """Everything written to fd 1 while the benchmark runs (NCCL's version banner, library chatter) goes to
    stderr, so that stdout carries exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    with StdoutToStderr():
        line = run()
    if line is not None:
        print(line, flush=True)


class Reducer:
    """max / sum over ranks on the device (NCCL) — the only exchange of the benchmark: timers and counters."""

    def __init__(self, dist):
        self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def max(self, x):
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, *xs):
        if self.dist is None:
            return [float(x) for x in xs]
        import torch
        t = torch.tensor([float(x) for x in xs], device="cuda", dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]


def device_leg(L, capi, red, args, rank, local, games, steps, warmup, virtual_loss=1, steady=False):
    """`steps` stones for `games` resident games with everything in HBM; returns the metric dict."""
    cfg = capi.make_config(num_simulation=SIMS, epsilon=EPS, dirichlet_alpha=ALPHA, c_puct=CPUCT, seed=SEED, device=local,
                           slots=games, log2_records=args.log2_records, evaluator=capi.EVAL_PHILOX, virtual_loss=virtual_loss)
    e = L.c6_engine_create(C.byref(cfg))
    if not e:
        raise RuntimeError(L.c6_last_error().decode())
    try:
        # many more games queued than will finish: slots refill, the batch stays full
        assert L.c6_engine_start(e, rank * 100_000_000, 50_000_000) == 0, L.c6_last_error()
        ms = C.c_float()
        assert L.c6_engine_timed_steps(e, warmup, C.byref(ms)) == 0, L.c6_last_error()
        st0, st1 = capi.Stats(), capi.Stats()
        L.c6_engine_stats(e, C.byref(st0))
        km0, kl0, km1, kl1 = C.c_double(), C.c_longlong(), C.c_double(), C.c_longlong()
        L.c6_engine_kernel_times(e, C.byref(km0), C.byref(kl0))
        red.barrier()
        with ClockSampler(local) as clk:
            assert L.c6_engine_timed_steps(e, steps, C.byref(ms)) == 0, L.c6_last_error()
        L.c6_engine_stats(e, C.byref(st1))
        L.c6_engine_kernel_times(e, C.byref(km1), C.byref(kl1))
        local_ms = float(ms.value)
        elapsed_ms = red.max(local_ms)
        total_sims, total_launches = red.sum(st1.sims - st0.sims, st1.kernel_launches - st0.kernel_launches)
        # roofline of the dominant kernel (one launch per step) on this rank: algorithmic bytes of the launches in the timed
        # region / their own durations (one CUDA event pair per launch on the engine stream)
        bytes_alg, d = algorithmic_bytes(st0, st1, materialised_batch=False)
        peak, peak_src = measured_peak_gbs()
        k_launches = max(kl1.value - kl0.value, 1)
        k_ms = (km1.value - km0.value) / k_launches
        achieved = (bytes_alg / k_launches) / (k_ms * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(d["sims"] / k_launches) if virtual_loss == 1 else (None, None)
        out = {"value": total_sims / (elapsed_ms * 1e-3), "ms_per_step": elapsed_ms / steps, "gpu_launches": int(total_launches),
               "clocks": clk.summary(), "peak_records_per_game": int(st1.peak_rec), "bytes_per_game": int(st1.bytes_per_game),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                            "traffic_source": traffic_src,
                            "kernel": "k_sims_device<philox,noise> (1 launch/step: 800 searches per game, then the stone and node collection)",
                            "kernel_ms_per_launch": k_ms, "kernel_share_of_step": k_ms * k_launches / local_ms,
                            "algorithmic_bytes_per_launch": bytes_alg / k_launches, "peak_source": peak_src,
                            "algorithmic_bytes_per_sim": bytes_alg / max(d["sims"], 1),
                            "avg_depth": d["sum_depth"] / max(d["sims"], 1), "avg_sum_ck": d["sum_ck"] / max(d["sims"], 1),
                            "avg_cleaf": d["sum_cleaf"] / max(d["evals"], 1)}}
        # steady state: keep stepping until every slot has finished at least one game (slots refill from the queue, so
        # the resident games are then spread over all phases of a game), then time a window: simulations/s over whole
        # games and finished games/s (the second half of BASELINE.json's metric)
        if steady:
            run_in, window = 240, 60
            assert L.c6_engine_timed_steps(e, run_in, C.byref(ms)) == 0, L.c6_last_error()
            sa, sb = capi.Stats(), capi.Stats()
            L.c6_engine_stats(e, C.byref(sa))
            red.barrier()
            assert L.c6_engine_timed_steps(e, window, C.byref(ms)) == 0, L.c6_last_error()
            L.c6_engine_stats(e, C.byref(sb))
            w_ms = red.max(float(ms.value))
            w_sims, w_games = red.sum(sb.sims - sa.sims, sb.games_done - sa.games_done)
            out["steady_state"] = {"simulations_per_sec": w_sims / (w_ms * 1e-3), "games_per_sec": w_games / (w_ms * 1e-3),
                                   "stones_per_game": (w_sims / SIMS) / max(w_games, 1.0), "ms_per_step": w_ms / window,
                                   "window": f"{window} steps after {warmup + steps + run_in} (every slot refilled at least once); "
                                             f"{int(w_games)} games finished in the window"}
        return out
    finally:
        L.c6_engine_destroy(e)


def e2e_leg(L, capi, red, args, rank, local, games, callback_name, waves_w=8, waves=48):
    """`waves` simulations for each of `games` games through the reference-facing Callback contract: host buffers, D2H of the
    leaf boards and H2D of the evaluator's answer inside the timed region."""
    cfg = capi.make_config(num_simulation=SIMS, epsilon=EPS, dirichlet_alpha=ALPHA, c_puct=CPUCT, seed=SEED, device=local,
                           slots=games, log2_records=args.log2_records, evaluator=capi.EVAL_CALLBACK)
    e = L.c6_engine_create(C.byref(cfg))
    if not e:
        raise RuntimeError(L.c6_last_error().decode())
    try:
        assert L.c6_engine_start(e, rank * 100_000_000, 50_000_000) == 0
        cb = C.cast(getattr(L, callback_name), capi.CALLBACK)
        ms, h2d, d2h = C.c_float(), C.c_double(), C.c_double()
        assert L.c6_engine_timed_waves(e, cb, waves_w, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
        s0, s1 = capi.Stats(), capi.Stats()
        L.c6_engine_stats(e, C.byref(s0))
        red.barrier()
        hw0, hc0, hw1, hc1 = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        L.c6_engine_host_times(e, C.byref(hw0), C.byref(hc0))
        bw0 = L.c6_engine_host_blocked_waits(e)
        assert L.c6_engine_timed_waves(e, cb, waves, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
        L.c6_engine_stats(e, C.byref(s1))
        L.c6_engine_host_times(e, C.byref(hw1), C.byref(hc1))
        bw1 = L.c6_engine_host_blocked_waits(e)
        e2e_ms = red.max(float(ms.value))
        (e2e_sims,) = red.sum(s1.sims - s0.sims)
        # what the host link of THIS box carries (rank 0's GPU, nothing else running): the floor under ms_per_wave
        up, down, both = C.c_double(), C.c_double(), C.c_double()
        per_wave = int(h2d.value / waves)
        red.barrier()
        link = None
        if rank == 0 and L.c6_probe_pcie(per_wave, 8, local, C.byref(up), C.byref(down), C.byref(both)) == 0:
            link = {"h2d_GBps": up.value, "d2h_GBps": down.value, "each_way_GBps_when_both_run": both.value,
                    "floor_ms_per_wave": per_wave / both.value / 1e6,
                    "note": "pinned-memory copies of one wave's bytes alone on rank 0's link, measured after the timed waves"}
        red.barrier()
        return {"value": e2e_sims / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d.value / waves, "d2h_bytes_per_step": d2h.value / waves,
                "step": f"one wave = one simulation for each of {games} games through the C-ABI Callback (host buffers, "
                        f"{callback_name}); {waves} waves timed",
                "ms_per_wave": e2e_ms / waves,
                "host_ms_per_wave": {"in_callback": 1e3 * (hc1.value - hc0.value) / waves, "blocked_on_gpu": 1e3 * (hw1.value - hw0.value) / waves,
                                     "waits_that_slept_per_wave": (bw1 - bw0) / waves, "of": "rank 0"},
                "host_link": link}
    finally:
        L.c6_engine_destroy(e)


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-steady", action="store_true", help="skip the whole-game steady-state window (games/s)")
    ap.add_argument("--no-config4", action="store_true", help="skip the 8,192-games-per-GPU leg of multi-GPU runs (BASELINE configs[3])")
    ap.add_argument("--virtual-loss", type=int, default=1, help="opt-in multi-leaf mode (NOT the reference's search); adds the key `multi_leaf`")
    ap.add_argument("--log2-records", type=int, default=18)
    ap.add_argument("--epsilon", type=float, default=None, help="override epsilon (0 disables Dirichlet noise)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    global EPS
    if args.epsilon is not None:
        EPS = args.epsilon
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3

    from pyconnect6 import _capi as capi
    L = capi.product()  # fails loudly when the CUDA library is not built
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    red = Reducer(dist)
    # several ranks share the host on the callback legs: keep each rank's host threads next to its GPU
    numa_node = L.c6_bind_host_to_device_numa(local) if world > 1 and os.environ.get("C6_NUMA_BIND", "1") != "0" else -1
    host = {"cores_per_rank": L.c6_host_cores_per_rank(), "spin_budget_us": L.c6_host_spin_budget_us(), "numa_node": numa_node,
            "cpus": os.cpu_count()}

    games = args.games
    dev = device_leg(L, capi, red, args, rank, local, games, args.steps, args.warmup, steady=not args.no_steady)
    out = {"metric": METRIC, "value": dev["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": dev["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": workload_config(world, games), "roofline": dev["roofline"],
           "gpu_launches": dev["gpu_launches"], "clocks": dev["clocks"], "peak_records_per_game": dev["peak_records_per_game"],
           "bytes_per_game": dev["bytes_per_game"], "host": host}
    if "steady_state" in dev:
        out["steady_state"] = dev["steady_state"]

    # BASELINE configs[3]: 65,536 games over 8 GPUs = 8,192 resident games per GPU (16,384 / 32,768 at 2 / 4 GPUs)
    if world > 1 and not args.no_config4 and games != 8192:
        c4 = device_leg(L, capi, red, args, rank, local, 8192, max(args.steps // 2, 3), 3)
        out["config4"] = {"workload": f"{8192 * world} concurrent games, 8192 per GPU, game-sharded over {world} GPUs, Dirichlet noise on",
                          "value": c4["value"], "unit": UNIT, "ms_per_step": c4["ms_per_step"], "roofline_frac": c4["roofline"]["frac"],
                          "peak_records_per_game": c4["peak_records_per_game"]}
    if args.virtual_loss > 1:
        vl = device_leg(L, capi, red, args, rank, local, games, max(args.steps // 2, 3), 3, virtual_loss=args.virtual_loss)
        out["multi_leaf"] = {"virtual_loss": args.virtual_loss, "value": vl["value"], "unit": UNIT, "ms_per_step": vl["ms_per_step"],
                             "note": "opt-in mode, not the reference's sequential search: up to K leaves per game and wave"}

    # end to end through the reference-facing evaluator contract: host callback, host buffers.  Headline: the write-only random
    # evaluator, the twin of the reference's own sample callback (Connect6/sample/cpp/main.cpp:6-28 fills values and policies and
    # never reads the boards); alongside it an evaluator that reads every board before overwriting it.
    if not args.no_e2e:
        out["e2e"] = e2e_leg(L, capi, red, args, rank, local, games, "c6_sample_random_callback")
        out["e2e_reading_evaluator"] = e2e_leg(L, capi, red, args, rank, local, games, "c6_sample_reading_callback", waves_w=4, waves=24)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ob = oracle_binding()
        O = ob.oracle()
        cores = O.oracle_hardware_concurrency()
        moves = 30  # ~10 s on every host core (the sample the task statement asks for: 10-30 s of CPU work)
        p = ob.OracleParams(SIMS, EPS, ALPHA, CPUCT, SEED, 1, 2, 1, 1, moves)
        outs = (ob.OracleGameOut * cores)()
        secs = O.oracle_self_play_many(None, C.byref(p), 30_000_000, cores, cores, outs)
        csims = sum(o.sims for o in outs)
        out["cpu_baseline"] = {"value": csims / secs, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{cores} games x first {moves} stones x {SIMS} sims, oracle in the reference's cost model "
                                         f"(SipHash board keys, full nodes), one OS thread per game like AsyncAgent; {secs:.1f} s"}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return json.dumps(out) if rank == 0 else None


if __name__ == "__main__":
    main()
