#!/usr/bin/env python
"""BASELINE.json configs[4]: the evaluator is a random-init conv policy/value net shaped like the reference's
libtorch trainer net (AlphaZero/cpp_weighted/model.hpp:9-95, 10 blocks x 256 channels), fed with 8,192 leaves per wave
(= 65,536 symmetric boards).  Two ways in: the reference `Callback` ABI (host pointers: the boards cross PCIe to the
host, the callback copies them to the GPU for the net and the answers back, the engine copies them to the GPU again) and
the device-pointer hook (the net reads the batch where it lies in HBM).  Prints one JSON line per way with the evaluator
time broken out.  One GPU, or one rank per GPU under torchrun (8 x 8,192 leaves = 65,536 leaves per wave as configs[4] is
written): every rank owns its games, its engine and its replica of the net; the only exchange is the reduction of the timers.
  python scripts/config5.py [--games 8192] [--waves 2] [--blocks 10] [--dtype fp32|tf32|bf16]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 scripts/config5.py
fp32 (the reference net's own arithmetic, TF32 off) is the default; tf32 / bf16 are what a deployment would run."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import c6lib  # noqa: E402
from pyconnect6 import torch_bridge  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--games", type=int, default=8192)
ap.add_argument("--waves", type=int, default=2)
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--blocks", type=int, default=10)
ap.add_argument("--dtype", default="fp32", choices=["fp32", "tf32", "bf16"])
ap.add_argument("--chunk", type=int, default=16384)
a = ap.parse_args()

RANK, WORLD, LOCAL = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
L = c6lib.product()
dev = torch.device("cuda", LOCAL)
torch.cuda.set_device(dev)
dist = None
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = a.dtype == "tf32"
torch.manual_seed(0)


def over_ranks(seconds, sims):
    """whole-job figures: time = max over ranks, simulations = sum over ranks"""
    if dist is None:
        return seconds, sims
    t = torch.tensor([seconds], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = torch.tensor([float(sims)], device=dev, dtype=torch.float64)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(t.item()), float(n.item())


def emit(d):
    if RANK == 0:
        print(json.dumps(d), flush=True)


net = torch_bridge.WeightedPolicy(num_block=a.blocks).to(dev).eval().to(memory_format=torch.channels_last)
amp = torch.bfloat16 if a.dtype == "bf16" else None
CELLS = 225


@torch.no_grad()
def run_net(player, boards):  # boards: (n, 225) float32 on the GPU -> values (n,), policies (n, 225)
    outs_v, outs_p = [], []
    for lo in range(0, boards.shape[0], a.chunk):
        b = boards[lo:lo + a.chunk]
        p = torch.full((b.shape[0],), float(player), device=dev)
        with torch.autocast("cuda", dtype=amp, enabled=amp is not None):
            pol, val = net(p, b)
        outs_v.append(val.reshape(-1).float()), outs_p.append(pol.float())
    return torch.cat(outs_v), torch.cat(outs_p)


def make_cfg(evaluator):
    return c6lib.make_config(num_simulation=800, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6, device=LOCAL, slots=a.games,
                             log2_records=18, evaluator=evaluator)


def net_alone_ms():
    """the net on one wave's worth of boards (two calls: Black leaves, White leaves), nothing else on the GPU"""
    b = torch.zeros((4 * a.games, CELLS), device=dev)
    run_net(-1, b)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    run_net(-1, b), run_net(1, b)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e)


NET_MS = net_alone_ms()


# ---- device-pointer hook ---------------------------------------------------------------------------------------------
def dcb(player, d_values, d_policies, length, stream, user):
    with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
        pol = torch_bridge.device_tensor(d_policies, (length, CELLS), dev)
        val = torch_bridge.device_tensor(d_values, (length,), dev)
        v, p = run_net(player, pol)
        val.copy_(v), pol.copy_(p)


e1 = L.c6_engine_create(C.byref(make_cfg(c6lib.EVAL_DEVICE_CALLBACK)))
assert e1, L.c6_last_error()
assert L.c6_engine_start(e1, RANK * 100_000_000, 50_000_000) == 0
ccb = c6lib.DEVICE_CALLBACK(dcb)
assert L.c6_engine_waves_device_callback(e1, ccb, None, a.warm) == 0, L.c6_last_error()
L.c6_engine_sync(e1)
s0, s1 = c6lib.Stats(), c6lib.Stats()
L.c6_engine_stats(e1, C.byref(s0))
if dist is not None:
    dist.barrier()
t0 = time.perf_counter()
assert L.c6_engine_waves_device_callback(e1, ccb, None, a.waves) == 0, L.c6_last_error()
L.c6_engine_sync(e1)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
L.c6_engine_stats(e1, C.byref(s1))
dt, sims = over_ranks(dt, s1.sims - s0.sims)
emit({"config": f"configs[4], {WORLD} GPU(s), device-pointer evaluator hook", "n_gpus": WORLD, "games_per_gpu": a.games, "waves": a.waves,
      "net": f"{a.blocks} blocks x 256 ch, {a.dtype}, channels_last, random init", "boards_per_wave_per_gpu": 8 * a.games,
      "simulations_per_s": sims / dt, "ms_per_wave": 1e3 * dt / a.waves,
      "net_alone_ms_per_wave": NET_MS, "everything_else_ms_per_wave": 1e3 * dt / a.waves - NET_MS, "h2d_d2h_bytes": 0})
L.c6_engine_destroy(e1)

# ---- reference Callback ABI (host pointers) --------------------------------------------------------------------------
host_s = [0.0]


def hcb(player, values_p, policies_p, n):
    t = time.perf_counter()
    pol = np.ctypeslib.as_array(policies_p, shape=(n, CELLS))
    val = np.ctypeslib.as_array(values_p, shape=(n,))
    b = torch.from_numpy(pol).to(dev, non_blocking=False)
    v, p = run_net(int(player), b)
    pol[:] = p.cpu().numpy()
    val[:] = v.cpu().numpy()
    host_s[0] += time.perf_counter() - t


e2 = L.c6_engine_create(C.byref(make_cfg(c6lib.EVAL_CALLBACK)))
assert e2, L.c6_last_error()
assert L.c6_engine_start(e2, RANK * 100_000_000, 50_000_000) == 0
cb = c6lib.CALLBACK(hcb)
ms, h2d, d2h = C.c_float(), C.c_double(), C.c_double()
assert L.c6_engine_timed_waves(e2, cb, a.warm, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
host_s[0] = 0.0
L.c6_engine_stats(e2, C.byref(s0))
if dist is not None:
    dist.barrier()
assert L.c6_engine_timed_waves(e2, cb, a.waves, C.byref(ms), C.byref(h2d), C.byref(d2h)) == 0, L.c6_last_error()
L.c6_engine_stats(e2, C.byref(s1))
dt, sims = over_ranks(ms.value * 1e-3, s1.sims - s0.sims)
emit({"config": f"configs[4], {WORLD} GPU(s), reference Callback ABI (host buffers)", "n_gpus": WORLD, "games_per_gpu": a.games, "waves": a.waves,
      "net": f"{a.blocks} blocks x 256 ch, {a.dtype}, channels_last, random init", "boards_per_wave_per_gpu": 8 * a.games,
      "simulations_per_s": sims / dt, "ms_per_wave": 1e3 * dt / a.waves,
      "callback_ms_per_wave": 1e3 * host_s[0] / a.waves, "net_alone_ms_per_wave": NET_MS,
      "engine_h2d_MB_per_wave": h2d.value / a.waves / 1e6, "engine_d2h_MB_per_wave": d2h.value / a.waves / 1e6})
L.c6_engine_destroy(e2)
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
