"""Game-index sharding of self-play across the GPUs of one box (one process per GPU).

The reference's only parallelism is one game per task on a thread pool
(libconnect6/src/agent/async_agent/mod.rs:100-137); games share nothing, so here they shard by game
id across ranks with NO collective on the data path.  A game's search is a pure function of
(seed, game id) — Philox streams are keyed by the id — so the records do not depend on how many
ranks played them.  The only exchange is the optional gather of finished packed records (240 B per
game) to rank 0 over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import ctypes as C

import numpy as np

from . import _capi

RECORD_BYTES = C.sizeof(_capi.Record)  # 240


def shard_range(n_games, rank, world):
    """Contiguous block of game ids for `rank`: sizes differ by at most one, blocks tile [0, n_games)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} of {world}")
    base, extra = divmod(int(n_games), world)
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def play_shard_device(first_id, n_games, *, num_simulation=800, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=0xC6,
                      device=0, slots=None, log2_records=18, evaluator=_capi.EVAL_PHILOX, max_moves=0):
    """`n_games` games with ids first_id.. on `device` with a device-resident evaluator; returns the
    packed records as a (n_games, 240) uint8 array and the engine statistics."""
    L = _capi.product()
    rec = (_capi.Record * max(n_games, 1))()
    st = _capi.Stats()
    if n_games > 0:
        cfg = _capi.make_config(num_simulation=num_simulation, epsilon=epsilon, dirichlet_alpha=dirichlet_alpha, c_puct=c_puct,
                                seed=seed, device=device, slots=min(slots or n_games, n_games), log2_records=log2_records,
                                evaluator=evaluator, max_moves=max_moves)
        if L.c6_self_play_device(C.byref(cfg), first_id, n_games, rec, C.byref(st)) != 0:
            raise RuntimeError(L.c6_last_error().decode())
    out = np.frombuffer(rec, dtype=np.uint8).reshape(-1, RECORD_BYTES)[:n_games].copy()
    return out, st


def gather_records(local, n_games, dist, device=None):
    """Gather every rank's (n_local, 240) uint8 record block to rank 0, in game-id order.

    Shards are padded to the largest block so one `gather` moves everything (65,536 games x 240 B =
    15.7 MB: microseconds over NVLink, so this is never on the critical path).  Returns the
    (n_games, 240) array on rank 0 and None elsewhere."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [shard_range(n_games, r, world)[1] for r in range(world)]
    cap = max(max(sizes), 1)
    buf = torch.zeros((cap, RECORD_BYTES), dtype=torch.uint8)
    if len(local):
        buf[:len(local)] = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        buf = buf.to(device)
    parts = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, parts, dst=0)
    if rank != 0:
        return None
    return np.concatenate([parts[r][:sizes[r]].cpu().numpy() for r in range(world)], axis=0)


def self_play_sharded(n_games, dist=None, produce=play_shard_device, gather=True, device=None, **kw):
    """Play `n_games` games split over the ranks of `dist` (None = single process).

    `produce(first_id, n, **kw) -> ((n, 240) uint8, stats)` runs one shard — the GPU engine by default;
    the CPU tests inject a stub so the sharding / gather logic is exercised without a GPU."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        rec, st = produce(0, n_games, **kw)
        return rec, st
    lo, n = shard_range(n_games, dist.get_rank(), dist.get_world_size())
    rec, st = produce(lo, n, **kw)
    if not gather:
        return rec, st
    return gather_records(rec, n_games, dist, device=device), st


def records_from_bytes(arr):
    """(n, 240) uint8 -> list of dicts {winner, n_moves, finished, game_id, moves}."""
    out = []
    for row in np.asarray(arr, dtype=np.uint8).reshape(-1, RECORD_BYTES):
        r = _capi.Record.from_buffer_copy(row.tobytes())
        out.append({"winner": int(r.winner), "n_moves": int(r.n_moves), "finished": int(r.finished), "game_id": int(r.game_id),
                    "moves": [int(r.moves[i]) for i in range(r.n_moves)]})
    return out
