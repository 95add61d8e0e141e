"""CPU suite, part 3: the N>1 path (game-index sharding + record gather) on world_size-2 gloo.

The shard producer is injected: here a stub that fabricates records as a pure function of the game id
(the GPU engine has exactly that property: Philox streams are keyed by game id), so the host logic —
block partition, padding, gather to rank 0 in game-id order — is covered without a GPU.  The GPU suite
repeats the check with the real engine (test_parity_gpu.py::test_sharded_equals_unsharded).
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import c6lib  # noqa: F401  (puts the package on sys.path)
from pyconnect6 import sharded


def fake_records(first_id, n, **kw):
    out = np.zeros((n, sharded.RECORD_BYTES), dtype=np.uint8)
    for i in range(n):
        gid = first_id + i
        rng = np.random.default_rng(gid)
        m = int(rng.integers(11, 60))
        out[i, 0] = np.uint8(np.int8([-1, 0, 1][gid % 3]))
        out[i, 1] = m
        out[i, 2] = 1
        out[i, 4:8] = np.frombuffer(np.uint32(gid).tobytes(), dtype=np.uint8)
        out[i, 8:8 + m] = rng.permutation(225)[:m].astype(np.uint8)
    return out, None


def test_shard_range_tiles_exactly():
    for n in (0, 1, 2, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharded.shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1
    with pytest.raises(ValueError):
        sharded.shard_range(4, 2, 2)


def _worker(rank, world, port, n_games, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rec, _ = sharded.self_play_sharded(n_games, dist=dist, produce=fake_records)
        if rank == 0:
            q.put(rec)
        else:
            assert rec is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n_games", [9, 2, 1])
def test_two_rank_gather_equals_single_process(n_games):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_games, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    want, _ = fake_records(0, n_games)
    assert got.shape == want.shape and np.array_equal(got, want)
    recs = sharded.records_from_bytes(got)
    assert [r["game_id"] for r in recs] == list(range(n_games))


def test_single_process_path_needs_no_process_group():
    rec, _ = sharded.self_play_sharded(5, dist=None, produce=fake_records)
    assert np.array_equal(rec, fake_records(0, 5)[0])
