"""Packed records -> training tuples (the replay-buffer feed next to the hot path), checked against what the REFERENCE's own
replay buffer holds after `push_game` of the same games (tests/golden/replay_buffer.json, made by make_replay_golden.py, which
imports /root/reference/AlphaZero/py_weighted/buffer.py).  The GPU feed (c6_replay.cu) is checked in the `gpu` tests below."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import c6lib  # noqa: F401
from pyconnect6 import replay

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay_buffer.json")))


def decode(samples):
    return [(w, p, ["X.O".index(ch) - 1 for ch in b], pos) for (w, p, b, pos) in samples]


def gold_records():
    return np.array(GOLD["records"], dtype=np.uint8)


def test_training_tuples_equal_the_reference_buffer():
    want = decode(GOLD["full"])
    winners, players, boards, positions = replay.training_tuples(gold_records())
    assert len(winners) == GOLD["n_samples"] == len(want)
    for i, (win, player, board, pos) in enumerate(want):
        assert (winners[i], players[i], positions[i]) == (win, player, pos), i
        assert boards[i].tolist() == board, i


def test_as_game_results_feed_the_reference_buffer_format():
    flat = []
    for win, path in replay.as_game_results(gold_records()):  # Buffer.push_game (py_weighted/buffer.py:22-31)
        for player, board, (row, col) in path:
            flat.append((win, player, board, row * 15 + col))
    assert flat == decode(GOLD["full"])


def test_packed_buffer_keeps_what_the_reference_buffer_keeps():
    """Same retention rule (newest max_size samples) and clear_half as the reference Buffer, from packed records and from the
    nested form."""
    for feed in ("records", "nested"):
        buf = replay.PackedBuffer(GOLD["small_max_size"], 15, 8)
        if feed == "records":
            buf.push_records(gold_records())
        else:
            for g in replay.as_game_results(gold_records()):
                buf.push_game(g)
        assert len(buf) == len(GOLD["small_after_push"]) == GOLD["small_max_size"]
        assert buf.contents() == decode(GOLD["small_after_push"])
        buf.clear_half()
        assert buf.contents() == decode(GOLD["small_after_clear_half"])
        np.random.seed(0)
        wins, players, boards, poses = buf.sample()
        live = decode(GOLD["small_after_clear_half"])
        assert len(wins) == 8 and all((w, p, b, q) in live for w, p, b, q in zip(wins, players, boards, poses))


def test_empty_and_single_stone_records():
    rec = np.zeros((2, 240), dtype=np.uint8)
    rec[1, 1] = 1
    rec[1, 8] = 112
    w, p, b, pos = replay.training_tuples(rec)
    assert len(w) == 1 and p[0] == -1 and pos[0] == 112 and not b[0].any()
    assert replay.as_game_results(rec)[0] == (0, [])


def test_stone_colours_follow_the_turn_rule():
    assert replay.stone_colours(7).tolist() == [-1, 1, 1, -1, -1, 1, 1]


@pytest.mark.gpu
def test_gpu_feed_equals_the_reference_buffer():
    """c6_expand_records_host (one warp per game on the GPU): the four float32 arrays of the reference trainer's sample() layout."""
    L = c6lib.product()
    rec = gold_records()
    n = rec.shape[0]
    rows = GOLD["n_samples"]
    FP = C.POINTER(C.c_float)
    w, p, pos = (np.full(rows, np.nan, dtype=np.float32) for _ in range(3))
    b = np.full((rows, 225), np.nan, dtype=np.float32)
    n_rows = C.c_longlong()
    crec = (c6lib.Record * n).from_buffer_copy(rec.tobytes())
    assert L.c6_expand_records_host(crec, n, 0, w.ctypes.data_as(FP), p.ctypes.data_as(FP), b.ctypes.data_as(FP), pos.ctypes.data_as(FP),
                                    C.byref(n_rows)) == 0, L.c6_last_error()
    assert n_rows.value == rows
    want = decode(GOLD["full"])
    assert w.tolist() == [float(x[0]) for x in want] and p.tolist() == [float(x[1]) for x in want] and pos.tolist() == [float(x[3]) for x in want]
    assert np.array_equal(b, np.array([x[2] for x in want], dtype=np.float32))


@pytest.mark.gpu
def test_gpu_feed_from_engine_records_at_scale():
    """Records of 2,048 engine games expanded in HBM for a torch trainer (no host copy of the boards) == the host feed."""
    import torch
    L = c6lib.product()
    n = 2048
    cfg = c6lib.make_config(num_simulation=2, epsilon=0.25, seed=3, evaluator=c6lib.EVAL_PHILOX, slots=1024, log2_records=12)
    rec = (c6lib.Record * n)()
    assert L.c6_self_play_device(C.byref(cfg), 0, n, rec, None) == 0, L.c6_last_error()
    raw = np.frombuffer(bytes(rec), dtype=np.uint8).reshape(n, 240)
    w, p, b, pos = replay.training_tuples_device(torch.from_numpy(raw.copy()).cuda())
    hw, hp, hb, hpos = replay.training_tuples(raw)
    assert b.is_cuda and b.dtype == torch.float32 and b.shape == (len(hw), 225)
    assert np.array_equal(w.cpu().numpy(), hw.astype(np.float32)) and np.array_equal(p.cpu().numpy(), hp.astype(np.float32))
    assert np.array_equal(pos.cpu().numpy(), hpos.astype(np.float32)) and np.array_equal(b.cpu().numpy(), hb.astype(np.float32))
    # empty input
    w0, _, b0, _ = replay.training_tuples_device(np.zeros((0, 240), dtype=np.uint8))
    assert w0.numel() == 0 and b0.shape == (0, 225)
