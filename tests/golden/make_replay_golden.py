#!/usr/bin/env python
"""Generates tests/golden/replay_buffer.json with the REFERENCE's own replay buffer as the checker:

  python tests/golden/make_replay_golden.py      (needs /root/reference; run in the build container)

Games of the pinned oracle are put into the form `pyconnect6.self_play` returns — (winner, [(turn, board, (row, col)), ...]) —
and pushed through `Buffer.push_game` of /root/reference/AlphaZero/py_weighted/buffer.py (imported, not copied), once with room
for everything and once with a small `max_size` (retention of the newest samples) followed by `clear_half`.  The fixture holds
the packed 240-byte records of the same games and what the reference buffer then contains; tests/test_replay.py checks
pyconnect6.replay (host and GPU feeds, PackedBuffer) against it.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, "/root/reference/AlphaZero/py_weighted")
import c6lib  # noqa: E402
from buffer import Buffer  # noqa: E402  (the reference's)


def oracle_games():
    O = c6lib.oracle()
    games = []
    for gid, mm in ((0, 0), (1, 37), (2, 0), (3, 1), (4, 12)):
        p = c6lib.OracleParams(2, 0.25, 0.03, 1.0, 0xC6, 1, c6lib.ORACLE_TIE[1], 0, 1, mm)
        out = c6lib.OracleGameOut()
        O.oracle_self_play(None, C.byref(p), gid, C.byref(out), None, None, None)
        games.append((out.winner, [out.moves[i] for i in range(out.n_moves)], [out.turns[i] for i in range(out.n_moves)]))
    return games


def nested(winner, moves, turns):
    board = [0] * 225
    path = []
    for m, t in zip(moves, turns):
        path.append((t, list(board), (m // 15, m % 15)))
        board[m] = t
    return winner, path


def packed(games):
    rec = np.zeros((len(games), 240), dtype=np.uint8)
    for g, (winner, moves, _) in enumerate(games):
        rec[g, 0] = np.uint8(winner & 0xFF)
        rec[g, 1] = len(moves)
        rec[g, 2] = 1
        rec[g, 4:8] = np.frombuffer(np.uint32(g).tobytes(), dtype=np.uint8)
        rec[g, 8:8 + len(moves)] = moves
    return rec


def dump(buf):
    # board as a 225-character string: X black (-1), . empty, O white (+1)
    return [[int(w), int(p), "".join("X.O"[int(x) + 1] for x in b), int(pos)] for (w, p, b, pos) in buf.buffer]


def main():
    games = oracle_games()
    full = Buffer(max_size=100000, board_size=15, num_sample=8)
    small = Buffer(max_size=150, board_size=15, num_sample=8)
    for g in games:
        full.push_game(nested(*g))
        small.push_game(nested(*g))
    after_push = dump(small)
    small.clear_half()
    out = {"records": packed(games).tolist(), "n_samples": len(full), "full": dump(full), "small_max_size": 150,
           "small_after_push": after_push, "small_after_clear_half": dump(small),
           "cite": "AlphaZero/py_weighted/buffer.py:22-44 (push_game, retention, clear_half), imported from /root/reference"}
    with open(os.path.join(HERE, "replay_buffer.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("replay_buffer.json:", len(full), "samples from", len(games), "games")


if __name__ == "__main__":
    main()
