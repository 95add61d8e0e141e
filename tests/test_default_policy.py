"""DefaultPolicy (the reference's pure MCTS with random rollouts, LC/policy/default_policy/mod.rs:94-271) on the GPU
against the CPU oracle restatement (oracle/c6_default_policy.hpp) on the same Philox streams: every stone, the winner,
the root node's visit / black_win counters at every move and the search statistics must be identical (integer work:
bit-exact).  The oracle itself is pinned by ports of the reference's five default_policy tests (oracle_tests.cpp).
"""
import ctypes as C

import numpy as np
import pytest

import c6lib

STAT_NAMES = ("searches", "rollouts", "rollout_stones", "select_steps", "children_scanned", "replaced", "nodes")


def oracle_games(num_iter, seed, first_id, n, max_moves=0):
    O = c6lib.oracle()
    outs = (c6lib.OracleDpOut * n)()
    O.oracle_default_policy_many(num_iter, seed, first_id, n, max_moves, 0, outs)
    return outs


def test_oracle_default_policy_properties():
    """CPU: Agent::play with DefaultPolicy on the oracle — the assertions of default_policy/tests.rs:134-150 plus
    determinism of the Philox seam."""
    a = oracle_games(10, 0xC6, 0, 4)
    b = oracle_games(10, 0xC6, 0, 4)
    for x, y in zip(a, b):
        assert list(x.moves[:x.n_moves]) == list(y.moves[:y.n_moves]) and x.winner == y.winner
        assert x.n_moves > 0 and len(set(x.moves[:x.n_moves])) == x.n_moves  # legal: no cell twice
        assert x.searches == 10 * x.n_moves
        assert x.nodes <= x.rollouts + x.n_moves + 1
        if x.winner != 0:  # the last mover is the winner (tests.rs:143-147)
            last = x.n_moves - 1
            turn = -1 if last == 0 else (1 if ((last - 1) // 2) % 2 == 0 else -1)
            assert turn == x.winner
    assert any(list(a[0].moves[:a[0].n_moves]) != list(a[i].moves[:a[i].n_moves]) for i in range(1, 4))  # game id matters


@pytest.mark.gpu
@pytest.mark.parametrize("num_iter,n_games,max_moves,seed,slots", [(50, 48, 0, 0xC6, 0), (2, 40, 0, 7, 0), (200, 8, 30, 11, 0), (1, 16, 0, 1, 0),
                                                                 (50, 48, 0, 0xC6, 5), (3, 150, 0, 9, 37)])
def test_default_policy_gpu_matches_oracle(num_iter, n_games, max_moves, seed, slots, monkeypatch):
    """slots > 0: fewer tree slots than games (C6_DP_SLOTS) — a lane whose game is over takes the next game into the same slot,
    on a tree it has to empty first; 5 and 37 slots also leave a warp partly filled."""
    if slots:
        monkeypatch.setenv("C6_DP_SLOTS", str(slots))
    L = c6lib.product()
    first = 1000
    rec = (c6lib.Record * n_games)()
    rv = np.zeros((n_games, 225), dtype=np.int32)
    rb = np.zeros((n_games, 225), dtype=np.int32)
    st = c6lib.DpStats()
    rc = L.c6_default_policy_play(num_iter, seed, first, n_games, max_moves, 0, rec, rv.ctypes.data_as(C.POINTER(C.c_int32)),
                                  rb.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(st))
    assert rc == 0, L.c6_last_error()
    want = oracle_games(num_iter, seed, first, n_games, max_moves)
    for g in range(n_games):
        w = want[g]
        assert rec[g].game_id == first + g
        assert [rec[g].moves[i] for i in range(rec[g].n_moves)] == list(w.moves[:w.n_moves]), f"game {g}"
        assert rec[g].winner == w.winner
        assert list(rv[g, :w.n_moves]) == list(w.root_visit[:w.n_moves]), f"game {g} root visits"
        assert list(rb[g, :w.n_moves]) == list(w.root_black_win[:w.n_moves]), f"game {g} root black wins"
    for name in STAT_NAMES:
        assert getattr(st, name) == sum(getattr(w, name) for w in want), name
    assert st.errors == 0


@pytest.mark.gpu
def test_default_policy_gpu_at_scale_properties():
    """Size-independent properties at a size the oracle would need minutes for: every game legal, finished, won by its
    last mover (or a 225-stone draw), searches = num_iter x stones, and the run is a pure function of (seed, game id):
    a sub-range replays identically."""
    L = c6lib.product()
    n, it = 8192, 50
    rec = (c6lib.Record * n)()
    st = c6lib.DpStats()
    assert L.c6_default_policy_play(it, 0xC6, 0, n, 0, 0, rec, None, None, C.byref(st)) == 0, L.c6_last_error()
    stones = 0
    for g in range(n):
        r = rec[g]
        mv = [r.moves[i] for i in range(r.n_moves)]
        assert r.finished == 1 and len(set(mv)) == len(mv) and max(mv) < 225
        stones += r.n_moves
        if r.winner == 0:
            assert r.n_moves == 225
        else:
            last = r.n_moves - 1
            assert r.winner == (-1 if last == 0 else (1 if ((last - 1) // 2) % 2 == 0 else -1))
    assert st.searches == it * stones and st.rollouts <= st.searches and st.errors == 0
    sub = (c6lib.Record * 64)()
    assert L.c6_default_policy_play(it, 0xC6, 4000, 64, 0, 0, sub, None, None, None) == 0
    for i in range(64):
        assert bytes(sub[i].moves) == bytes(rec[4000 + i].moves) and sub[i].winner == rec[4000 + i].winner


@pytest.mark.gpu
def test_pyconnect6_default_policy_play():
    """Python surface: records expanded to the reference's (winner, [(turn, board, (row, col))]) shape."""
    from pyconnect6 import pyconnect6 as native
    games = native.default_policy_play(num_iter=10, num_game=3, seed=0xC6, first_game_id=1000)
    want = oracle_games(10, 0xC6, 1000, 3)
    assert len(games) == 3
    for (winner, path), w in zip(games, want):
        assert winner == w.winner and [r * 15 + c for _, _, (r, c) in path] == list(w.moves[:w.n_moves])
        board = [0] * 225
        for turn, b, (r, c) in path:  # Path.board is the board BEFORE the stone (agent_impl/mod.rs:139-143)
            assert b == board and b[r * 15 + c] == 0
            board[r * 15 + c] = turn


@pytest.mark.gpu
def test_default_policy_rejects_budgets_whose_node_ids_do_not_fit():
    """child ids are packed in 24 bits; a budget that could exceed them is refused instead of corrupting moves."""
    L = c6lib.product()
    rec = (c6lib.Record * 1)()
    assert L.c6_default_policy_play(80000, 1, 0, 1, 0, 0, rec, None, None, None) != 0
    assert b"2^24" in L.c6_last_error()
    assert L.c6_default_policy_play(0, 1, 0, 1, 0, 0, rec, None, None, None) != 0  # num_iter 0 panics in the reference
