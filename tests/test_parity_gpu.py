"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): winners, move sequences and visit counts bit-exact; Q/value
statistics within 1e-5 relative (they come out bit-exact too, which is what is asserted, with the
1e-5 bound as the documented tolerance).
"""
import ctypes as C
import os

import numpy as np
import pytest

import c6lib
from c6lib import EVAL_CALLBACK, EVAL_DET, EVAL_NEG, EVAL_PHILOX

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north_star tolerance for Q/value statistics


def make_prefix(n_stones, seed):
    """A legal opening of n_stones cells (row*15+col, in playing order) that nobody has won yet."""
    O = c6lib.oracle()
    rng = np.random.default_rng(seed)
    board = np.zeros(225, dtype=np.int8)
    cells = []
    while len(cells) < n_stones:
        c = int(rng.integers(225))
        if board[c] != 0:
            continue
        i = len(cells)
        board[c] = -1 if i == 0 else (1 if ((i - 1) // 2) % 2 == 0 else -1)
        if O.oracle_search(board.ctypes.data_as(C.POINTER(C.c_int8))) != 0:
            board[c] = 0
            continue
        cells.append(c)
    return cells


def run_engine(n_games, callback_name=None, first_id=0, prefixes=None, **cfg_kw):
    L = c6lib.product()
    cfg_kw.setdefault("slots", n_games)
    cfg_kw.setdefault("trace", 1)
    cfg_kw.setdefault("log2_records", 18)
    cfg_kw.setdefault("tie_break", 1)  # the product default: exact PUCT ties are broken uniformly (Philox), as the reference does
    cfg = c6lib.make_config(**cfg_kw)
    e = L.c6_engine_create(C.byref(cfg))
    assert e, L.c6_last_error()
    try:
        assert L.c6_engine_start(e, first_id, n_games) == 0, L.c6_last_error()
        for g, cells in enumerate(prefixes or []):  # stones placed from outside before the first search (slot g = game g)
            for c in cells:
                assert L.c6_engine_external_move(e, g, c // 15, c % 15) == 0, L.c6_last_error()
        if cfg.evaluator == EVAL_CALLBACK:
            cb = c6lib.CALLBACK(c6lib.oracle_callback(callback_name))
            assert L.c6_engine_run_callback(e, cb) == 0, L.c6_last_error()
        else:
            assert L.c6_engine_run_device(e) == 0, L.c6_last_error()
        rec = (c6lib.Record * n_games)()
        assert L.c6_engine_records(e, rec, n_games) == 0, L.c6_last_error()
        cv = np.zeros((n_games, 225, 225), dtype=np.int32)
        rn = np.zeros((n_games, 225), dtype=np.int32)
        rw = np.zeros((n_games, 225), dtype=np.float32)
        if cfg.trace:
            assert L.c6_engine_trace(e, cv.ctypes.data_as(C.POINTER(C.c_int32)), rn.ctypes.data_as(C.POINTER(C.c_int32)),
                                     rw.ctypes.data_as(C.POINTER(C.c_float))) == 0, L.c6_last_error()
        st = c6lib.Stats()
        assert L.c6_engine_stats(e, C.byref(st)) == 0
        return rec, cv, rn, rw, st
    finally:
        L.c6_engine_destroy(e)


def run_oracle(game_id, callback_name, evaluator, num_simulation, epsilon, seed, max_moves, alpha=0.03, c_puct=1.0, prefix=None, tie=1,
               virtual_loss=1):
    O = c6lib.oracle()
    p = c6lib.OracleParams(num_simulation, epsilon, alpha, c_puct, seed, 1 if epsilon != 0 else 0, c6lib.ORACLE_TIE[tie], 0, evaluator, max_moves,
                           virtual_loss)
    out = c6lib.OracleGameOut()
    cv = np.zeros((225, 225), dtype=np.int32)
    rn = np.zeros(225, dtype=np.int32)
    rw = np.zeros(225, dtype=np.float32)
    cb = c6lib.oracle_callback(callback_name) if callback_name else None
    pre = prefix or []
    O.oracle_self_play_from(cb, C.byref(p), game_id, (C.c_int * max(len(pre), 1))(*pre), len(pre), C.byref(out),
                            cv.ctypes.data_as(C.POINTER(C.c_int32)), rn.ctypes.data_as(C.POINTER(C.c_int32)),
                            rw.ctypes.data_as(C.POINTER(C.c_float)))
    return out, cv, rn, rw


def compare_games(n_games, rec, cv, rn, rw, first_id, callback_name, evaluator, num_simulation, epsilon, seed, max_moves, c_puct=1.0,
                  prefixes=None, tie=1, virtual_loss=1):
    for g in range(n_games):
        pre = prefixes[g] if prefixes else []
        out, ocv, orn, orw = run_oracle(first_id + g, callback_name, evaluator, num_simulation, epsilon, seed, max_moves, c_puct=c_puct,
                                        prefix=pre, tie=tie, virtual_loss=virtual_loss)
        r = rec[g]
        where = f"game {first_id + g}"
        assert r.game_id == first_id + g
        gm = [r.moves[i] for i in range(r.n_moves)]
        om = [out.moves[i] for i in range(out.n_moves)]
        if gm != om:
            first = next((i for i in range(min(len(gm), len(om))) if gm[i] != om[i]), min(len(gm), len(om)))
            # locate the divergence for the report
            raise AssertionError(f"{where}: move sequences diverge at stone {first}: engine {gm[first:first+3]} oracle {om[first:first+3]}; "
                                 f"visit rows equal before: {np.array_equal(cv[g][:first], ocv[:first])}")
        assert r.winner == out.winner, where
        assert r.finished == out.finished, where
        n, n0 = r.n_moves, len(pre)  # no search happens before the stones of a given opening
        # the engine records a trace row for every next() call, including a final call that returns None
        assert np.array_equal(cv[g][n0:n], ocv[n0:n]), f"{where}: root child visit counts differ"
        assert np.array_equal(rn[g][n0:n], orn[n0:n]), f"{where}: root visits differ"
        a, b = rw[g][n0:n], orw[n0:n]
        assert np.all(np.abs(a - b) <= REL_TOL * np.maximum(np.abs(b), 1e-30)), f"{where}: root q_sum beyond 1e-5 rel"
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"{where}: root q_sum not bit-identical"


# ---------------------------------------------------------------------------------------------
# win detection / playouts (BASELINE configs[2])
# ---------------------------------------------------------------------------------------------
def test_win_check_random_boards():
    """BASELINE configs[2] / SURVEY 8d (3a) at its full size: 16,777,216 random boards (cells iid over {Black, None, White}),
    every winner compared with the oracle's search() (threaded batch call on the bit-packed boards)."""
    L, O = c6lib.product(), c6lib.oracle()
    n = 1 << 24
    ms = C.c_float()
    boards = np.zeros((n, 16), dtype=np.uint32)
    winners = np.zeros(n, dtype=np.int8)
    assert L.c6_bench_win_check(n, 0xC6, 1, 0, C.byref(ms), boards.ctypes.data_as(C.POINTER(C.c_uint32)),
                                winners.ctypes.data_as(C.POINTER(C.c_int8)), n) == 0, L.c6_last_error()
    cells = c6lib.unpack_boards(boards[:1 << 16])
    # generator sanity: three symbols, roughly uniform
    frac = [(cells == v).mean() for v in (-1, 0, 1)]
    assert all(abs(f - 1 / 3) < 0.01 for f in frac), frac
    ow = np.zeros(n, dtype=np.int8)
    O.oracle_search_batch_packed(boards.ctypes.data_as(C.POINTER(C.c_uint32)), n, ow.ctypes.data_as(C.POINTER(C.c_int8)), 0)
    assert np.array_equal(winners, ow), f"{int((winners != ow).sum())} of {n} winners differ, first at {int(np.argmax(winners != ow))}"
    assert len(set(ow[:1 << 16].tolist())) == 3  # all three outcomes occur
    # the unpacked entry point of the oracle (the restatement proper) agrees with the packed one on a slice
    ow2 = np.zeros(1 << 16, dtype=np.int8)
    O.oracle_search_batch(cells.ctypes.data_as(C.POINTER(C.c_int8)), 1 << 16, ow2.ctypes.data_as(C.POINTER(C.c_int8)), 0)
    assert np.array_equal(ow2, ow[:1 << 16])


def test_win_check_reference_vectors():
    """search_winner/tests.rs:102-216 replayed through the kernel (incl. the both-colours scan-order case)."""
    L, O = c6lib.product(), c6lib.oracle()
    boards = []

    def snap(t):
        boards.append(np.array(t, dtype=np.int8).reshape(225).copy())

    t = np.zeros((15, 15), dtype=np.int8)
    for i in range(5):
        t[0][i] = -1
    snap(t); t[0][5] = -1; snap(t)
    for i in range(5):
        t[5 - i][0] = 1
    snap(t); t[0][0] = 1; snap(t)
    for i in range(5):
        t[5 - i][5 - i] = -1
    snap(t); t[0][0] = -1; snap(t)
    for i in range(5):
        t[5 - i][i] = 1
    snap(t); t[0][5] = 1; snap(t)
    expected = [0, -1, -1, 1, 1, -1, -1, 1]
    for base in (0, 14):
        for init in (0, 1, 9):
            for kind in ("row", "col", "rd", "ld"):
                for upto in (5, 6):
                    b = np.zeros((15, 15), dtype=np.int8)
                    for i in range(upto):
                        if kind == "row":
                            b[base][init + i] = 1
                        elif kind == "col":
                            b[init + i][base] = -1
                        elif kind == "rd":
                            b[init + i][init + i] = 1
                        else:
                            b[init + i][14 - init - i] = -1
                    snap(b)
                    expected.append(0 if upto == 5 else (1 if kind in ("row", "rd") else -1))
    # overline (7) and a full board without runs are also decided like the oracle
    b = np.zeros((15, 15), dtype=np.int8)
    b[7][3:10] = -1
    snap(b); expected.append(-1)
    cells = np.stack(boards)
    n = cells.shape[0]
    packed = c6lib.pack_boards(cells)
    w = np.zeros(n, dtype=np.int8)
    assert L.c6_win_check_host(packed.ctypes.data_as(C.POINTER(C.c_uint32)), w.ctypes.data_as(C.POINTER(C.c_int8)), n, 0) == 0
    ow = np.array([O.oracle_search(cells[i].ctypes.data_as(C.POINTER(C.c_int8))) for i in range(n)], dtype=np.int8)
    assert w.tolist() == expected
    assert ow.tolist() == expected


def test_win_check_empty_batch():
    L = c6lib.product()
    assert L.c6_win_check_host(None, None, 0, 0) == 0


def test_playouts_match_oracle():
    """1,048,576 random playouts compared end to end (winner, stone count, final board) with the oracle's
    rollout on the same Philox stream."""
    L, O = c6lib.product(), c6lib.oracle()
    n, first = 1 << 20, 123456789012
    ms = C.c_float()
    w = np.zeros(n, dtype=np.int8)
    s = np.zeros(n, dtype=np.uint8)
    b = np.zeros((n, 16), dtype=np.uint32)
    tot = C.c_uint64()
    assert L.c6_bench_playouts(first, n, 0xC6, 1, 0, C.byref(ms), w.ctypes.data_as(C.POINTER(C.c_int8)),
                               s.ctypes.data_as(C.POINTER(C.c_uint8)), b.ctypes.data_as(C.POINTER(C.c_uint32)), n, C.byref(tot)) == 0, L.c6_last_error()
    assert tot.value == int(s.astype(np.int64).sum())
    ow, os_, ob = np.zeros(n, dtype=np.int8), np.zeros(n, dtype=np.uint8), np.zeros((n, 16), dtype=np.uint32)
    O.oracle_playout_batch(0xC6, first, n, ow.ctypes.data_as(C.POINTER(C.c_int8)), os_.ctypes.data_as(C.POINTER(C.c_uint8)),
                           ob.ctypes.data_as(C.POINTER(C.c_uint32)), 0)
    assert np.array_equal(w, ow) and np.array_equal(s, os_)
    assert np.array_equal(b, ob)
    # the single-playout entry point (unpacked board) agrees on a few
    cells = c6lib.unpack_boards(b[:64])
    for i in range(0, 64, 8):
        w1, s1 = C.c_int(), C.c_int()
        b1 = np.zeros(225, dtype=np.int8)
        O.oracle_playout(0xC6, first + i, C.byref(w1), C.byref(s1), b1.ctypes.data_as(C.POINTER(C.c_int8)))
        assert (w[i], s[i]) == (w1.value, s1.value) and np.array_equal(cells[i], b1)
    # property: the winner is the colour of the last stone, stone counts follow the turn rule
    last = s.astype(np.int64) - 1
    turn = np.where(last == 0, -1, np.where(((last - 1) // 2) % 2 == 0, 1, -1))
    assert np.all((w == 0) | (w == turn))


# ---------------------------------------------------------------------------------------------
# self-play: callback path (C-ABI evaluator contract) and device evaluators
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cbname,sims,eps,n_games,max_moves,tie", [
    ("const_eval_callback", 4, 0.0, 1, 12, 0),     # SURVEY B.1 hand trace (derived for the lowest-index tie-break)
    ("const_eval_callback", 24, 0.0, 3, 12, 1),    # constant priors: every select step is a many-way exact tie
    ("det_eval_callback", 2, 0.0, 6, 0, 1),        # reference sample budget, full games
    ("neg_eval_callback", 2, 0.25, 6, 0, 1),       # light_weight() defaults incl. noise
    ("det_eval_callback", 48, 0.0, 4, 30, 1),
    ("neg_eval_callback", 48, 0.0, 4, 30, 0),
    ("neg_eval_callback", 32, 0.25, 4, 24, 1),
    ("neg_eval_callback", 32, 0.25, 4, 24, 0),
])
def test_selfplay_callback_path(cbname, sims, eps, n_games, max_moves, tie):
    seed = 0xC6
    rec, cv, rn, rw, st = run_engine(n_games, cbname, first_id=10, num_simulation=sims, epsilon=eps, seed=seed,
                                     evaluator=EVAL_CALLBACK, max_moves=max_moves, tie_break=tie)
    compare_games(n_games, rec, cv, rn, rw, 10, cbname, 0, sims, eps, seed, max_moves, tie=tie)
    if cbname == "const_eval_callback" and tie == 0:
        assert [rec[0].moves[i] for i in range(7)] == [2, 4, 5, 6, 7, 8, 9]  # (0,2),(0,4),(0,5),... as derived by hand


@pytest.mark.parametrize("evaluator,cbname,sims,eps,n_games,max_moves,tie", [
    (EVAL_PHILOX, None, 2, 0.0, 8, 0, 1),
    (EVAL_PHILOX, None, 2, 0.25, 8, 0, 1),
    (EVAL_PHILOX, None, 64, 0.0, 4, 24, 1),
    (EVAL_PHILOX, None, 64, 0.25, 4, 24, 1),
    (EVAL_PHILOX, None, 64, 0.25, 4, 24, 0),
    (EVAL_DET, "det_eval_callback", 40, 0.0, 4, 24, 1),
    (EVAL_DET, "det_eval_callback", 40, 0.0, 4, 24, 0),
    (EVAL_NEG, "neg_eval_callback", 40, 0.25, 4, 24, 1),
])
def test_selfplay_device_path(evaluator, cbname, sims, eps, n_games, max_moves, tie):
    seed = 0xC6
    rec, cv, rn, rw, st = run_engine(n_games, None, first_id=77, num_simulation=sims, epsilon=eps, seed=seed,
                                     evaluator=evaluator, max_moves=max_moves, tie_break=tie)
    compare_games(n_games, rec, cv, rn, rw, 77, cbname, 1 if evaluator == EVAL_PHILOX else 0, sims, eps, seed, max_moves, tie=tie)


@pytest.mark.parametrize("alpha", [0.3, 1.0, 2.5])
def test_selfplay_other_dirichlet_alphas(alpha):
    """Noise with another shape: alpha < 1 runs the same f32 log-space sampler, alpha >= 1 the f64 Marsaglia-Tsang one."""
    L = c6lib.product()
    seed, sims, eps, n, mm = 0xC6, 24, 0.25, 3, 10
    cfg = c6lib.make_config(num_simulation=sims, epsilon=eps, dirichlet_alpha=alpha, seed=seed, evaluator=EVAL_PHILOX, slots=n, max_moves=mm,
                            log2_records=16)
    rec = (c6lib.Record * n)()
    assert L.c6_self_play_device(C.byref(cfg), 900, n, rec, None) == 0, L.c6_last_error()
    for g in range(n):
        out, *_ = run_oracle(900 + g, None, 1, sims, eps, seed, mm, alpha=alpha)
        assert [rec[g].moves[i] for i in range(rec[g].n_moves)] == [out.moves[i] for i in range(out.n_moves)], (alpha, g)


@pytest.mark.parametrize("evaluator,cbname,sims,c_puct,n_stones", [
    # Nearly full boards searched with a low c_puct: every simulation runs many stones deep, so the rare rules of
    # expand_leaf decide the result (co-parents through an earlier path stone, expanded nodes orphaned by the
    # collector; oracle/pool_model.cpp counts how often they fire in positions like these), full-board leaves
    # and single-child nodes occur, and the games end with policy() returning None on a full board.
    (EVAL_NEG, "neg_eval_callback", 300, 0.2, 214),
    (EVAL_DET, "det_eval_callback", 120, 0.3, 205),
    (EVAL_NEG, "neg_eval_callback", 64, 1.0, 190),
    (EVAL_CALLBACK, "neg_eval_callback", 150, 0.2, 212),   # the same through the split descend / expand kernels
    (EVAL_NEG, "neg_eval_callback", 500, 0.0, 203),        # c_puct = 0: average path of 16-19 stones, far past the fused
                                                           # kernel's 12-deep shared-memory path cache
])
def test_selfplay_deep_trees_late_game(evaluator, cbname, sims, c_puct, n_stones):
    seed, n_games = 0xC6, 3
    prefixes = [make_prefix(n_stones + g, 1000 * n_stones + g) for g in range(n_games)]
    rec, cv, rn, rw, st = run_engine(n_games, cbname if evaluator == EVAL_CALLBACK else None, first_id=300, prefixes=prefixes,
                                     num_simulation=sims, epsilon=0.0, c_puct=c_puct, seed=seed, evaluator=evaluator)
    compare_games(n_games, rec, cv, rn, rw, 300, cbname, 0, sims, 0.0, seed, 0, c_puct=c_puct, prefixes=prefixes)
    assert st.sum_depth > 1.9 * st.sims  # the trees really are deep


@pytest.mark.parametrize("evaluator,cbname,sims,c_puct,n_stones", [
    # Dirichlet noise ON near the end of the game: nodes with 2..40 children, so the rank-indexed per-child loops run
    # one or two partly filled rounds, the slow-path draws of a step fit one partly filled pass, and single-child /
    # full-board shortcuts sit next to noisy steps.
    (EVAL_PHILOX, None, 200, 1.0, 186),
    (EVAL_PHILOX, None, 120, 0.3, 205),
    (EVAL_NEG, "neg_eval_callback", 90, 0.5, 216),
    (EVAL_CALLBACK, "det_eval_callback", 60, 1.0, 196),
])
def test_selfplay_late_game_with_noise(evaluator, cbname, sims, c_puct, n_stones):
    seed, n_games, eps = 0xC6, 3, 0.25
    prefixes = [make_prefix(n_stones + g, 7000 * n_stones + g) for g in range(n_games)]
    rec, cv, rn, rw, st = run_engine(n_games, cbname if evaluator == EVAL_CALLBACK else None, first_id=700, prefixes=prefixes,
                                     num_simulation=sims, epsilon=eps, c_puct=c_puct, seed=seed, evaluator=evaluator)
    compare_games(n_games, rec, cv, rn, rw, 700, cbname, 1 if evaluator == EVAL_PHILOX else 0, sims, eps, seed, 0, c_puct=c_puct,
                  prefixes=prefixes)


@pytest.mark.parametrize("evaluator,cbname,sims,eps,K,n_games,max_moves", [
    # opt-in multi-leaf mode (virtual loss): waves of up to K descents per game, leaves expanded afterwards; the oracle twin
    # (AlphaZero::search_wave) runs the same schedule on the reference's map-based tree
    (EVAL_PHILOX, None, 64, 0.25, 2, 4, 20),
    (EVAL_PHILOX, None, 64, 0.25, 4, 4, 20),
    (EVAL_PHILOX, None, 50, 0.0, 8, 3, 20),      # budget not a multiple of K
    (EVAL_PHILOX, None, 2, 0.25, 4, 6, 0),       # whole games, budget < K: the second descent of the first wave collides
    (EVAL_DET, "det_eval_callback", 40, 0.0, 4, 3, 20),
    (EVAL_NEG, "neg_eval_callback", 40, 0.25, 3, 3, 20),
    (EVAL_CALLBACK, "neg_eval_callback", 40, 0.25, 4, 5, 20),   # K leaves per game in every evaluator batch
    (EVAL_CALLBACK, "det_eval_callback", 33, 0.0, 8, 5, 16),
    (EVAL_CALLBACK, "neg_eval_callback", 2, 0.25, 2, 4, 0),
])
def test_selfplay_multi_leaf_mode(evaluator, cbname, sims, eps, K, n_games, max_moves):
    seed = 0xC6
    rec, cv, rn, rw, st = run_engine(n_games, cbname if evaluator == EVAL_CALLBACK else None, first_id=40, num_simulation=sims, epsilon=eps,
                                     seed=seed, evaluator=evaluator, max_moves=max_moves, virtual_loss=K)
    compare_games(n_games, rec, cv, rn, rw, 40, cbname, 1 if evaluator == EVAL_PHILOX else 0, sims, eps, seed, max_moves, virtual_loss=K)
    stones = sum(rec[g].n_moves for g in range(n_games))
    assert st.sims == sims * stones if max_moves else st.sims >= sims * (stones - n_games)  # every stone got its full budget


@pytest.mark.parametrize("evaluator,cbname,sims,c_puct,n_stones,K", [
    (EVAL_NEG, "neg_eval_callback", 200, 0.2, 212, 4),     # deep late-game trees: transposed children between waiting leaves
    (EVAL_CALLBACK, "neg_eval_callback", 120, 0.3, 208, 3),
    (EVAL_PHILOX, None, 150, 1.0, 196, 8),
])
def test_multi_leaf_mode_late_game(evaluator, cbname, sims, c_puct, n_stones, K):
    seed, n_games, eps = 0xC6, 3, 0.25
    prefixes = [make_prefix(n_stones + g, 3000 * n_stones + g) for g in range(n_games)]
    rec, cv, rn, rw, st = run_engine(n_games, cbname if evaluator == EVAL_CALLBACK else None, first_id=800, prefixes=prefixes,
                                     num_simulation=sims, epsilon=eps, c_puct=c_puct, seed=seed, evaluator=evaluator, virtual_loss=K)
    compare_games(n_games, rec, cv, rn, rw, 800, cbname, 1 if evaluator == EVAL_PHILOX else 0, sims, eps, seed, 0, c_puct=c_puct,
                  prefixes=prefixes, virtual_loss=K)


def test_selfplay_800_sims_reference_default():
    """Reference default budget (800 simulations, eps .25, alpha .03) for the first stones of two games."""
    seed, sims, eps, mm = 0xC6, 800, 0.25, 6
    rec, cv, rn, rw, st = run_engine(2, None, first_id=0, num_simulation=sims, epsilon=eps, seed=seed, evaluator=EVAL_PHILOX, max_moves=mm)
    compare_games(2, rec, cv, rn, rw, 0, None, 1, sims, eps, seed, mm)
    assert st.sims == 2 * sims * mm


def test_slot_refill_and_game_order():
    """More games than resident slots: finished slots are refilled, records come back by game id."""
    seed, sims = 3, 2
    rec, cv, rn, rw, st = run_engine(10, None, first_id=5, slots=3, num_simulation=sims, epsilon=0.0, seed=seed, evaluator=EVAL_PHILOX)
    compare_games(10, rec, cv, rn, rw, 5, None, 1, sims, 0.0, seed, 0)
    assert st.games_done == 10
    for g in range(10):
        r = rec[g]
        if r.winner != 0:  # path.last().turn == winner (alphazero_policy/tests.rs:198-202)
            last = r.n_moves - 1
            turn = -1 if last == 0 else (1 if ((last - 1) // 2) % 2 == 0 else -1)
            assert r.winner == turn


def test_engine_on_recycled_blocks_equals_engine_on_fresh_memory():
    """A destroyed engine leaves its device and pinned blocks to the next one of the same shape (c6_engine.cu: cached_alloc): the
    second engine starts on whatever the first left there — another seed's tree, another budget's counters — and must play
    exactly what an engine on fresh memory plays, on the device path and on the callback path."""
    L = c6lib.product()
    L.c6_release_cached_memory()
    kw = dict(slots=3, num_simulation=24, epsilon=0.25, evaluator=EVAL_PHILOX, max_moves=9)
    fresh = [bytes(run_engine(5, None, seed=s, **kw)[0]) for s in (11, 12)]
    L.c6_release_cached_memory()
    first = bytes(run_engine(5, None, seed=11, **kw)[0])  # fresh memory again
    second = bytes(run_engine(5, None, seed=12, **kw)[0])  # the blocks seed 11 has just left
    third = bytes(run_engine(5, None, seed=11, **kw)[0])
    assert [first, second, third] == [fresh[0], fresh[1], fresh[0]]
    kw = dict(slots=3, num_simulation=24, epsilon=0.25, evaluator=EVAL_CALLBACK, max_moves=9)
    a = bytes(run_engine(5, "det_eval_callback", seed=11, **kw)[0])
    b = bytes(run_engine(5, "neg_eval_callback", seed=12, **kw)[0])
    L.c6_release_cached_memory()
    assert bytes(run_engine(5, "neg_eval_callback", seed=12, **kw)[0]) == b and bytes(run_engine(5, "det_eval_callback", seed=11, **kw)[0]) == a
    L.c6_release_cached_memory()


def test_zero_simulations():
    rec, *_ = run_engine(2, None, num_simulation=0, epsilon=0.0, evaluator=EVAL_PHILOX)
    assert all(rec[g].n_moves == 0 and rec[g].winner == 0 for g in range(2))


def test_pool_too_small_for_the_budget_is_refused_loudly():
    """A ring that cannot hold even a fresh tree of one stone's search is refused at construction, naming the right knob."""
    L = c6lib.product()
    cfg = c6lib.make_config(num_simulation=64, epsilon=0.0, evaluator=EVAL_PHILOX, slots=1, log2_records=10)
    assert not L.c6_engine_create(C.byref(cfg))
    assert b"raise log2_records" in L.c6_last_error() and b"at least 14" in L.c6_last_error()
    cfg = c6lib.make_config(num_simulation=64, epsilon=0.0, evaluator=EVAL_PHILOX, slots=1, virtual_loss=9)
    assert not L.c6_engine_create(C.byref(cfg)) and b"virtual_loss" in L.c6_last_error()


@pytest.mark.parametrize("evaluator,cbname,log2", [(EVAL_PHILOX, None, 15), (EVAL_CALLBACK, "peaked_eval_callback", 15),
                                                   (EVAL_NEG, "neg_eval_callback", 15)])
def test_full_pool_restarts_the_search_instead_of_failing_the_batch(evaluator, cbname, log2):
    """a game whose node pool fills up (here: a ring of 2^15 records for 143 simulations, which holds a fresh tree —
    143 x 225 + 452 records at most — but not a fresh tree next to the subtree kept from the stone before) drops its tree and searches the stone again from a fresh root; the batch goes
    on, every game ends as a legal finished game, and the restarts are counted."""
    L = c6lib.product()
    n, sims = 6, 143
    cfg = c6lib.make_config(num_simulation=sims, epsilon=0.25, seed=21, evaluator=evaluator, slots=n, log2_records=log2)
    e = L.c6_engine_create(C.byref(cfg))
    assert e, L.c6_last_error()
    try:
        assert L.c6_engine_start(e, 0, n) == 0
        if evaluator == EVAL_CALLBACK:
            assert L.c6_engine_run_callback(e, c6lib.CALLBACK(c6lib.oracle_callback(cbname))) == 0, L.c6_last_error()
        else:
            assert L.c6_engine_run_device(e) == 0, L.c6_last_error()
        rec = (c6lib.Record * n)()
        assert L.c6_engine_records(e, rec, n) == 0
        st = c6lib.Stats()
        assert L.c6_engine_stats(e, C.byref(st)) == 0
    finally:
        L.c6_engine_destroy(e)
    assert st.games_done == n and st.degraded > 0 and st.peak_rec <= (1 << log2)
    for g in range(n):
        r = rec[g]
        mv = [r.moves[i] for i in range(r.n_moves)]
        assert r.finished == 1 and len(set(mv)) == len(mv) and (r.winner != 0 or r.n_moves == 225)


def test_peaked_evaluator_whole_games_at_800_simulations():
    """a concentrated evaluator (90 % of the prior on one cell) over whole games at the reference default budget in
    the default pool (2^18 records): the games finish, and whenever no restart was needed they equal the oracle's."""
    L = c6lib.product()
    n, sims, seed = 2, 800, 77
    cfg = c6lib.make_config(num_simulation=sims, epsilon=0.25, seed=seed, evaluator=EVAL_CALLBACK, slots=n, log2_records=18)
    e = L.c6_engine_create(C.byref(cfg))
    assert e, L.c6_last_error()
    try:
        assert L.c6_engine_start(e, 0, n) == 0
        assert L.c6_engine_run_callback(e, c6lib.CALLBACK(c6lib.oracle_callback("peaked_eval_callback"))) == 0, L.c6_last_error()
        rec = (c6lib.Record * n)()
        assert L.c6_engine_records(e, rec, n) == 0
        st = c6lib.Stats()
        assert L.c6_engine_stats(e, C.byref(st)) == 0
    finally:
        L.c6_engine_destroy(e)
    assert st.games_done == n
    if st.degraded == 0:
        for g in range(n):
            out, *_ = run_oracle(g, "peaked_eval_callback", 0, sims, 0.25, seed, 0)
            assert [rec[g].moves[i] for i in range(rec[g].n_moves)] == [out.moves[i] for i in range(out.n_moves)], g
    print("peak records per game", st.peak_rec, "restarts", st.degraded)


def test_legacy_cpp_self_play_matches_oracle(monkeypatch):
    """cpp_self_play (connect6.hpp:50-58) end to end: host evaluator buffers, Path[] records."""
    L = c6lib.product()
    os.environ["C6_SEED"] = "198"
    try:
        cb = c6lib.CALLBACK(c6lib.oracle_callback("neg_eval_callback"))
        v = L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 3, 0.25, 0.03, 1.0, False, 4)
    finally:
        del os.environ["C6_SEED"]
    assert v.len == 4
    for g in range(4):
        out, *_ = run_oracle(g, "neg_eval_callback", 0, 3, 0.25, 198, 0)
        r = v.vec[g]
        assert r.winner == out.winner and r.len == out.n_moves
        board = np.zeros((15, 15), dtype=np.int32)
        for m in range(r.len):
            p = r.paths[m]
            assert (p.row * 15 + p.col, p.turn) == (out.moves[m], out.turns[m])
            assert np.array_equal(np.ctypeslib.as_array(p.board), board)  # board BEFORE the stone
            board[p.row][p.col] = p.turn


# ---------------------------------------------------------------------------------------------
# committed golden traces (tests/golden/selfplay_traces.json, made by the pinned oracle): no oracle
# needed at run time
# ---------------------------------------------------------------------------------------------
import json  # noqa: E402

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_TRACES = json.load(open(os.path.join(_GOLDEN, "selfplay_traces.json")))


@pytest.mark.parametrize("case", _TRACES, ids=lambda c: c["name"])
def test_engine_reproduces_golden_traces(case):
    n = len(case["games"])
    first = case["first_id"]
    kw = dict(num_simulation=case["num_simulation"], epsilon=case["epsilon"], seed=case["seed"], max_moves=case["max_moves"],
              tie_break=case["tie_break"], virtual_loss=case.get("virtual_loss", 1))
    if case["evaluator"] == 1:
        rec, cv, rn, rw, st = run_engine(n, None, first_id=first, evaluator=EVAL_PHILOX, **kw)
    else:
        rec, cv, rn, rw, st = run_engine(n, case["callback"], first_id=first, evaluator=EVAL_CALLBACK, **kw)
    tot = dict(sims=0, sum_ck=0, sum_cleaf=0, sum_depth=0)
    for g, want in enumerate(case["games"]):
        r = rec[g]
        m = r.n_moves
        assert [r.moves[i] for i in range(m)] == want["moves"], want["game_id"]
        assert (r.winner, r.finished) == (want["winner"], want["finished"])
        assert rn[g][:m].tolist() == want["root_visit"]
        assert rw[g][:m].view(np.uint32).tolist() == want["root_qsum_bits"]
        assert [int((cv[g][i].astype(np.int64) * np.arange(1, 226)).sum()) for i in range(m)] == want["child_visit_checksum"]
        for k in tot:
            tot[k] += want[k]
    # the device counters behind the roofline's algorithmic-bytes formula equal the oracle's
    assert (st.sims, st.sum_ck, st.sum_cleaf, st.sum_depth) == (tot["sims"], tot["sum_ck"], tot["sum_cleaf"], tot["sum_depth"])


_HEADLINE = json.load(open(os.path.join(_GOLDEN, "selfplay_headline.json")))


@pytest.mark.parametrize("case", _HEADLINE, ids=lambda c: c["name"])
def test_headline_whole_games_in_full_engine(case):
    """Whole games at the headline configuration (800 simulations per stone, eps .25, RandomEvaluator) and 64 whole games at 50
    simulations, played INSIDE a 4,096-slot engine with slot refill: the golden game ids are scattered among filler games
    (some start in a refilled slot), and every stone, root visit count, root value sum (bits) and child-visit checksum must
    equal the committed oracle trace (tests/golden/make_headline.py)."""
    L = c6lib.product()
    total, slots = case["total_games"], case["slots"]
    cfg = c6lib.make_config(num_simulation=case["num_simulation"], epsilon=case["epsilon"], dirichlet_alpha=case["dirichlet_alpha"],
                            c_puct=case["c_puct"], seed=case["seed"], evaluator=EVAL_PHILOX, slots=slots, trace=1, log2_records=18,
                            tie_break=case["tie_break"])
    e = L.c6_engine_create(C.byref(cfg))
    assert e, L.c6_last_error()
    try:
        assert L.c6_engine_start(e, 0, total) == 0, L.c6_last_error()
        assert L.c6_engine_run_device(e) == 0, L.c6_last_error()
        rec = (c6lib.Record * total)()
        assert L.c6_engine_records(e, rec, total) == 0, L.c6_last_error()
        st = c6lib.Stats()
        assert L.c6_engine_stats(e, C.byref(st)) == 0
        assert st.games_done == total
        assert st.peak_rec <= (1 << 18)
        cv = np.zeros((225, 225), dtype=np.int32)
        rn = np.zeros(225, dtype=np.int32)
        rw = np.zeros(225, dtype=np.float32)
        for want in case["games"]:
            g = want["game_id"]
            r = rec[g]
            m = r.n_moves
            got = [r.moves[i] for i in range(m)]
            if got != want["moves"]:
                first = next((i for i in range(min(m, len(want["moves"]))) if got[i] != want["moves"][i]), min(m, len(want["moves"])))
                raise AssertionError(f"{case['name']} game {g}: stones diverge at {first}: engine {got[first:first + 3]} oracle {want['moves'][first:first + 3]}")
            assert (r.winner, r.finished, r.game_id) == (want["winner"], 1, g)
            assert L.c6_engine_trace_game(e, g, cv.ctypes.data_as(C.POINTER(C.c_int32)), rn.ctypes.data_as(C.POINTER(C.c_int32)),
                                          rw.ctypes.data_as(C.POINTER(C.c_float))) == 0, L.c6_last_error()
            assert rn[:m].tolist() == want["root_visit"], g
            assert rw[:m].view(np.uint32).tolist() == want["root_qsum_bits"], g
            assert [int((cv[i].astype(np.int64) * np.arange(1, 226)).sum()) for i in range(m)] == want["child_visit_checksum"], g
        # every filler game is a legal finished game too
        for g in range(total):
            r = rec[g]
            assert r.finished == 1 and r.game_id == g and len({r.moves[i] for i in range(r.n_moves)}) == r.n_moves
    finally:
        L.c6_engine_destroy(e)


def test_first_level_visits_spread_over_the_board():
    """with the uniform tie-break the first stone of 10,000 two-simulation games (= the child the second
    simulation picked among 225 exactly tied ones) is spread over the whole board; in lowest-index mode it is always cell 0."""
    L = c6lib.product()
    n = 10000
    rec = (c6lib.Record * n)()
    cfg = c6lib.make_config(num_simulation=2, epsilon=0.25, seed=9, evaluator=EVAL_PHILOX, slots=2048, max_moves=1, log2_records=10)
    assert L.c6_self_play_device(C.byref(cfg), 0, n, rec, None) == 0, L.c6_last_error()
    cnt = np.bincount([rec[g].moves[0] for g in range(n)], minlength=225)
    assert cnt.min() > 0 and cnt.max() < 4 * n / 225, (cnt.min(), cnt.max())
    assert abs(cnt[0] - n / 225) < 6 * np.sqrt(n / 225)
    cfg.tie_break = 0
    assert L.c6_self_play_device(C.byref(cfg), 0, 64, rec, None) == 0, L.c6_last_error()
    assert all(rec[g].moves[0] == 0 for g in range(64))


def test_sharded_equals_unsharded():
    """Records are a pure function of (seed, game id): two half-shards == one run."""
    from pyconnect6 import sharded
    kw = dict(num_simulation=8, epsilon=0.25, seed=11, max_moves=20)
    whole, _ = sharded.play_shard_device(0, 6, **kw)
    parts = []
    for r in range(2):
        lo, n = sharded.shard_range(6, r, 2)
        parts.append(sharded.play_shard_device(lo, n, **kw)[0])
    assert np.array_equal(np.concatenate(parts), whole)
    assert [r["game_id"] for r in sharded.records_from_bytes(whole)] == list(range(6))


def test_full_games_properties_at_scale():
    """Size-independent properties on many full games (2 simulations, noise on): every record is a legal
    game — distinct cells, stone count == path length, winner == colour of the last stone and a six
    exists exactly when a winner is reported (checked with the win-check kernel itself)."""
    L = c6lib.product()
    n = 2048
    cfg = c6lib.make_config(num_simulation=2, epsilon=0.25, seed=5, evaluator=EVAL_PHILOX, slots=512, log2_records=12)
    rec = (c6lib.Record * n)()
    st = c6lib.Stats()
    assert L.c6_self_play_device(C.byref(cfg), 1000, n, rec, C.byref(st)) == 0, L.c6_last_error()
    assert st.games_done == n
    cells = np.zeros((n, 225), dtype=np.int8)
    prev = np.zeros((n, 225), dtype=np.int8)
    for g in range(n):
        r = rec[g]
        assert r.finished == 1 and r.game_id == 1000 + g
        mv = [r.moves[i] for i in range(r.n_moves)]
        assert len(set(mv)) == len(mv) and all(0 <= m < 225 for m in mv)
        for i, m in enumerate(mv):
            turn = -1 if i == 0 else (1 if ((i - 1) // 2) % 2 == 0 else -1)
            cells[g, m] = turn
            if i < len(mv) - 1:
                prev[g, m] = turn
        if r.winner != 0:
            last = r.n_moves - 1
            assert r.winner == (-1 if last == 0 else (1 if ((last - 1) // 2) % 2 == 0 else -1))
        else:
            assert r.n_moves == 225
    w = np.zeros(n, dtype=np.int8)
    wp = np.zeros(n, dtype=np.int8)
    pk, pkp = c6lib.pack_boards(cells), c6lib.pack_boards(prev)
    assert L.c6_win_check_host(pk.ctypes.data_as(C.POINTER(C.c_uint32)), w.ctypes.data_as(C.POINTER(C.c_int8)), n, 0) == 0
    assert L.c6_win_check_host(pkp.ctypes.data_as(C.POINTER(C.c_uint32)), wp.ctypes.data_as(C.POINTER(C.c_int8)), n, 0) == 0
    assert np.array_equal(w, np.array([rec[g].winner for g in range(n)], dtype=np.int8))
    assert not wp.any()  # no six before the last stone: the game stopped at the first win


def test_pyconnect6_self_play():
    """Connect6/pybind/test_pybind/test_pybind.py:16-24 and sample/python/random_policy.py."""
    import pyconnect6

    class RandomPolicy:
        def __call__(self, turn, board):
            size = len(board)
            assert turn in (-1, 1) and size % 8 == 0 and len(board[0]) == 225
            return np.random.rand(size), np.random.rand(size, 225)

    param = pyconnect6.default_param()
    param["num_simulation"] = 2
    param["num_game_thread"] = 2
    result = pyconnect6.self_play(RandomPolicy(), param)
    assert len(result) == 2
    for winner, path in result:
        assert winner in (-1, 0, 1) and len(path) >= 11
        if winner != 0:
            assert path[-1][0] == winner
    param["num_game_thread"] = 1
    winner, path = pyconnect6.self_play(RandomPolicy(), param)
    turn, board, (row, col) = path[0]
    assert turn == -1 and board == [0] * 225 and 0 <= row < 15 and 0 <= col < 15


def test_pyconnect6_evaluator_error_is_raised():
    import pyconnect6

    def bad(turn, board):
        raise KeyError("boom")

    param = pyconnect6.default_param()
    param["num_simulation"] = 2
    with pytest.raises(pyconnect6.pyconnect6.EvaluatorError):
        pyconnect6.self_play(bad, param)


def test_reference_catch2_suite_full():
    """The reference's Catch2 acceptance suite (test_cppbind, incl. Connect6::self_play 2 sims x 2 games and
    Connect6::play), prebuilt by __graft_entry__.build() from the reference's sources in place, linked
    against this library."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "test_cppbind")
    assert os.path.exists(exe), "tests/_build/test_cppbind is missing: __graft_entry__.build() makes it where /root/reference is mounted, and it travels with the snapshot"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "All tests passed" in r.stdout


def test_noise_sampler_bitwise_vs_oracle():
    """Every log2-Gamma draw of the engine's Dirichlet-noise sampler (alpha = 0.03 default, 0.003, 0.7, and the f64 alpha >= 1
    path), its f32 log2 / exp2, its f64 log / exp and its tie-break picks equal the oracle's independent restatement bit for bit."""
    L, O = c6lib.product(), c6lib.oracle()
    n = 1 << 17
    FP = C.POINTER(C.c_float)
    for shape in (0.03, 0.003, 0.7, 1.0, 2.5):
        out = np.zeros(n, dtype=np.float32)
        assert L.c6_probe_gamma(shape, 0xABCDEF01, 77, out.ctypes.data_as(FP), n, 0) == 0, L.c6_last_error()
        want = np.zeros(n, dtype=np.float32)
        O.oracle_gamma_log2_many(7, 9, 0, n, 0xABCDEF01, 77, shape, want.ctypes.data_as(FP))
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), shape
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-130, 127, 50000), rng.uniform(-1, 1, 50000), np.exp(rng.uniform(-87, 88, 50000)),
                         1.0 - np.arange(1, 3000) * 2.0 ** -24, 1.0 + np.arange(0, 3000) * 2.0 ** -23,
                         [1.1754944e-38, 1.0, 2.0, 0.5, 0.70710677, 0.70710683, 1.4142135, 1.4142137, -125.0, -125.00001, -126.0, 3.4e38]]).astype(np.float32)
    lg, ex, wl, we = (np.zeros_like(xs) for _ in range(4))
    assert L.c6_probe_math32(xs.ctypes.data_as(FP), lg.ctypes.data_as(FP), ex.ctypes.data_as(FP), len(xs), 0) == 0
    O.oracle_math32_many(xs.ctypes.data_as(FP), len(xs), wl.ctypes.data_as(FP), we.ctypes.data_as(FP))
    ok = xs < 127.0  # ex2_det is defined for t < 127
    assert np.array_equal(lg.view(np.uint32), wl.view(np.uint32))
    assert np.array_equal(ex[ok].view(np.uint32), we[ok].view(np.uint32))
    pos = xs > 0
    ref = np.log2(xs[pos].astype(np.float64))
    assert np.all(np.abs(lg[pos] - ref) <= 2.0 ** -23 * np.maximum(1.0, np.abs(ref)))
    xs = np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-1, 1, 20000), np.exp(rng.uniform(-700, 700, 20000)),
                         [5e-324, 1e-310, 1.0, 1.0 - 2 ** -53, 1.0 + 2 ** -52, 2.0, 0.5, -800.0, 800.0, 1e308]])
    logs, exps = np.zeros_like(xs), np.zeros_like(xs)
    assert L.c6_probe_math(xs.ctypes.data_as(C.POINTER(C.c_double)), logs.ctypes.data_as(C.POINTER(C.c_double)),
                           exps.ctypes.data_as(C.POINTER(C.c_double)), len(xs), 0) == 0
    wl = np.array([O.oracle_det_log(x) if x > 0 else 0.0 for x in xs])
    we = np.array([O.oracle_det_exp(x) for x in xs])
    assert np.array_equal(logs.view(np.uint64), wl.view(np.uint64))
    assert np.array_equal(exps.view(np.uint64), we.view(np.uint64))
    for count in (2, 3, 224, 225):
        got = np.zeros(4096, dtype=np.uint32)
        assert L.c6_probe_tie(0xC6, 5, count, got.ctypes.data_as(C.POINTER(C.c_uint32)), len(got), 0) == 0
        assert got.tolist() == [O.oracle_tie_pick(i, i % 7, 0xC6, 5, count) for i in range(len(got))]
        assert got.max() < count and len(set(got.tolist())) == min(count, 225)


@pytest.mark.parametrize("pipeline", ["1", "3", "7"])
def test_callback_pipeline_groups_do_not_change_results(pipeline, monkeypatch):
    """The callback path splits the resident games into pipeline groups (own stream, own batch region);
    records must not depend on the split, and slot refill must work across groups."""
    monkeypatch.setenv("C6_PIPELINE", pipeline)
    seed, sims, eps, n = 0xC6, 6, 0.25, 12
    rec, cv, rn, rw, st = run_engine(n, "neg_eval_callback", first_id=300, slots=7, num_simulation=sims, epsilon=eps, seed=seed,
                                     evaluator=EVAL_CALLBACK, max_moves=14)
    compare_games(n, rec, cv, rn, rw, 300, "neg_eval_callback", 0, sims, eps, seed, 14)
    assert st.games_done == n


def test_cpp_play_with_matches_oracle():
    """cpp_play_with (connect6.hpp:60-65): AlphaZero as Black on the GPU against White moves typed on stdin.
    The tree must be reused across White's stones exactly as the reference's `init` reuses it; checked against
    the oracle's MultiPolicy(AlphaZero, scripted White) with the same evaluator, seed and tie-break."""
    import subprocess
    import sys
    sims, eps, seed = 24, 0.25, 4242
    O = c6lib.oracle()
    p = c6lib.OracleParams(sims, eps, 0.03, 1.0, seed, 1, c6lib.ORACLE_TIE[1], 0, 0, 0)
    for attempt in range(50):  # a White script that does not collide with the stones Black turns out to play
        rng = np.random.default_rng(attempt)
        white = [(int(c) // 15, int(c) % 15) for c in rng.choice(225, size=10, replace=False)]
        out = c6lib.OracleGameOut()
        wm = (C.c_int * (2 * len(white)))(*[x for m in white for x in m])
        O.oracle_play_with(c6lib.oracle_callback("neg_eval_callback"), C.byref(p), 0, wm, len(white), C.byref(out))
        if out.finished == 1 and out.n_moves >= 12:
            break
    assert out.finished == 1
    want = [out.moves[i] for i in range(out.n_moves)]
    code = (
        "import ctypes as C, c6lib\n"
        "L = c6lib.product()\n"
        "L.cpp_play_with.restype = c6lib.PlayResult\n"
        "L.cpp_play_with.argtypes = [c6lib.CALLBACK, c6lib.ALLOC_PATH, C.c_int, C.c_float, C.c_double, C.c_float]\n"
        "cb = c6lib.CALLBACK(c6lib.oracle_callback('neg_eval_callback'))\n"
        f"r = L.cpp_play_with(cb, c6lib.alloc_path, {sims}, {eps}, 0.03, 1.0)\n"
        "import sys; sys.stdout.flush()\n"
        "print('RESULT', r.winner, ' '.join(str(r.paths[i].row * 15 + r.paths[i].col) for i in range(r.len)))\n"
    )
    stdin = "zz\n" + "".join(f"{chr(0x61 + r)}{chr(0x41 + c)}\n" for r, c in white)  # first line is rejected: "invalid input, retry"
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.abspath(__file__)), C6_SEED=str(seed))
    r = subprocess.run([sys.executable, "-c", code], input=stdin, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][0].split()
    assert int(line[1]) == out.winner
    assert [int(x) for x in line[2:]] == want
    assert "invalid input, retry" in r.stdout and "0 A B C D E F G H I J K L M N O" in r.stdout  # IoPolicy + Agent::debug transcript


def test_host_alphazero_policy_single_game_equals_batched_engine():
    """cpp_self_play with one game (the reference's n == 1 path) and with several give the same game 0."""
    L = c6lib.product()
    os.environ["C6_SEED"] = "77"
    try:
        cb = c6lib.CALLBACK(c6lib.oracle_callback("det_eval_callback"))
        one = L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 5, 0.25, 0.03, 1.0, False, 1)
        many = L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 5, 0.25, 0.03, 1.0, False, 3)
    finally:
        del os.environ["C6_SEED"]
    a, b = one.vec[0], many.vec[0]
    assert (a.winner, a.len) == (b.winner, b.len)
    assert all((a.paths[i].row, a.paths[i].col) == (b.paths[i].row, b.paths[i].col) for i in range(a.len))


def test_reference_cpp_sample_runs():
    """Connect6/sample/cpp/main.cpp unchanged (BASELINE configs[0] in its C++ form: one game, random evaluator,
    2 simulations, debug transcript), prebuilt by build() against the reference's header."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "sample_cpp")
    assert os.path.exists(exe), "tests/_build/sample_cpp is missing: __graft_entry__.build() makes it where /root/reference is mounted, and it travels with the snapshot"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Winner : " in r.stdout and "0 A B C D E F G H I J K L M N O" in r.stdout
    assert any(l.startswith(("Black (", "White (")) and "remain" in l for l in r.stdout.splitlines())


def test_in_process_sharding_equals_single_engine(monkeypatch):
    """c6_self_play_device_sharded / C6_DEVICES: one host thread + engine per listed device, games split by id.
    Listing device 0 twice exercises the threaded path on a one-GPU box; records must equal the unsharded run."""
    L = c6lib.product()
    n = 7
    cfg = c6lib.make_config(num_simulation=6, epsilon=0.25, seed=31, evaluator=EVAL_PHILOX, slots=4, max_moves=18, log2_records=14)
    whole = (c6lib.Record * n)()
    assert L.c6_self_play_device(C.byref(cfg), 40, n, whole, None) == 0, L.c6_last_error()
    parts = (c6lib.Record * n)()
    st = c6lib.Stats()
    devs = (C.c_int * 3)(0, 0, 0)
    assert L.c6_self_play_device_sharded(C.byref(cfg), devs, 3, 40, n, parts, C.byref(st)) == 0, L.c6_last_error()
    assert bytes(whole) == bytes(parts)
    assert st.games_done == n
    # the reference entry point with C6_DEVICES
    monkeypatch.setenv("C6_SEED", "5")
    cb = c6lib.CALLBACK(c6lib.oracle_callback("det_eval_callback"))
    one = L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 3, 0.25, 0.03, 1.0, False, 5)
    monkeypatch.setenv("C6_DEVICES", "0,0")
    two = L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 3, 0.25, 0.03, 1.0, False, 5)
    for g in range(5):
        a, b = one.vec[g], two.vec[g]
        assert (a.winner, a.len) == (b.winner, b.len)
        assert all((a.paths[i].row, a.paths[i].col) == (b.paths[i].row, b.paths[i].col) for i in range(a.len))


def test_device_pointer_evaluator_matches_host_callback_path():
    """C6_EVAL_DEVICE_CALLBACK + the torch bridge: the leaf batch is consumed in HBM through device pointers on the
    engine's stream.  The evaluator here round-trips the batch through the oracle's deterministic host evaluator,
    so the games must equal the oracle's (plumbing, ordering and layout are what is under test)."""
    import torch
    from pyconnect6 import torch_bridge
    neg = c6lib.CALLBACK(c6lib.oracle_callback("neg_eval_callback"))

    def evaluate(player, boards):
        n = boards.shape[0]
        pol = boards.reshape(n, 225).cpu().numpy().copy()
        val = np.zeros(n, dtype=np.float32)
        neg(player, val.ctypes.data_as(C.POINTER(C.c_float)), pol.ctypes.data_as(C.POINTER(C.c_float)), n)
        return torch.from_numpy(val).cuda(), torch.from_numpy(pol).cuda()

    n, sims, eps, seed, mm = 5, 12, 0.25, 9, 16
    rec, st = torch_bridge.self_play_torch(evaluate, n, num_simulation=sims, epsilon=eps, seed=seed, max_moves=mm, first_id=60, slots=3,
                                           log2_records=14)
    from pyconnect6 import sharded
    got = sharded.records_from_bytes(rec)
    for g in range(n):
        out, *_ = run_oracle(60 + g, "neg_eval_callback", 0, sims, eps, seed, mm)
        assert got[g]["moves"] == [out.moves[i] for i in range(out.n_moves)], g
        assert got[g]["winner"] == out.winner
    assert st.games_done == n


def test_torch_policy_value_net_drives_self_play():
    """A random-init net of the reference trainer's shape (cpp_weighted/model.hpp) as the evaluator, zero-copy."""
    import torch
    from pyconnect6 import torch_bridge, sharded
    torch.manual_seed(0)
    net = torch_bridge.WeightedPolicy(num_block=3, channels=32).cuda().eval()
    rec, st = torch_bridge.self_play_torch(torch_bridge.module_evaluator(net), 16, num_simulation=8, seed=1, max_moves=12, log2_records=13)
    games = sharded.records_from_bytes(rec)
    assert all(g["n_moves"] == 12 and len(set(g["moves"])) == 12 for g in games)
    assert st.sims == 16 * 8 * 12 and st.evals > 0


# ---------------------------------------------------------------------------------------------
# the reference's own callers, unchanged: byte-compiled / compiled in place by __graft_entry__.build()
# ---------------------------------------------------------------------------------------------
_BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")
_PYPKG = os.path.join(c6lib.PKG, "python")


def test_reference_pybind_tests_run_unchanged():
    """PB/test_pybind/test_pybind.py:16-49 (test_result_length: self_play 2 simulations x 2 games; test_echo_pyeval),
    byte-compiled from the reference tree, imported and run against our pyconnect6."""
    import subprocess
    import sys
    pyc = os.path.join(_BUILD, "ref_test_pybind.pycode")
    assert os.path.exists(pyc), "tests/_build/ref_test_pybind.pycode is missing: __graft_entry__.build() makes it where /root/reference is mounted"
    code = ("import importlib.machinery, importlib.util, sys\n"
            f"loader = importlib.machinery.SourcelessFileLoader('ref_test_pybind', {pyc!r})\n"
            "spec = importlib.util.spec_from_loader('ref_test_pybind', loader)\n"
            "m = importlib.util.module_from_spec(spec); loader.exec_module(m)\n"
            "tests = [n for n in dir(m) if n.startswith('test_')]\n"
            "assert sorted(tests) == ['test_echo_pyeval', 'test_result_length'], tests\n"
            "for n in tests: getattr(m, n)()\n"
            "print('REFERENCE PYTESTS PASSED', len(tests))\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=_PYPKG), timeout=600)
    assert r.returncode == 0 and "REFERENCE PYTESTS PASSED 2" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_reference_python_sample_runs_unchanged():
    """Connect6/sample/python/random_policy.py:28-37 (BASELINE configs[0] in its Python form: one game, random evaluator, 2
    simulations, debug transcript), byte-compiled from the reference tree and executed as a script."""
    import subprocess
    import sys
    pyc = os.path.join(_BUILD, "ref_random_policy.pycode")
    assert os.path.exists(pyc), "tests/_build/ref_random_policy.pycode is missing: __graft_entry__.build() makes it where /root/reference is mounted"
    code = ("import importlib.machinery, importlib.util\n"
            f"loader = importlib.machinery.SourcelessFileLoader('__main__', {pyc!r})\n"
            "spec = importlib.util.spec_from_loader('__main__', loader)\n"
            "m = importlib.util.module_from_spec(spec); loader.exec_module(m)\n")  # runs the script body, as `python random_policy.py` does
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=_PYPKG), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    last = r.stdout.strip().splitlines()[-1]
    assert last.startswith("winner ") and ", len " in last, last
    assert "0 A B C D E F G H I J K L M N O" in r.stdout and "remain" in r.stdout  # debug=True prints the game


def test_device_pointer_overload_of_the_reference_cpp_api():
    """include/connect6_device.hpp on top of the reference's own connect6.hpp: Connect6::self_play(DeviceCallback, ...) with a
    CUDA-kernel evaluator (constant value .5 / prior .25); every stone must equal the oracle's game with const_eval_callback."""
    import subprocess
    exe = os.path.join(_BUILD, "device_overload")
    assert os.path.exists(exe), "tests/_build/device_overload is missing: __graft_entry__.build() makes it where /root/reference is mounted"
    games, sims, seed = 5, 12, 0x51
    r = subprocess.run([exe, str(games), str(sims), str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l.split() for l in r.stdout.splitlines() if l.startswith("GAME")]
    assert len(lines) == games
    for g, l in enumerate(lines):
        out, *_ = run_oracle(g, "const_eval_callback", 0, sims, 0.25, seed, 0)
        assert int(l[1]) == out.winner and [int(x) for x in l[2:]] == [out.moves[i] for i in range(out.n_moves)], g
    tail = r.stdout.strip().splitlines()[-1].split()
    assert tail[0] == "CALLS" and int(tail[1]) > 0 and int(tail[5]) == 0  # boards on entry, values pre-zeroed


def test_host_mirror_random_evaluator_and_async_agent():
    """RandomEvaluator, AsyncAgent::run (generic thread-per-game form and the batched GPU form) and AlphaZero of the C++ mirror
    of the reference's Rust API (async_agent/tests.rs:8-47, alphazero_policy/tests.rs:186-204)."""
    import subprocess
    exe = os.path.join(_BUILD, "host_mirror")
    assert os.path.exists(exe), "tests/_build/host_mirror is missing: __graft_entry__.build() makes it"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "host mirror ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
