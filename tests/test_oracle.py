"""CPU suite, part 1: the oracle against the reference's known-answer vectors (tests/golden/, transcribed
from the reference's own tests with file:line) and against its own committed self-play traces.

The reference is Rust and cannot run here; these vectors are what pins the oracle.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

import c6lib

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def i8(a):
    return np.ascontiguousarray(a, dtype=np.int8).ctypes.data_as(C.POINTER(C.c_int8))


def test_reference_test_ports_pass():
    """oracle/oracle_tests.cpp: C++ ports of every deterministic reference test on the path."""
    c6lib.build_oracle()
    r = subprocess.run([c6lib.ORACLE_TESTS], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout


def test_win_detection_known_answers(oracle):
    vec = load("reference_known_answers.json")["win_detection"]
    assert len(vec) == 8 + 24 + 12
    for v in vec:
        assert oracle.oracle_search(i8(v["board"])) == v["winner"], v["cite"]


def test_symmetry_known_answers(oracle):
    s = load("reference_known_answers.json")["symmetries"]
    for kind, name in enumerate(["rotate_left", "rotate_right", "flip_vertical", "flip_horizontal"]):
        out = np.zeros(225, dtype=np.int8)
        oracle.oracle_transform(kind, i8(s["board"]), i8(out) if False else out.ctypes.data_as(C.POINTER(C.c_int8)))
        for r, c, want in s[name]:
            assert out[r * 15 + c] == want, (name, r, c)


def test_augment_recover_round_trip(oracle):
    """augment_tests.rs:115-142 on seeded random boards: 8 distinct planes, recover(augment(x)) == x."""
    rng = np.random.default_rng(6)
    for _ in range(20):
        b = rng.integers(-1, 2, size=225).astype(np.int8)
        planes = np.zeros((8, 225), dtype=np.int8)
        oracle.oracle_augment_way8(i8(b), planes.ctypes.data_as(C.POINTER(C.c_int8)))
        assert len({p.tobytes() for p in planes}) == 8
        f = planes.astype(np.float32)
        rec = np.zeros(225, dtype=np.float32)
        oracle.oracle_recover_way8(f.ctypes.data_as(C.POINTER(C.c_float)), rec.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(rec, b.astype(np.float32))


def test_rules_known_answers(oracle):
    steps = load("reference_known_answers.json")["rules"]["steps"]
    n = len(steps)
    pos = (C.c_int * (2 * n))(*[x for s in steps for x in s["pos"]])
    code, player, remain, ta, ra = [(C.c_int * n)() for _ in range(5)]
    oracle.oracle_game_replay(pos, n, code, player, remain, ta, ra)
    for i, s in enumerate(steps):
        if s["ok"]:
            assert code[i] == 0 and player[i] == s["player"] and remain[i] == s["num_remain"]
            assert (ta[i], ra[i]) == (s["turn_after"], s["remain_after"])
        else:
            assert code[i] == {"invalid position": 1, "already set position": 2}[s["err"]]


def test_scripted_game_known_answer(oracle):
    g = load("reference_known_answers.json")["scripted_game"]
    n = len(g["moves"])
    moves = (C.c_int * (2 * n))(*[x for m in g["moves"] for x in m])
    winner = C.c_int()
    paths = (C.c_int * (228 * 225))()
    err = C.create_string_buffer(64)
    ln = oracle.oracle_script_game(moves, n, C.byref(winner), paths, err)
    assert ln == g["path_len"] and winner.value == g["winner"]
    board = [0] * 225
    for k in range(ln):
        p = paths[k * 228:(k + 1) * 228]
        assert p[0] == g["turns"][k]
        assert list(p[1:226]) == board  # board BEFORE the stone (agent_impl/mod.rs:139-143)
        assert [p[226], p[227]] == list(g["moves"][k])
        board[p[226] * 15 + p[227]] = p[0]
    # error propagation (agent_impl/tests.rs:37-63): replaying a stone is Err("already set position")
    bad = (C.c_int * 4)(0, 0, 0, 0)
    assert oracle.oracle_script_game(bad, 2, C.byref(winner), paths, err) == -1
    assert err.value == b"already set position"


def run_oracle_case(case, g, faithful=0):
    O = c6lib.oracle()
    p = c6lib.OracleParams(case["num_simulation"], case["epsilon"], case["dirichlet_alpha"], case["c_puct"], case["seed"],
                           1 if case["epsilon"] != 0 else 0, c6lib.ORACLE_TIE[case.get("tie_break", 0)], faithful, case["evaluator"],
                           case["max_moves"], case.get("virtual_loss", 1))
    out = c6lib.OracleGameOut()
    cv = np.zeros((225, 225), dtype=np.int32)
    rn = np.zeros(225, dtype=np.int32)
    rw = np.zeros(225, dtype=np.float32)
    cb = c6lib.oracle_callback(case["callback"]) if case["callback"] else None
    O.oracle_self_play(cb, C.byref(p), g["game_id"], C.byref(out), cv.ctypes.data_as(C.POINTER(C.c_int32)),
                       rn.ctypes.data_as(C.POINTER(C.c_int32)), rw.ctypes.data_as(C.POINTER(C.c_float)))
    return out, cv, rn, rw


def check_against_trace(g, winner, finished, moves, cv, rn, rw):
    assert moves == g["moves"]
    assert (winner, finished) == (g["winner"], g["finished"])
    m = len(moves)
    assert rn[:m].tolist() == g["root_visit"]
    assert rw[:m].view(np.uint32).tolist() == g["root_qsum_bits"]
    assert [int((cv[i].astype(np.int64) * np.arange(1, 226)).sum()) for i in range(m)] == g["child_visit_checksum"]


@pytest.mark.parametrize("case", load("selfplay_traces.json"), ids=lambda c: c["name"])
def test_oracle_reproduces_committed_traces(case):
    for g in case["games"]:
        out, cv, rn, rw = run_oracle_case(case, g)
        check_against_trace(g, out.winner, out.finished, [out.moves[i] for i in range(out.n_moves)], cv, rn, rw)
        assert (out.sims, out.sum_ck, out.sum_cleaf, out.sum_depth) == (g["sims"], g["sum_ck"], g["sum_cleaf"], g["sum_depth"])


def test_lean_cost_model_equals_reference_cost_model():
    """The oracle's fast mode (Zobrist keys, superset GC) must give the results of the reference's own
    data structures (SipHash over the whole board, nothing freed)."""
    case = [c for c in load("selfplay_traces.json") if c["name"] == "philox_64sims_noise"][0]
    case = dict(case, max_moves=8)
    g = case["games"][0]
    a = run_oracle_case(case, g, faithful=0)
    b = run_oracle_case(case, g, faithful=1)
    assert [a[0].moves[i] for i in range(a[0].n_moves)] == [b[0].moves[i] for i in range(b[0].n_moves)]
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert np.array_equal(a[3].view(np.uint32), b[3].view(np.uint32))


def test_pool_model():
    """oracle/pool_model.cpp: the device node pool's way of sharing transposed children (sibling link runs,
    key table of expanded nodes, orphan flags — DESIGN.md §3) played side by side with the map-based
    restatement: every move, root statistic and child visit count identical, rare rules exercised."""
    import subprocess
    exe = os.path.join(c6lib.ORACLE_DIR, "_build", "pool_model")
    if not os.path.exists(exe):
        c6lib.build_oracle()
    r = subprocess.run([exe, "2", "quick"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "all cases identical to the oracle" in r.stdout
    late = [l for l in r.stdout.splitlines() if l.startswith("neg_late")][0].split()
    assert int(late[late.index("S1") + 4]) > 0 and int(late[late.index("S2") + 4]) > 0  # S1 hits, S2 hits


def test_b1_hand_trace(oracle):
    """derived by hand from the reference source, independent of any code here."""
    b1 = load("b1_hand_trace.json")
    case = {"num_simulation": 4, "epsilon": 0.0, "dirichlet_alpha": 0.03, "c_puct": 1.0, "seed": 1, "evaluator": 0,
            "callback": "const_eval_callback", "max_moves": 7}
    out, cv, rn, rw = run_oracle_case(case, {"game_id": 0})
    assert [[out.moves[i] // 15, out.moves[i] % 15] for i in range(7)] == b1["moves_rc"]
    assert (rn[0], rw[0]) == (b1["move1_root"]["visit"], b1["move1_root"]["q_sum"])
    assert (rn[1], rw[1]) == (b1["move2_root"]["visit"], b1["move2_root"]["q_sum"])


def test_philox_known_answer(oracle):
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors): zero, all-ones and the pi digits."""
    def run(ctr, key):
        out = (C.c_uint32 * 4)()
        oracle.oracle_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        return list(out)
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert run([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert run([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_deterministic_log_exp(oracle):
    """The noise sampler's own log/exp (no libm, so CPU and GPU agree bit for bit) are accurate."""
    import math
    for x in [1e-300, 1e-10, 0.03, 0.5, 1.0, 2.0, 7.25, 1e10]:
        assert abs(oracle.oracle_det_log(x) - math.log(x)) <= 1e-13 * max(1.0, abs(math.log(x)))
    for x in [-700.0, -30.0, -1.0, -1e-9, 0.0, 0.3, 1.0, 20.0]:
        assert abs(oracle.oracle_det_exp(x) - math.exp(x)) <= 1e-13 * math.exp(x)


def test_gamma_sampler_distribution(oracle):
    """The search's sampler (log2 of a Gamma(alpha) draw, f32): the distribution is what is contractual (rand 0.6's own
    stream is unseedable).  Kolmogorov-Smirnov against the Gamma CDF in log space, mean = variance = alpha, and
    E[ln X] = digamma(alpha), for the default alpha = 0.03, two other shapes below 1 and the alpha >= 1 path."""
    from scipy import special, stats
    n = 400000
    FP = C.POINTER(C.c_float)
    for shape in (0.03, 0.3, 0.7, 1.0, 2.5):
        lg = np.zeros(n, dtype=np.float32)
        oracle.oracle_gamma_log2_many(3, 1, 0, n, 123, 7, shape, lg.ctypes.data_as(FP))
        x = np.exp2(lg.astype(np.float64))
        ks = stats.kstest(special.gammainc(shape, x), "uniform")
        assert ks.statistic < 0.004, (shape, ks)          # 1.36 / sqrt(n) = 0.0021 is the 5 % critical value
        assert abs(x.mean() - shape) < 6 * np.sqrt(shape / n) + 1e-4, (shape, x.mean())
        assert abs(x.var() - shape) < 0.05 * shape + 0.01, (shape, x.var())
        sd = np.sqrt(special.polygamma(1, shape) / n)
        assert abs(lg.astype(np.float64).mean() * np.log(2) - special.digamma(shape)) < 6 * sd, shape
    # two draws share a Philox block (k and k ^ 32): they must still be uncorrelated
    lg = np.zeros(1 << 16, dtype=np.float32)
    oracle.oracle_gamma_log2_many(3, 1, 0, len(lg), 123, 7, 0.03, lg.ctypes.data_as(FP))
    a = lg.reshape(-1, 2, 32)
    assert abs(np.corrcoef(a[:, 0, :].ravel(), a[:, 1, :].ravel())[0, 1]) < 0.02


def test_f32_log2_exp2_accuracy(oracle):
    """lg2_det / ex2_det (tables + fixed polynomials over fmaf, so that CPU and GPU agree bit for bit) against libm."""
    FP = C.POINTER(C.c_float)
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-87, 88, 100000)), 1.0 - np.arange(1, 5000) * 2.0 ** -24, 1.0 + np.arange(5000) * 2.0 ** -23]).astype(np.float32)
    lg, ex = np.zeros_like(xs), np.zeros_like(xs)
    oracle.oracle_math32_many(xs.ctypes.data_as(FP), len(xs), lg.ctypes.data_as(FP), ex.ctypes.data_as(FP))
    ref = np.log2(xs.astype(np.float64))
    assert np.all(np.abs(lg - ref) <= 2.0 ** -25 + 2.0 ** -23 * np.abs(ref))
    near = np.abs(xs - 1) < 0.01
    assert np.all(np.abs(lg[near] - ref[near]) <= 1.5e-7 * np.abs(ref[near]) + 1e-38)   # relative accuracy where the largest draws live
    ts = rng.uniform(-125, 0, 100000).astype(np.float32)
    oracle.oracle_math32_many(ts.ctypes.data_as(FP), len(ts), lg[:len(ts)].ctypes.data_as(FP), ex[:len(ts)].ctypes.data_as(FP))
    r = np.exp2(ts.astype(np.float64))
    assert np.all(np.abs(ex[:len(ts)] - r) <= 1.2e-7 * r)
    one = np.array([0.0, -126.0, -1.0], dtype=np.float32)
    oracle.oracle_math32_many(one.ctypes.data_as(FP), 3, lg[:3].ctypes.data_as(FP), ex[:3].ctypes.data_as(FP))
    assert ex[:3].tolist() == [1.0, 0.0, 0.5]


def test_tie_break_is_uniform(oracle):
    """TieBreak::Philox: floor(u n / 2^32) over the tied children, one draw per (simulation, depth)."""
    for n in (2, 7, 225):
        picks = np.array([oracle.oracle_tie_pick(s, s % 5, 0xC6, 3, n) for s in range(20000)])
        assert picks.min() == 0 and picks.max() == n - 1
        cnt = np.bincount(picks, minlength=n)
        assert np.all(np.abs(cnt - 20000 / n) < 6 * np.sqrt(20000 / n) + 1)


def test_first_level_visits_are_not_biased_to_the_top_left(oracle):
    """With the uniform tie-break the second visit of every fresh node goes to a uniformly chosen child, not to the first
    empty cell: over many games the root's first visited child is spread over the board."""
    first_children = []
    for g in range(300):
        p = c6lib.OracleParams(2, 0.25, 0.03, 1.0, 0xC6, 1, c6lib.ORACLE_TIE[1], 0, 1, 1)
        out = c6lib.OracleGameOut()
        oracle.oracle_self_play(None, C.byref(p), g, C.byref(out), None, None, None)
        first_children.append(out.moves[0])
    assert len(set(first_children)) > 150 and min(first_children) < 40 and max(first_children) > 180
    p = c6lib.OracleParams(2, 0.25, 0.03, 1.0, 0xC6, 1, c6lib.ORACLE_TIE[0], 0, 1, 1)
    out = c6lib.OracleGameOut()
    oracle.oracle_self_play(None, C.byref(p), 0, C.byref(out), None, None, None)
    assert out.moves[0] == 0  # lowest-index mode: 2 simulations always end on cell (0, 0)


def test_random_evaluator_twin_is_iid_uniform(oracle):
    """The RandomEvaluator twin (alphazero_policy/mod.rs:137-169: every value and policy entry ~ U(-1, 1)): the eight draws of
    every empty cell — sixteen 24-bit draws out of three Philox blocks per pair of cells — are uniform on [-1, 1), uncorrelated
    within a cell, between the two cells of a pair and between planes; occupied cells get none."""
    from scipy import stats
    FP = C.POINTER(C.c_float)
    rng = np.random.default_rng(5)
    board = np.zeros(225, dtype=np.int8)
    board[rng.choice(225, size=40, replace=False)] = rng.choice([-1, 1], size=40)
    sims = 600
    pol = np.zeros((sims, 8, 225), dtype=np.float32)
    val = np.zeros((sims, 8), dtype=np.float32)
    for s in range(sims):
        oracle.oracle_random_evaluator(0xC6, 3, s, board.ctypes.data_as(C.POINTER(C.c_int8)), val[s].ctypes.data_as(FP), pol[s].ctypes.data_as(FP))
    # plane 6 is the leaf itself (rotate_left four times): original-frame draw number 6 of every cell
    ident = pol[:, 6, :]
    empty = board == 0
    assert not ident[:, ~empty].any()
    x = ident[:, empty].astype(np.float64)
    assert x.min() >= -1.0 and x.max() < 1.0
    assert stats.kstest((x.ravel() + 1) / 2, "uniform").statistic < 0.006
    assert abs(x.mean()) < 0.01 and abs(x.var() - 1 / 3) < 0.01
    # the two cells of a rank pair (k, k + 32) share Philox blocks: still uncorrelated; so are different planes of one cell
    ranks = np.flatnonzero(empty)
    a, b = ident[:, ranks[:32]].ravel(), ident[:, ranks[32:64]].ravel()
    assert abs(np.corrcoef(a, b)[0, 1]) < 0.02
    # every plane is a permutation of its own draw: compare plane 7 (flip of the identity) mapped back against plane 6
    p7 = pol[:, 7, :].reshape(sims, 15, 15)[:, :, ::-1].reshape(sims, 225)
    assert abs(np.corrcoef(p7[:, empty].ravel(), ident[:, empty].ravel())[0, 1]) < 0.02
    assert stats.kstest((val.ravel().astype(np.float64) + 1) / 2, "uniform").statistic < 0.03
