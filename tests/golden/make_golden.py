#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/.

  python tests/golden/make_golden.py

1. reference_known_answers.json — the reference's OWN deterministic known-answer vectors for the path,
   transcribed as data (inputs + the value the reference test asserts), with the file:line of every
   assertion.  The reference is Rust and cannot be executed here (no cargo/rustc in the image), so these
   transcriptions are what pins the oracle to the reference.
     LC = /root/reference/Connect6/libconnect6/src
2. selfplay_traces.json — outputs of the pinned oracle (oracle/c6_oracle.hpp) for seeded self-play:
   move list, winner and per-move root statistics.  The CPU suite checks the oracle still reproduces
   them (drift guard); the GPU suite checks the CUDA engine against them without needing the oracle.
3. b1_hand_trace.json — the hand-derived trace of SURVEY.md Appendix B.1 (independent of any code).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import c6lib  # noqa: E402

N = 15
B, W, E = -1, 1, 0


def board():
    return [[0] * N for _ in range(N)]


def snap(t):
    return [v for row in t for v in row]


def win_vectors():
    out = []

    def add(t, expect, cite):
        out.append({"board": snap(t), "winner": expect, "cite": cite})

    # LC/game/search_winner/tests.rs:102-139 test_base_search (one table mutated through 8 assertions)
    t = board()
    for i in range(5):
        t[0][i] = B
    add(t, E, "search_winner/tests.rs:108")
    t[0][5] = B
    add(t, B, "search_winner/tests.rs:111")
    for i in range(5):
        t[5 - i][0] = W
    add(t, B, "search_winner/tests.rs:117")
    t[0][0] = W
    add(t, W, "search_winner/tests.rs:120")
    for i in range(5):
        t[5 - i][5 - i] = B
    add(t, W, "search_winner/tests.rs:126")
    t[0][0] = B
    add(t, B, "search_winner/tests.rs:129")
    for i in range(5):
        t[5 - i][i] = W
    add(t, B, "search_winner/tests.rs:135")
    t[0][5] = W
    add(t, W, "search_winner/tests.rs:138 (both colours own a six: scan order decides)")
    # :141-182 boundary_test! row => White, col => Black; bases 0 / 14, inits 0, 1, 9
    for base in (0, N - 1):
        for init in (0, 1, N - 6):
            t = board()
            for i in range(5):
                t[base][init + i] = W
            add(t, E, "search_winner/tests.rs:147 row")
            t[base][init + 5] = W
            add(t, W, "search_winner/tests.rs:150 row")
    for base in (0, N - 1):
        for init in (0, 1, N - 6):
            t = board()
            for i in range(5):
                t[init + i][base] = B
            add(t, E, "search_winner/tests.rs:157 col")
            t[init + 5][base] = B
            add(t, B, "search_winner/tests.rs:160 col")
    # :184-216 cross_test! right_down => White, left_down => Black
    for init in (0, 1, N - 6):
        t = board()
        for i in range(5):
            t[init + i][init + i] = W
        add(t, E, "search_winner/tests.rs:191 right_down")
        t[init + 5][init + 5] = W
        add(t, W, "search_winner/tests.rs:194 right_down")
    for init in (0, 1, N - 6):
        t = board()
        for i in range(5):
            t[init + i][N - 1 - init - i] = B
        add(t, E, "search_winner/tests.rs:201 left_down")
        t[init + 5][N - init - 6] = B
        add(t, B, "search_winner/tests.rs:204 left_down")
    return out


def corner_board():
    # the 8-stone corner pattern shared by augment_tests.rs:6-14,27-35,48-56,69-77
    t = board()
    t[0][0] = B
    t[0][1] = B
    t[0][N - 1] = W
    t[1][N - 1] = B
    t[N - 1][N - 1] = W
    t[N - 1][N - 2] = W
    t[N - 1][0] = B
    t[N - 2][0] = W
    return t


def symmetry_vectors():
    c = snap(corner_board())
    return {
        "board": c,
        # (row, col, expected) after the transform, augment_tests.rs:17-24 / 38-45 / 59-66 / 80-87
        "rotate_left": [[N - 1, 0, B], [N - 2, 0, B], [0, 0, W], [0, 1, B], [0, N - 1, W], [1, N - 1, W], [N - 1, N - 1, B], [N - 1, N - 2, W]],
        "rotate_right": [[N - 1, 0, W], [N - 2, 0, W], [0, 0, B], [0, 1, W], [0, N - 1, B], [1, N - 1, B], [N - 1, N - 1, W], [N - 1, N - 2, B]],
        "flip_vertical": [[N - 1, 0, W], [N - 1, 1, W], [0, 0, W], [1, 0, B], [0, N - 1, B], [0, N - 2, B], [N - 1, N - 1, B], [N - 2, N - 1, W]],
        "flip_horizontal": [[N - 1, 0, B], [N - 1, 1, B], [0, 0, B], [1, 0, W], [0, N - 1, W], [0, N - 2, W], [N - 1, N - 1, W], [N - 2, N - 1, B]],
        "cite": "LC/policy/alphazero_policy/augment_tests.rs:5-94",
    }


def scripted_game():
    # LC/agent/agent_impl/tests.rs:66-129 test_play: winner White, 11 path entries
    moves = [(0, 0), (0, 1), (0, 2), (1, 0), (2, 0), (0, 3), (0, 4), (3, 0), (4, 0), (0, 5), (0, 6)]
    turns = [B, W, W, B, B, W, W, B, B, W, W]
    return {"moves": moves, "turns": turns, "winner": W, "path_len": 11, "cite": "LC/agent/agent_impl/tests.rs:66-129"}


def rules_vectors():
    # LC/game/game_impl/tests.rs:15-39: (pos) -> SetResult{player, num_remain, position} / error strings
    return {
        "steps": [
            {"pos": [0, 0], "ok": True, "player": B, "num_remain": 0, "turn_after": W, "remain_after": 2},
            {"pos": [0, 0], "ok": False, "err": "already set position"},
            {"pos": [N, N], "ok": False, "err": "invalid position"},
        ],
        "cite": "LC/game/game_impl/tests.rs:15-39 (error strings: game_impl/mod.rs:64-100)",
    }


SELFPLAY_CASES = [
    # name, evaluator (1 = Philox RandomEvaluator twin, 0 = callback), callback, sims, eps, seed, first_id, n_games, max_moves,
    # tie_break (1 = uniform among exact ties from Philox, the engine default; 0 = lowest index)
    ("philox_2sims_noise_full_games", 1, None, 2, 0.25, 0xC6, 0, 4, 0, 1),
    ("philox_64sims_eps0", 1, None, 64, 0.0, 0xC6, 77, 3, 16, 1),
    ("philox_64sims_noise", 1, None, 64, 0.25, 0xC6, 77, 3, 16, 1),
    ("philox_64sims_noise_lowest_index", 1, None, 64, 0.25, 0xC6, 77, 3, 16, 0),
    ("det_40sims_eps0", 0, "det_eval_callback", 40, 0.0, 0xC6, 5, 3, 16, 1),
    ("det_40sims_eps0_lowest_index", 0, "det_eval_callback", 40, 0.0, 0xC6, 5, 3, 16, 0),
    ("neg_40sims_noise", 0, "neg_eval_callback", 40, 0.25, 0xC6, 5, 3, 16, 1),
    ("philox_800sims_reference_default", 1, None, 800, 0.25, 0xC6, 0, 1, 4, 1),
    # opt-in multi-leaf mode (virtual_loss = 4): not the reference's schedule; drift guard of the oracle twin + engine check
    ("philox_64sims_noise_multi_leaf4", 1, None, 64, 0.25, 0xC6, 77, 3, 16, 1, 4),
]


def selfplay_traces():
    O = c6lib.oracle()
    cases = []
    for name, ev, cbname, sims, eps, seed, first, n, mm, tie, *rest in SELFPLAY_CASES:
        vl = rest[0] if rest else 1
        games = []
        for g in range(n):
            p = c6lib.OracleParams(sims, eps, 0.03, 1.0, seed, 1 if eps != 0 else 0, c6lib.ORACLE_TIE[tie], 0, ev, mm, vl)
            out = c6lib.OracleGameOut()
            cv = np.zeros((225, 225), dtype=np.int32)
            rn = np.zeros(225, dtype=np.int32)
            rw = np.zeros(225, dtype=np.float32)
            cb = c6lib.oracle_callback(cbname) if cbname else None
            O.oracle_self_play(cb, C.byref(p), first + g, C.byref(out), cv.ctypes.data_as(C.POINTER(C.c_int32)),
                               rn.ctypes.data_as(C.POINTER(C.c_int32)), rw.ctypes.data_as(C.POINTER(C.c_float)))
            m = out.n_moves
            games.append({
                "game_id": first + g, "winner": out.winner, "finished": out.finished,
                "moves": [out.moves[i] for i in range(m)],
                "root_visit": rn[:m].tolist(),
                "root_qsum_bits": rw[:m].view(np.uint32).tolist(),
                # per move: checksum of the 225 root-child visit counts (exact sum of v*(cell+1)) and the argmax cell
                "child_visit_checksum": [int((cv[i].astype(np.int64) * np.arange(1, 226)).sum()) for i in range(m)],
                "sims": int(out.sims), "sum_ck": int(out.sum_ck), "sum_cleaf": int(out.sum_cleaf), "sum_depth": int(out.sum_depth),
            })
        cases.append({"name": name, "evaluator": ev, "callback": cbname, "num_simulation": sims, "epsilon": eps,
                      "dirichlet_alpha": 0.03, "c_puct": 1.0, "seed": seed, "first_id": first, "max_moves": mm, "tie_break": tie, "virtual_loss": vl, "games": games})
    return cases


def main():
    ka = {"win_detection": win_vectors(), "symmetries": symmetry_vectors(), "scripted_game": scripted_game(),
          "rules": rules_vectors()}
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(ka, f, separators=(",", ":"))
    with open(os.path.join(HERE, "selfplay_traces.json"), "w") as f:
        json.dump(selfplay_traces(), f, separators=(",", ":"))
    b1 = {"evaluator": "const (value .5, policy .25)", "num_simulation": 4, "epsilon": 0.0, "c_puct": 1.0,
          "moves_rc": [[0, 2], [0, 4], [0, 5], [0, 6], [0, 7], [0, 8], [0, 9]],
          "move1_root": {"visit": 4, "q_sum": 1.5}, "move2_root": {"visit": 5, "q_sum": 2.0},
          "cite": "SURVEY.md Appendix B.1 (derived by hand from alphazero_policy/mod.rs:243-452)"}
    with open(os.path.join(HERE, "b1_hand_trace.json"), "w") as f:
        json.dump(b1, f, separators=(",", ":"))
    print("golden fixtures written")


if __name__ == "__main__":
    main()
