#!/usr/bin/env python
"""Generates tests/golden/selfplay_headline.json: oracle traces of WHOLE games at the headline configuration
(BASELINE.json configs[1]: 800 simulations per stone, epsilon .25, alpha .03, RandomEvaluator twin, uniform tie-break) and of
64 whole games at 50 simulations.  The GPU suite replays them inside a 4,096-slot engine with slot refill, the golden game
ids scattered among filler games (tests/test_parity_gpu.py::test_headline_whole_games_in_full_engine).

  python tests/golden/make_headline.py          # ~2 minutes on 8 cores (lean cost model of the pinned oracle)
"""
import ctypes as C
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import c6lib  # noqa: E402

CASES = [
    # name, sims, eps, seed, total games queued in the engine, resident slots, golden game ids
    ("headline_800sims", 800, 0.25, 0xC6, 4608, 4096,
     [3, 257, 511, 1024, 1500, 2047, 2048, 3000, 3500, 4000, 4095, 4096, 4200, 4300, 4500, 4607]),
    ("whole_games_50sims", 50, 0.25, 0xC6, 5120, 4096, sorted(set(int(x) for x in np.linspace(0, 5119, 64)))),
]


def one_game(args):
    sims, eps, seed, gid = args
    O = c6lib.oracle()
    p = c6lib.OracleParams(sims, eps, 0.03, 1.0, seed, 1, c6lib.ORACLE_TIE[1], 0, 1, 0)
    out = c6lib.OracleGameOut()
    cv = np.zeros((225, 225), dtype=np.int32)
    rn = np.zeros(225, dtype=np.int32)
    rw = np.zeros(225, dtype=np.float32)
    O.oracle_self_play(None, C.byref(p), gid, C.byref(out), cv.ctypes.data_as(C.POINTER(C.c_int32)),
                       rn.ctypes.data_as(C.POINTER(C.c_int32)), rw.ctypes.data_as(C.POINTER(C.c_float)))
    m = out.n_moves
    return {"game_id": gid, "winner": out.winner, "moves": [out.moves[i] for i in range(m)], "root_visit": rn[:m].tolist(),
            "root_qsum_bits": rw[:m].view(np.uint32).tolist(),
            "child_visit_checksum": [int((cv[i].astype(np.int64) * np.arange(1, 226)).sum()) for i in range(m)],
            "sims": int(out.sims), "max_nodes": int(out.max_nodes)}


def main():
    c6lib.oracle()
    cases = []
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as ex:
        for name, sims, eps, seed, total, slots, ids in CASES:
            games = list(ex.map(one_game, [(sims, eps, seed, g) for g in ids]))
            cases.append({"name": name, "num_simulation": sims, "epsilon": eps, "dirichlet_alpha": 0.03, "c_puct": 1.0, "seed": seed,
                          "tie_break": 1, "total_games": total, "slots": slots, "games": games})
            print(name, len(games), "games,", sum(len(g["moves"]) for g in games), "stones")
    with open(os.path.join(HERE, "selfplay_headline.json"), "w") as f:
        json.dump(cases, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
