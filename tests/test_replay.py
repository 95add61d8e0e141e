"""CPU suite, part 4: packed records -> training tuples (the replay-buffer feed next to the hot path)."""
import numpy as np

import c6lib  # noqa: F401
from pyconnect6 import replay
from test_sharding_gloo import fake_records


def test_training_tuples_equal_the_nested_reference_form():
    rec, _ = fake_records(3, 40)
    winners, players, boards, positions = replay.training_tuples(rec)
    flat = []
    for win, path in replay.as_game_results(rec):  # Buffer.push_game (py_weighted/buffer.py:22-31)
        for player, board, (row, col) in path:
            flat.append((win, player, board, row * 15 + col))
    assert len(flat) == len(winners) == int(rec[:, 1].sum())
    for i, (win, player, board, pos) in enumerate(flat):
        assert (winners[i], players[i], positions[i]) == (win, player, pos)
        assert boards[i].tolist() == board


def test_empty_and_single_stone_records():
    rec = np.zeros((2, 240), dtype=np.uint8)
    rec[1, 1] = 1
    rec[1, 8] = 112
    w, p, b, pos = replay.training_tuples(rec)
    assert len(w) == 1 and p[0] == -1 and pos[0] == 112 and not b[0].any()
    assert replay.as_game_results(rec)[0] == (0, [])


def test_stone_colours_follow_the_turn_rule():
    assert replay.stone_colours(7).tolist() == [-1, 1, 1, -1, -1, 1, 1]
