"""pyconnect6 — Python face of the B200 Connect6 self-play engine.

Same surface as the reference package (Connect6/pybind/pyconnect6/__init__.py:5-79): `board_size`,
`self_play`, `play_with`, `default_param`, `param_to_tuple`, `dump_param`, `load_param`, and the
native-module namespace `pyconnect6.pyconnect6` (`self_play`, `play_with`, `test_echo_pyeval`,
Connect6/pybind/libpyconnect6/src/lib.rs:35-85).  The search runs on the GPU through the C-ABI of
libcppconnect6.so (include/connect6_b200.h); there is no CPU fallback.

Evaluator protocol (unchanged): `policy(turn, boards) -> (values, probs)` with `turn` in {-1, +1},
`boards` a list of `len` boards of 225 ints in {-1, 0, +1}, `values` a sequence of `len` floats and
`probs` a `len` x 225 sequence.  The engine batches the leaves of every concurrent game into one call
per side to move, so `len` is 8 x (number of games whose leaf has that side to move) instead of 8.
"""
import json

from . import pyconnect6  # noqa: F401  (the native-module namespace of the reference)
from . import replay, sharded  # noqa: F401

_PARAM_ORDER = ("num_simulation", "epsilon", "dirichlet_alpha", "c_puct", "debug", "num_game_thread")


def board_size():
    return 15


def default_param():
    """HyperParameter::default() (libconnect6 alphazero_policy/mod.rs:97-108) plus the binding's
    `debug` / `num_game_thread` (pyconnect6/__init__.py:47-56)."""
    return dict(zip(_PARAM_ORDER, (800, 0.25, 0.03, 1, False, 1)))


def param_to_tuple(param):
    return tuple(param[k] for k in _PARAM_ORDER)


def self_play(policy, param=None):
    """`num_game_thread` games of AlphaZero self-play, all concurrently on the GPU.

    Returns `(winner, [(turn, board, (row, col)), ...])` for one game, a tuple of those otherwise
    (libpyconnect6/src/lib.rs:104-152)."""
    return pyconnect6.self_play(policy, *param_to_tuple(default_param() if param is None else param))


def play_with(policy, param=None):
    """AlphaZero (Black, GPU search) against moves typed on stdin as "aB" (White)."""
    return pyconnect6.play_with(policy, *param_to_tuple(default_param() if param is None else param)[:4])


def dump_param(path, param=None):
    with open(path + "_mcts.json", "w") as f:
        json.dump(default_param() if param is None else param, f)


def load_param(path):
    with open(path + "_mcts.json") as f:
        return json.load(f)
