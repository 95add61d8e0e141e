"""The `pyconnect6.pyconnect6` native-module surface of the reference
(Connect6/pybind/libpyconnect6/src/lib.rs:35-85), served by the B200 engine through its C-ABI.

  self_play(obj, num_simulation, epsilon, dirichlet_alpha, c_puct, debug, num_game_thread)
  play_with(obj, num_simulation, epsilon, dirichlet_alpha, c_puct)
  test_echo_pyeval(obj, player, boards)

`obj` is the evaluator callable.  PyEval of the reference (pybind/pyeval/mod.rs:42-111) turns each
board into a list of ints, calls `obj(turn, boards)` under the GIL and parses `(values, policies)`
from any iterables of floats; the same conversion happens here inside a C callback that the engine
invokes at most twice per simulation wave (ctypes releases the GIL around the native call and takes
it back for the callback, like `py.allow_threads` + `Python::acquire_gil`, lib.rs:130, pyeval/mod.rs:70).
"""
import ctypes as C
import os

import numpy as np

from . import _capi

CELLS = 225


class EvaluatorError(RuntimeError):
    """The Python evaluator raised or returned malformed output (the reference panics here:
    alphazero_policy/mod.rs:378-380 "couldn't get value, prob from pyobject")."""


def _make_eval_callback(obj, on_error):
    """PyEval::eval twin (pyeval/mod.rs:66-110) as a C `Callback` (connect6.hpp:14)."""

    def cb(player, values_p, policies_p, length):
        try:
            pol = np.ctypeslib.as_array(policies_p, shape=(length, CELLS))
            val = np.ctypeslib.as_array(values_p, shape=(length,))
            boards = pol.astype(np.int64).tolist()  # list of lists of ints, as pylist_from_multiple builds
            res = obj(int(player), boards)
            v, p = res
            v = np.asarray(v, dtype=np.float32).reshape(-1)
            p = np.asarray(p, dtype=np.float32)
            if v.shape[0] != length or p.shape != (length, CELLS):
                raise ValueError(f"evaluator returned shapes {v.shape}, {p.shape} for {length} boards")
            val[:] = v
            pol[:] = p
        except BaseException as e:  # noqa: BLE001 - must not unwind through the C frames
            on_error(e)

    return _capi.CALLBACK(cb)


def _expand(record):
    """packed record -> (winner, [(turn, board[225], (row, col)), ...]) (pybind_impl/mod.rs:67-106)."""
    board = [0] * CELLS
    path = []
    for m in range(record.n_moves):
        turn = -1 if m == 0 else (1 if ((m - 1) // 2) % 2 == 0 else -1)
        cell = record.moves[m]
        path.append((turn, list(board), (cell // 15, cell % 15)))
        board[cell] = turn
    return int(record.winner), path


def self_play(obj, num_simulation, epsilon, dirichlet_alpha, c_puct, debug, num_game_thread):
    L = _capi.product()
    n = int(num_game_thread)
    if n <= 0:
        return ()
    err = []
    handle = [None]

    def on_error(e):
        err.append(e)
        if handle[0]:
            L.c6_engine_cancel(handle[0])

    cb = _make_eval_callback(obj, on_error)
    seed = int(os.environ.get("C6_SEED", "0"), 0) if os.environ.get("C6_SEED") else int.from_bytes(os.urandom(4), "little")
    cfg = _capi.make_config(num_simulation=int(num_simulation), epsilon=float(epsilon), dirichlet_alpha=float(dirichlet_alpha),
                            c_puct=float(c_puct), seed=seed, device=int(os.environ.get("C6_DEVICE", "0")),
                            slots=min(n, int(os.environ.get("C6_MAX_CONCURRENT", "8192"))),
                            log2_records=int(os.environ.get("C6_LOG2_RECORDS", "18")), evaluator=_capi.EVAL_CALLBACK)
    e = L.c6_engine_create(C.byref(cfg))
    if not e:
        raise RuntimeError(L.c6_last_error().decode())
    handle[0] = e
    try:
        if L.c6_engine_start(e, 0, n) != 0:
            raise RuntimeError(L.c6_last_error().decode())
        rc = L.c6_engine_run_callback(e, cb)
        if err:
            raise EvaluatorError("alpha_zero::expand couldn't get value, prob from pyobject") from err[0]
        if rc != 0:
            raise RuntimeError(L.c6_last_error().decode())
        rec = (_capi.Record * n)()
        if L.c6_engine_records(e, rec, n) != 0:
            raise RuntimeError(L.c6_last_error().decode())
    finally:
        handle[0] = None
        L.c6_engine_destroy(e)
    results = [_expand(rec[i]) for i in range(n)]
    if debug:
        for i, (winner, path) in enumerate(results):
            _print_game(path, winner) if n == 1 else print(f"run: {i}")
    return results[0] if n == 1 else tuple(results)


def default_policy_play(num_iter=50, num_game=1, seed=None, first_game_id=0, device=None):
    """`Agent::new(&mut DefaultPolicy::with_num_iter(num_iter)).play()` for num_game games on the GPU (pure MCTS with
    random rollouts, libconnect6/src/policy/default_policy/mod.rs:94-271; not exported by the reference's pybind).
    Returns a tuple of (winner, path) in the format of self_play."""
    L = _capi.product()
    n = int(num_game)
    if n <= 0:
        return ()
    if seed is None:
        seed = int(os.environ.get("C6_SEED", "0"), 0) if os.environ.get("C6_SEED") else int.from_bytes(os.urandom(4), "little")
    if device is None:
        device = int(os.environ.get("C6_DEVICE", "0"))
    rec = (_capi.Record * n)()
    if L.c6_default_policy_play(int(num_iter), int(seed), int(first_game_id), n, 0, int(device), rec, None, None, None) != 0:
        raise RuntimeError(L.c6_last_error().decode())
    return tuple(_expand(rec[i]) for i in range(n))


def _print_game(path, winner):
    """Agent::debug transcript (agent_impl/mod.rs:127-129,147-161; game_impl/mod.rs:217-237)."""
    def show(board):
        print("0 A B C D E F G H I J K L M N O")
        for r in range(15):
            print(chr(0x61 + r) + " " + " ".join("X" if board[r * 15 + c] < 0 else "O" if board[r * 15 + c] > 0 else "_" for c in range(15)) + " ")
    board = [0] * CELLS
    for m, (turn, board, (row, col)) in enumerate(path):
        show(board)
        remain = 0 if m == 0 else 1 - ((m - 1) % 2)
        print(f"{'Black' if turn < 0 else 'White'} ({chr(0x61 + row)}, {chr(0x41 + col)}), remain {remain}, 0.0 elapsed")
    if path:
        board = list(path[-1][1])
        board[path[-1][2][0] * 15 + path[-1][2][1]] = path[-1][0]
    show(board)


def play_with(obj, num_simulation, epsilon, dirichlet_alpha, c_puct):
    L = _capi.product()
    err = []
    cb = _make_eval_callback(obj, err.append)
    L.cpp_play_with.restype = _capi.PlayResult
    L.cpp_play_with.argtypes = [_capi.CALLBACK, _capi.ALLOC_PATH, C.c_int, C.c_float, C.c_double, C.c_float]
    r = L.cpp_play_with(cb, _capi.alloc_path, int(num_simulation), float(epsilon), float(dirichlet_alpha), float(c_puct))
    if err:
        raise EvaluatorError("alpha_zero::expand couldn't get value, prob from pyobject") from err[0]
    path = []
    for m in range(r.len):
        p = r.paths[m]
        path.append((int(p.turn), [int(p.board[a][b]) for a in range(15) for b in range(15)], (int(p.row), int(p.col))))
    return int(r.winner), path


def test_echo_pyeval(obj, player, boards):
    """lib.rs:46-82: run `obj` through the evaluator conversion on `boards` (flat, len x 225 ints)
    and return `(values, policies)` as lists.  Goes through the C-ABI's `test_echo_cppeval`."""
    L = _capi.product()
    boards = [int(b) for b in boards]
    n = len(boards) // CELLS
    err = []
    cb = _make_eval_callback(obj, err.append)
    arr = (C.c_int * len(boards))(*boards)
    out = L.test_echo_cppeval(int(player), arr, n, cb, _capi.alloc_float)
    if err:
        raise EvaluatorError("evaluator failed") from err[0]
    flat = [out.vec[i] for i in range(out.len)]
    values = flat[:n]
    policies = [flat[n + i * CELLS:n + (i + 1) * CELLS] for i in range(n)]
    return values, policies


test_echo_pyeval.__test__ = False  # a library function that happens to be called test_*, not a pytest test
