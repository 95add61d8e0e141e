"""CPU suite, part 2: the C-ABI boundary and the host-side mirror (no GPU compute is called).

  * the library loads and exports every symbol include/connect6_b200.h declares;
  * struct layouts are the reference's (connect6.hpp:26-42);
  * the reference's ABI echo helpers (ffi_test/mod.rs) and evaluator / policy conversions behave like
    the reference's own tests say (rawobj/tests.rs, cppeval/tests.rs:30-92, cpp_policy/tests.rs:41-58,
    test_cppbind/*.cpp, test_pybind.py:27-49);
  * cpp_play (host game loop over a user policy callback: no search in it) reproduces the reference's
    scripted game (agent_impl/tests.rs:66-129) and aborts on an invalid move like `.unwrap()` does.
"""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import c6lib

ROOT = c6lib.ROOT
GOLDEN = os.path.join(ROOT, "tests", "golden")


def header_symbols():
    text = open(os.path.join(ROOT, "include", "connect6_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"typedef[^;{]*(\{[^}]*\})?[^;]*;", "", text)  # structs and function-pointer typedefs
    names = re.findall(r"^[A-Za-z_][A-Za-z0-9_ \*]*?\b(\w+)\s*\([^;{]*\)\s*;", text, flags=re.M)
    return [n for n in names if not n.startswith("c6_alloc") and n not in ("c6_callback", "c6_policy_callback")]


def test_library_exports_every_declared_symbol(lib):
    names = header_symbols()
    assert len(names) >= 35, names
    for want in c6lib.REFERENCE_SYMBOLS:
        assert want in names
    for n in names:
        assert getattr(lib, n) is not None, n


def test_struct_layouts():
    assert C.sizeof(c6lib.Path) == 912 and c6lib.Path.row.offset == 904
    assert C.sizeof(c6lib.PlayResult) == 24 and c6lib.PlayResult.paths.offset == 8 and c6lib.PlayResult.len.offset == 16
    assert C.sizeof(c6lib.Vec) == 16
    assert C.sizeof(c6lib.Record) == 240


def test_raw_path_helpers(lib):
    p = lib.test_new_raw_path()
    assert (p.turn, p.row, p.col) == (0, 0, 0) and not np.ctypeslib.as_array(p.board).any()
    p = lib.test_with_raw_path()  # ffi_test/mod.rs:19-34, checked by test_cppbind/rawobj.cpp
    assert (p.turn, p.row, p.col) == (1, 0, 1)
    b = np.ctypeslib.as_array(p.board)
    assert all(b[i][j] == (i * 15 + j) % 3 - 1 for i in range(15) for j in range(15))
    rng = np.random.default_rng(1)
    board = rng.integers(-1, 2, size=225).astype(np.int32)
    p = lib.test_echo_raw_path(-1, board.ctypes.data_as(C.POINTER(C.c_int)), 3, 9)
    assert (p.turn, p.row, p.col) == (-1, 3, 9)
    assert np.array_equal(np.ctypeslib.as_array(p.board).reshape(-1), board)


def test_raw_play_result_helpers(lib):
    r = lib.test_with_raw_play_result(c6lib.alloc_path)  # ffi_test/mod.rs:61-86
    assert (r.winner, r.len) == (-1, 10)
    for i in range(10):
        p = r.paths[i]
        assert (p.turn, p.row, p.col) == (-1 if i % 2 == 0 else 1, i, i + 1)
        b = np.ctypeslib.as_array(p.board)
        want = np.zeros((15, 15), dtype=np.int32)
        for j in range(i + 1):
            want[j][j] = (i + j) % 3 - 1
        assert np.array_equal(b, want)
    paths = (c6lib.Path * 3)()
    for i in range(3):
        paths[i].turn, paths[i].row, paths[i].col = (1, -1, 0)[i], i, 2 * i
        paths[i].board[i][i] = 1
    e = lib.test_echo_raw_play_result(1, paths, 3, c6lib.alloc_path)
    assert (e.winner, e.len) == (1, 3)
    for i in range(3):
        assert (e.paths[i].turn, e.paths[i].row, e.paths[i].col) == ((1, -1, 0)[i], i, 2 * i)
        assert e.paths[i].board[i][i] == 1


def test_raw_vec_helpers(lib):
    v = lib.test_with_raw_vec(c6lib.alloc_int)
    assert [v.vec[i] for i in range(v.len)] == [0, 1, 2, 3, 4, 5]
    src = (C.c_int * 5)(5, -4, 3, -2, 1)
    v = lib.test_echo_raw_vec(src, 5, c6lib.alloc_int)
    assert [v.vec[i] for i in range(v.len)] == [5, -4, 3, -2, 1]
    v = lib.test_echo_raw_vec(src, 0, c6lib.alloc_int)  # empty input
    assert v.len == 0


def test_cppeval_echo(lib):
    """cppeval/tests.rs:63-92 / test_cppbind/cppeval.cpp: values[i] = i, policy = 2 x board."""
    def cb(player, values, policies, n):
        for i in range(n):
            values[i] = float(i) + player
        for i in range(n * 225):
            policies[i] = policies[i] * 2.0
    rng = np.random.default_rng(2)
    n = 7
    boards = rng.integers(-1, 2, size=n * 225).astype(np.int32)
    out = lib.test_echo_cppeval(1, boards.ctypes.data_as(C.POINTER(C.c_int)), n, c6lib.CALLBACK(cb), c6lib.alloc_float)
    assert out.len == n + n * 225
    flat = np.array([out.vec[i] for i in range(out.len)], dtype=np.float32)
    assert np.array_equal(flat[:n], np.arange(n, dtype=np.float32) + 1)
    assert np.array_equal(flat[n:], 2.0 * boards.astype(np.float32))


def test_cpp_policy_conversion(lib):
    """cpp_policy/tests.rs:41-58: the callback sees the float board; {-1,-1} means no move."""
    seen = {}

    def cb(board, pos):
        seen["board"] = [board[i] for i in range(225)]
        seen["pre"] = (pos[0], pos[1])
        pos[0], pos[1] = 4, 11
    board = np.zeros(225, dtype=np.float32)
    board[7], board[100] = -1.0, 1.0
    v = lib.test_cpp_policy(board.ctypes.data_as(C.POINTER(C.c_float)), c6lib.POLICY_CALLBACK(cb), c6lib.alloc_int)
    assert [v.vec[i] for i in range(v.len)] == [4, 11]
    assert seen["pre"] == (-1, -1) and seen["board"] == board.tolist()
    v = lib.test_cpp_policy(board.ctypes.data_as(C.POINTER(C.c_float)), c6lib.POLICY_CALLBACK(lambda b, p: None), c6lib.alloc_int)
    assert v.len == 0


def scripted_policy(moves):
    it = iter(moves)

    def cb(board, pos):
        m = next(it, None)
        if m is not None:
            pos[0], pos[1] = m
    return c6lib.POLICY_CALLBACK(cb)


def test_cpp_play_scripted_game(lib):
    g = json.load(open(os.path.join(GOLDEN, "reference_known_answers.json")))["scripted_game"]
    v = lib.cpp_play(scripted_policy([tuple(m) for m in g["moves"]]), c6lib.alloc_path, c6lib.alloc_result, False, 1)
    assert v.len == 1
    r = v.vec[0]
    assert (r.winner, r.len) == (g["winner"], g["path_len"])
    board = np.zeros((15, 15), dtype=np.int32)
    for k in range(r.len):
        p = r.paths[k]
        assert p.turn == g["turns"][k] and [p.row, p.col] == list(g["moves"][k])
        assert np.array_equal(np.ctypeslib.as_array(p.board), board)
        board[p.row][p.col] = p.turn


def test_cpp_play_no_move_ends_with_no_winner(lib):
    v = lib.cpp_play(scripted_policy([(0, 0), (5, 5)]), c6lib.alloc_path, c6lib.alloc_result, False, 1)
    assert (v.vec[0].winner, v.vec[0].len) == (0, 2)
    v = lib.cpp_play(scripted_policy([]), c6lib.alloc_path, c6lib.alloc_result, False, 0)
    assert v.len == 0


def run_child(code):
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, "tests"))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)


def test_cpp_play_invalid_move_aborts_like_unwrap():
    """lib.rs:40: Agent::play Err -> .unwrap() -> abort.  Same here, with the reference's error text."""
    r = run_child("import c6lib, test_abi as t; L = c6lib.product(); "
                  "L.cpp_play(t.scripted_policy([(0,0),(0,0)]), c6lib.alloc_path, c6lib.alloc_result, False, 1)")
    assert r.returncode != 0 and "already set position" in r.stderr


def test_self_play_without_gpu_fails_loudly():
    """No CPU fallback: without a CUDA device the reference entry point aborts with a message and the
    extended interface returns an error."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    L = c6lib.product()
    cfg = c6lib.make_config(num_simulation=2, evaluator=c6lib.EVAL_PHILOX, slots=1)
    assert not L.c6_engine_create(C.byref(cfg))
    assert b"no CUDA device" in L.c6_last_error()
    r = run_child("import c6lib, ctypes as C; L = c6lib.product(); "
                  "cb = c6lib.CALLBACK(lambda *a: None); L.cpp_self_play(cb, c6lib.alloc_path, c6lib.alloc_result, 2, .25, .03, 1., False, 1)")
    assert r.returncode != 0 and "no CUDA device" in r.stderr


def test_missing_library_fails_loudly():
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import c6lib; c6lib.product()" % os.path.join(ROOT, "tests")],
                       capture_output=True, text=True, env=dict(os.environ, C6_LIB="/nonexistent/libcppconnect6.so"), timeout=120)
    assert r.returncode != 0 and "not built" in r.stderr


def test_pyconnect6_surface_and_echo():
    """test_pybind.py:27-49 (test_echo_pyeval) through the C-ABI, plus the wrapper's parameter helpers."""
    import pyconnect6
    assert pyconnect6.board_size() == 15
    p = pyconnect6.default_param()
    assert pyconnect6.param_to_tuple(p) == (800, 0.25, 0.03, 1, False, 1)
    rng = np.random.default_rng(3)
    turn = int(rng.integers(3)) - 1
    n = 10
    boards = [int(x) - 1 for x in rng.integers(3, size=n * 225)]

    def double_policy(turn, board):
        assert isinstance(board, list) and isinstance(board[0], list) and isinstance(board[0][0], int)
        return [i + turn for i in range(len(board))], [[i * 2 for i in x] for x in board]

    value, policy = pyconnect6.pyconnect6.test_echo_pyeval(double_policy, turn, boards)
    norm = {-1: -1, 0: 0, 1: 1}[turn]
    idx = 0
    for i in range(n):
        assert value[i] == i + norm
        for x in policy[i]:
            assert x == 2 * boards[idx]
            idx += 1
    with pytest.raises(pyconnect6.pyconnect6.EvaluatorError):
        pyconnect6.pyconnect6.test_echo_pyeval(lambda t, b: ([0.0], [[0.0]]), turn, boards)


def test_param_dump_load(tmp_path):
    import pyconnect6
    p = pyconnect6.default_param()
    p["num_simulation"] = 17
    pyconnect6.dump_param(str(tmp_path / "ckpt"), p)
    assert os.path.exists(tmp_path / "ckpt_mcts.json")
    assert pyconnect6.load_param(str(tmp_path / "ckpt")) == p


@pytest.mark.skipif(not os.path.isdir("/root/reference/Connect6/cppbind/test_cppbind"), reason="reference tree not mounted")
def test_reference_catch2_suite_against_our_library(tmp_path):
    """The reference's own Catch2 ABI tests (rawobj / cppeval / cpp_policy), compiled unmodified against
    the reference's own header, linked with our library: every case except Connect6::self_play, which needs
    the GPU (tests/test_parity_gpu.py runs the prebuilt binary in full there)."""
    exe = build_reference_catch2()
    r = subprocess.run([exe, '~"Check Connect6::self_play"'], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(c6lib.LIB_PATH)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "All tests passed" in r.stdout


CATCH2_EXE = os.path.join(ROOT, "tests", "_build", "test_cppbind")


def build_reference_catch2():
    """Compiles the reference's test sources IN PLACE (nothing is copied into the repo); the binary goes
    to tests/_build/ (git-ignored, travels to the GPU box with the snapshot)."""
    ref = "/root/reference"
    tdir = os.path.join(ref, "Connect6/cppbind/test_cppbind")
    exe = CATCH2_EXE
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(c6lib.LIB_PATH):
        return exe
    srcs = [os.path.join(tdir, f) for f in sorted(os.listdir(tdir)) if f.endswith(".cpp")]
    cmd = ["g++", "-std=c++17", "-O1", "-DCATCH_CONFIG_NO_POSIX_SIGNALS", "-I", os.path.join(ref, "Connect6/cppbind/headers"),
           "-I", os.path.join(ref, "External/Catch2/single_include"), *srcs, "-o", exe,
           "-L", os.path.dirname(c6lib.LIB_PATH), "-lcppconnect6", "-Wl,-rpath,$ORIGIN/../../alphazero-connect6_b200/lib", "-pthread", "-ldl"]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=600)
    return exe


SAMPLE_EXE = os.path.join(ROOT, "tests", "_build", "sample_cpp")


def build_reference_sample():
    """Connect6/sample/cpp/main.cpp (config 1 of BASELINE.json in its C++ form), compiled in place against the
    reference's header and linked with our library."""
    ref = "/root/reference"
    os.makedirs(os.path.dirname(SAMPLE_EXE), exist_ok=True)
    if os.path.exists(SAMPLE_EXE) and os.path.getmtime(SAMPLE_EXE) >= os.path.getmtime(c6lib.LIB_PATH):
        return SAMPLE_EXE
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ref, "Connect6/cppbind/headers"), os.path.join(ref, "Connect6/sample/cpp/main.cpp"),
           "-o", SAMPLE_EXE, "-L", os.path.dirname(c6lib.LIB_PATH), "-lcppconnect6", "-Wl,-rpath,$ORIGIN/../../alphazero-connect6_b200/lib",
           "-pthread", "-ldl"]
    subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=600)
    return SAMPLE_EXE


@pytest.mark.skipif(not os.path.isdir("/root/reference/Connect6/sample/cpp"), reason="reference tree not mounted")
def test_reference_cpp_sample_links_against_our_library():
    assert os.path.exists(build_reference_sample())


# ---------------------------------------------------------------------------------------------
# More of the reference's own acceptance material, built IN PLACE where /root/reference is mounted (nothing is copied into
# the repo); outputs go to tests/_build/ (git-ignored, travels to the GPU box with the snapshot) and the GPU suite runs them.
# ---------------------------------------------------------------------------------------------
BUILD_DIR = os.path.join(ROOT, "tests", "_build")
# (.pycode, not .pyc: the gpurun snapshot leaves *.pyc files behind)
REF_PY = {"ref_test_pybind.pycode": "/root/reference/Connect6/pybind/test_pybind/test_pybind.py",
          "ref_random_policy.pycode": "/root/reference/Connect6/sample/python/random_policy.py"}


def build_reference_python():
    """The reference's Python acceptance test and sample, byte-compiled from where they lie (PB/test_pybind/test_pybind.py:16-49,
    Connect6/sample/python/random_policy.py:28-37): the byte-code files are what the GPU box executes against our pyconnect6."""
    import py_compile
    os.makedirs(BUILD_DIR, exist_ok=True)
    out = []
    for name, src in REF_PY.items():
        out.append(py_compile.compile(src, cfile=os.path.join(BUILD_DIR, name), doraise=True))
    return out


DEVICE_OVERLOAD_EXE = os.path.join(BUILD_DIR, "device_overload")
HOST_MIRROR_EXE = os.path.join(BUILD_DIR, "host_mirror")


def _stale(exe, *srcs):
    return not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(s) for s in (c6lib.LIB_PATH, *srcs))


def build_device_overload():
    """tests/cpp/device_overload.cu: the reference's own connect6.hpp (in place) + include/connect6_device.hpp, the opt-in
    device-pointer Connect6::self_play overload, with a CUDA-kernel evaluator."""
    src = os.path.join(ROOT, "tests", "cpp", "device_overload.cu")
    os.makedirs(BUILD_DIR, exist_ok=True)
    if _stale(DEVICE_OVERLOAD_EXE, src, os.path.join(ROOT, "include", "connect6_device.hpp")):
        cmd = ["nvcc", "-std=c++17", "-O1", "-gencode", "arch=compute_100a,code=sm_100a", "-I", "/root/reference/Connect6/cppbind/headers",
               "-I", os.path.join(ROOT, "include"), src, "-o", DEVICE_OVERLOAD_EXE, "-L", os.path.dirname(c6lib.LIB_PATH), "-lcppconnect6",
               "-Xlinker", "-rpath=$ORIGIN/../../alphazero-connect6_b200/lib"]
        subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=600)
    return DEVICE_OVERLOAD_EXE


def build_host_mirror():
    """tests/cpp/host_mirror.cpp: RandomEvaluator, AsyncAgent (generic and batched) and AlphaZero of the host-side mirror."""
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror.cpp")
    os.makedirs(BUILD_DIR, exist_ok=True)
    csrc = os.path.join(c6lib.PKG, "csrc")
    if _stale(HOST_MIRROR_EXE, src, os.path.join(csrc, "connect6_alphazero.hpp"), os.path.join(csrc, "connect6_host.hpp")):
        cmd = ["g++", "-std=c++17", "-O1", "-I", csrc, "-I", "/usr/local/cuda/include", src, "-o", HOST_MIRROR_EXE,
               "-L", os.path.dirname(c6lib.LIB_PATH), "-lcppconnect6", "-Wl,-rpath,$ORIGIN/../../alphazero-connect6_b200/lib", "-pthread", "-ldl"]
        subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=600)
    return HOST_MIRROR_EXE


def test_win_check_against_cell_scan():
    """tests/cpp/win_check_scan.cpp: the kernels' bit-packed win detection (c6_board.cuh, compiled for the host) against a
    cell-by-cell scan of search()'s rule on 200,000 random boards and on runs laid across row ends and board edges."""
    src = os.path.join(ROOT, "tests", "cpp", "win_check_scan.cpp")
    exe = os.path.join(BUILD_DIR, "win_check_scan")
    csrc = os.path.join(c6lib.PKG, "csrc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(f) for f in (src, os.path.join(csrc, "c6_board.cuh"))):
        subprocess.run(["g++", "-std=c++17", "-O2", "-I", csrc, "-I", "/usr/local/cuda/include", src, "-o", exe], check=True,
                       capture_output=True, text=True, timeout=600)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr


def test_host_policies_random_and_multi():
    """tests/cpp/host_policies.cpp: RandomPolicy under Agent::play (port of random_policy/tests.rs:7-25 over 300 seeds), MultiPolicy
    dispatch, Game::set error strings — the host-only part of the mirror, no GPU and no library needed."""
    src = os.path.join(ROOT, "tests", "cpp", "host_policies.cpp")
    exe = os.path.join(BUILD_DIR, "host_policies")
    csrc = os.path.join(c6lib.PKG, "csrc")
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not os.path.exists(exe) or any(os.path.getmtime(exe) < os.path.getmtime(f) for f in (src, os.path.join(csrc, "connect6_host.hpp"))):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", csrc, "-I", "/usr/local/cuda/include", src, "-o", exe], check=True,
                       capture_output=True, text=True, timeout=600)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference/Connect6/pybind"), reason="reference tree not mounted")
def test_reference_pybind_echo_test_runs_in_place():
    """PB/test_pybind/test_pybind.py::test_echo_pyeval, the reference's own file run where it lies against our pyconnect6
    (its other test, test_result_length, searches: the GPU suite runs the byte-compiled file)."""
    env = dict(os.environ, PYTHONPATH=os.path.join(c6lib.PKG, "python"))
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "/root/reference/Connect6/pybind/test_pybind/test_pybind.py",
                        "-k", "test_echo_pyeval"], capture_output=True, text=True, env=env, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "1 passed" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/Connect6"), reason="reference tree not mounted")
def test_extra_acceptance_binaries_build():
    assert all(os.path.exists(p) for p in build_reference_python())
    assert os.path.exists(build_device_overload()) and os.path.exists(build_host_mirror())
