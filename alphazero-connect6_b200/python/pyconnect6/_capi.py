"""ctypes binding of the shipped C-ABI library (include/connect6_b200.h).

`product()` loads alphazero-connect6_b200/lib/libcppconnect6.so and fails loudly when it has not been
built; nothing here falls back to a CPU implementation.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB_PATH = os.environ.get("C6_LIB", os.path.join(PKG, "lib", "libcppconnect6.so"))

CALLBACK = C.CFUNCTYPE(None, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int)
POLICY_CALLBACK = C.CFUNCTYPE(None, C.POINTER(C.c_float), C.POINTER(C.c_int))


class Path(C.Structure):
    _fields_ = [("turn", C.c_int), ("board", (C.c_int * 15) * 15), ("row", C.c_int), ("col", C.c_int)]


class PlayResult(C.Structure):
    _fields_ = [("winner", C.c_int), ("paths", C.POINTER(Path)), ("len", C.c_int)]


class Vec(C.Structure):
    _fields_ = [("vec", C.POINTER(PlayResult)), ("len", C.c_int)]


class VecInt(C.Structure):
    _fields_ = [("vec", C.POINTER(C.c_int)), ("len", C.c_int)]


class VecFloat(C.Structure):
    _fields_ = [("vec", C.POINTER(C.c_float)), ("len", C.c_int)]


# allocator callbacks (connect6.hpp:17-23): ctypes callbacks return the address as an integer
ALLOC_PATH = ALLOC_RESULT = ALLOC_INT = ALLOC_FLOAT = C.CFUNCTYPE(C.c_void_p, C.c_int)


class Config(C.Structure):
    _fields_ = [("num_simulation", C.c_int32), ("epsilon", C.c_float), ("dirichlet_alpha", C.c_double),
                ("c_puct", C.c_float), ("seed", C.c_uint32), ("device", C.c_int32), ("slots", C.c_int32),
                ("log2_records", C.c_int32), ("evaluator", C.c_int32), ("trace", C.c_int32), ("max_moves", C.c_int32), ("virtual_loss", C.c_int32), ("tie_break", C.c_int32)]


class Record(C.Structure):
    _fields_ = [("winner", C.c_int8), ("n_moves", C.c_uint8), ("finished", C.c_uint8), ("pad", C.c_uint8),
                ("game_id", C.c_uint32), ("moves", C.c_uint8 * 232)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("sims", "sum_ck", "sum_cleaf", "sum_depth", "evals", "gc_copied", "games_done")] + \
               [("peak_rec", C.c_uint32), ("peak_lnk", C.c_uint32), ("kernel_launches", C.c_uint64), ("bytes_per_game", C.c_uint64),
                ("degraded", C.c_uint64)]


class DpStats(C.Structure):  # c6_dp_stats
    _fields_ = [(n, C.c_uint64) for n in ("searches", "rollouts", "rollout_stones", "select_steps", "children_scanned",
                                          "replaced", "nodes", "errors")] + [("kernel_ms", C.c_float)]


EVAL_CALLBACK, EVAL_PHILOX, EVAL_DET, EVAL_NEG, EVAL_DEVICE_CALLBACK = 0, 1, 2, 3, 4
DEVICE_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)

# symbols of the reference ABI (connect6.hpp:44-91)
REFERENCE_SYMBOLS = ["cpp_play", "cpp_self_play", "cpp_play_with", "test_new_raw_path", "test_with_raw_path",
                     "test_echo_raw_path", "test_with_raw_play_result", "test_echo_raw_play_result", "test_with_raw_vec",
                     "test_echo_raw_vec", "test_echo_cppeval", "test_cpp_policy"]

_product = None


def product():
    """The shipped library.  Fails loudly when it has not been built."""
    global _product
    if _product is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libcppconnect6.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.cpp_self_play.restype = Vec
        L.cpp_self_play.argtypes = [CALLBACK, ALLOC_PATH, ALLOC_RESULT, C.c_int, C.c_float, C.c_double, C.c_float, C.c_bool, C.c_int]
        L.cpp_play.restype = Vec
        L.cpp_play.argtypes = [POLICY_CALLBACK, ALLOC_PATH, ALLOC_RESULT, C.c_bool, C.c_int]
        L.test_new_raw_path.restype = Path
        L.test_with_raw_path.restype = Path
        L.test_echo_raw_path.restype = Path
        L.test_echo_raw_path.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int]
        L.test_with_raw_play_result.restype = PlayResult
        L.test_with_raw_play_result.argtypes = [ALLOC_PATH]
        L.test_echo_raw_play_result.restype = PlayResult
        L.test_echo_raw_play_result.argtypes = [C.c_int, C.POINTER(Path), C.c_int, ALLOC_PATH]
        L.test_with_raw_vec.restype = VecInt
        L.test_with_raw_vec.argtypes = [ALLOC_INT]
        L.test_echo_raw_vec.restype = VecInt
        L.test_echo_raw_vec.argtypes = [C.POINTER(C.c_int), C.c_int, ALLOC_INT]
        L.test_echo_cppeval.restype = VecFloat
        L.test_echo_cppeval.argtypes = [C.c_int, C.POINTER(C.c_int), C.c_int, CALLBACK, ALLOC_FLOAT]
        L.test_cpp_policy.restype = VecInt
        L.test_cpp_policy.argtypes = [C.POINTER(C.c_float), POLICY_CALLBACK, ALLOC_INT]
        L.c6_last_error.restype = C.c_char_p
        L.c6_engine_create.restype = C.c_void_p
        L.c6_engine_create.argtypes = [C.POINTER(Config)]
        L.c6_engine_destroy.argtypes = [C.c_void_p]
        L.c6_engine_start.argtypes = [C.c_void_p, C.c_uint32, C.c_int64]
        L.c6_engine_step_device.argtypes = [C.c_void_p]
        L.c6_engine_wave_callback.argtypes = [C.c_void_p, CALLBACK]
        L.c6_engine_waves_callback.argtypes = [C.c_void_p, CALLBACK, C.c_int]
        L.c6_engine_external_move.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.c6_engine_waves_device_callback.argtypes = [C.c_void_p, DEVICE_CALLBACK, C.c_void_p, C.c_int]
        L.c6_engine_run_device_callback.argtypes = [C.c_void_p, DEVICE_CALLBACK, C.c_void_p]
        L.c6_engine_run_device.argtypes = [C.c_void_p]
        L.c6_release_cached_memory.argtypes = []
        L.c6_release_cached_memory.restype = None
        L.c6_engine_run_callback.argtypes = [C.c_void_p, CALLBACK]
        L.c6_engine_sync.argtypes = [C.c_void_p]
        L.c6_engine_cancel.argtypes = [C.c_void_p]
        L.c6_engine_cancel.restype = None
        L.c6_engine_timed_steps.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        L.c6_engine_timed_waves.argtypes = [C.c_void_p, CALLBACK, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.c6_engine_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        L.c6_engine_host_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.c6_engine_kernel_times.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.c6_engine_games_done.restype = C.c_int64
        L.c6_engine_games_done.argtypes = [C.c_void_p]
        L.c6_engine_records.argtypes = [C.c_void_p, C.POINTER(Record), C.c_int64]
        L.c6_engine_trace.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.c6_engine_trace_game.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.c6_engine_validate.restype = C.c_longlong
        L.c6_engine_validate.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
        L.c6_self_play_device.argtypes = [C.POINTER(Config), C.c_uint32, C.c_int64, C.POINTER(Record), C.POINTER(Stats)]
        L.c6_self_play_device_sharded.argtypes = [C.POINTER(Config), C.POINTER(C.c_int), C.c_int, C.c_uint32, C.c_int64, C.POINTER(Record), C.POINTER(Stats)]
        L.c6_win_check_host.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_int8), C.c_int64, C.c_int]
        L.c6_bench_win_check.argtypes = [C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_int8), C.c_int64]
        L.c6_bench_playouts.argtypes = [C.c_int64, C.c_int64, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int8),
                                        C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_uint64)]
        L.c6_default_policy_play.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int64, C.c_int, C.c_int, C.POINTER(Record),
                                             C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(DpStats)]
        FP = C.POINTER(C.c_float)
        L.c6_expand_records_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.c6_expand_records_host.argtypes = [C.POINTER(Record), C.c_int, C.c_int, FP, FP, FP, FP, C.POINTER(C.c_longlong)]
        L.c6_probe_gamma.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.c_int, C.c_int]
        L.c6_probe_math32.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int]
        L.c6_probe_tie.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.c_int, C.c_int]
        L.c6_probe_math.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int, C.c_int]
        L.c6_probe_pcie.argtypes = [C.c_int64, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3
        _product = L
    return _product


def make_config(**kw):
    cfg = Config()
    product().c6_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


# keep allocations made through the C allocator callbacks alive (the tests "leak" like the Rust tests do)
_keep = []


def _mk_alloc(ctype):
    def f(n):
        arr = (ctype * max(n, 1))()
        _keep.append(arr)
        return C.addressof(arr)
    return ALLOC_PATH(f)


alloc_path = _mk_alloc(Path)
alloc_result = _mk_alloc(PlayResult)
alloc_int = _mk_alloc(C.c_int)
alloc_float = _mk_alloc(C.c_float)


def unpack_boards(words):
    """n x 16 u32 bitboards -> n x 225 int8 cells (-1 black, +1 white)."""
    import numpy as np
    words = np.asarray(words, dtype=np.uint32).reshape(-1, 16)
    n = words.shape[0]
    out = np.zeros((n, 225), dtype=np.int8)
    for r in range(15):
        for c in range(15):
            bit = 16 * r + c
            w, b = bit >> 5, bit & 31
            out[:, r * 15 + c] -= ((words[:, w] >> b) & 1).astype(np.int8)
            out[:, r * 15 + c] += ((words[:, 8 + w] >> b) & 1).astype(np.int8)
    return out


def pack_boards(cells):
    import numpy as np
    cells = np.asarray(cells, dtype=np.int8).reshape(-1, 225)
    out = np.zeros((cells.shape[0], 16), dtype=np.uint32)
    for r in range(15):
        for c in range(15):
            bit = 16 * r + c
            w, b = bit >> 5, bit & 31
            out[:, w] |= ((cells[:, r * 15 + c] == -1).astype(np.uint32) << np.uint32(b))
            out[:, 8 + w] |= ((cells[:, r * 15 + c] == 1).astype(np.uint32) << np.uint32(b))
    return out
