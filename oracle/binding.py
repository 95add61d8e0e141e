"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so (the CPU restatement of the reference's
path, oracle/c6_oracle.hpp).  Imported by tests/ (through tests/c6lib.py), __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs — never by the product package."""
import ctypes as C
import os
import subprocess

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
ORACLE_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")
ORACLE_TESTS = os.path.join(ORACLE_DIR, "_build", "oracle_tests")

TIE_LOWEST, TIE_PHILOX = 0, 1          # c6_config.tie_break
ORACLE_TIE = {0: 0, 1: 2}               # -> OracleParams.tie_mode (1 is the oracle's non-reproducible splitmix pick)


class OracleParams(C.Structure):
    _fields_ = [("num_simulation", C.c_int32), ("epsilon", C.c_float), ("dirichlet_alpha", C.c_double),
                ("c_puct", C.c_float), ("seed", C.c_uint32), ("noise_mode", C.c_int32), ("tie_mode", C.c_int32),
                ("faithful", C.c_int32), ("evaluator", C.c_int32), ("max_moves", C.c_int32), ("virtual_loss", C.c_int32)]


class OracleGameOut(C.Structure):
    _fields_ = [("winner", C.c_int32), ("n_moves", C.c_int32), ("finished", C.c_int32), ("moves", C.c_int32 * 225),
                ("turns", C.c_int32 * 225)] + [(n, C.c_uint64) for n in ("sims", "sum_ck", "sum_cleaf", "sum_depth", "evals", "max_nodes")]


class OracleDpOut(C.Structure):
    _fields_ = [("winner", C.c_int), ("n_moves", C.c_int), ("moves", C.c_int * 225), ("root_visit", C.c_int32 * 225),
                ("root_black_win", C.c_int32 * 225)] + \
               [(n, C.c_uint64) for n in ("searches", "rollouts", "rollout_stones", "select_steps", "children_scanned", "replaced", "nodes")]


_oracle = None


def build_oracle():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_PATH):
            build_oracle()
        L = C.CDLL(ORACLE_PATH)
        L.oracle_search.argtypes = [C.POINTER(C.c_int8)]
        L.oracle_augment_way8.argtypes = [C.POINTER(C.c_int8), C.POINTER(C.c_int8)]
        L.oracle_recover_way8.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_transform.argtypes = [C.c_int, C.POINTER(C.c_int8), C.POINTER(C.c_int8)]
        L.oracle_game_replay.argtypes = [C.POINTER(C.c_int), C.c_int] + [C.POINTER(C.c_int)] * 5
        L.oracle_zobrist.restype = C.c_uint64
        L.oracle_zobrist.argtypes = [C.c_int, C.c_int]
        L.oracle_philox.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.oracle_det_log.restype = C.c_double
        L.oracle_det_log.argtypes = [C.c_double]
        L.oracle_det_exp.restype = C.c_double
        L.oracle_det_exp.argtypes = [C.c_double]
        L.oracle_gamma.restype = C.c_double
        L.oracle_gamma.argtypes = [C.c_uint32] * 5 + [C.c_double]
        L.oracle_gamma_log2.restype = C.c_float
        L.oracle_gamma_log2.argtypes = [C.c_uint32] * 5 + [C.c_double]
        L.oracle_gamma_log2_many.argtypes = [C.c_uint32] * 3 + [C.c_int] + [C.c_uint32] * 2 + [C.c_double, C.POINTER(C.c_float)]
        L.oracle_math32_many.argtypes = [C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.oracle_tie_pick.restype = C.c_uint32
        L.oracle_tie_pick.argtypes = [C.c_uint32] * 5
        L.oracle_playout.argtypes = [C.c_uint32, C.c_uint64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int8)]
        L.oracle_script_game.argtypes = [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p]
        L.oracle_self_play.argtypes = [C.c_void_p, C.POINTER(OracleParams), C.c_uint32, C.POINTER(OracleGameOut),
                                       C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.oracle_self_play_from.argtypes = [C.c_void_p, C.POINTER(OracleParams), C.c_uint32, C.POINTER(C.c_int), C.c_int, C.POINTER(OracleGameOut),
                                            C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float)]
        L.oracle_self_play_many.restype = C.c_double
        L.oracle_self_play_many.argtypes = [C.c_void_p, C.POINTER(OracleParams), C.c_uint32, C.c_int, C.c_int, C.POINTER(OracleGameOut)]
        L.oracle_play_with.argtypes = [C.c_void_p, C.POINTER(OracleParams), C.c_uint32, C.POINTER(C.c_int), C.c_int, C.POINTER(OracleGameOut)]
        L.oracle_default_policy_play.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(OracleDpOut)]
        L.oracle_default_policy_many.restype = C.c_double
        L.oracle_default_policy_many.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int, C.POINTER(OracleDpOut)]
        L.oracle_search_batch.argtypes = [C.POINTER(C.c_int8), C.c_int64, C.POINTER(C.c_int8), C.c_int]
        L.oracle_search_batch_packed.argtypes = [C.POINTER(C.c_uint32), C.c_int64, C.POINTER(C.c_int8), C.c_int]
        L.oracle_playout_batch.argtypes = [C.c_uint32, C.c_uint64, C.c_int64, C.POINTER(C.c_int8), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.c_int]
        L.oracle_random_evaluator.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int8), C.POINTER(C.c_float), C.POINTER(C.c_float)]
        _oracle = L
    return _oracle


def oracle_callback(name):
    """Address of one of the oracle library's deterministic evaluators, usable as the reference Callback."""
    fn = getattr(oracle(), name)
    return C.cast(fn, C.c_void_p).value
