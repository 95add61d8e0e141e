"""Test-side bindings: the shipped library (re-exported from the package's own ctypes binding) and the
CPU oracle (oracle/binding.py over oracle/_build/liboracle.so) — test infrastructure only, never used by the product path.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alphazero-connect6_b200")
sys.path.insert(0, os.path.join(PKG, "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from pyconnect6._capi import *  # noqa: E402,F401,F403
from pyconnect6._capi import product, make_config, alloc_path, alloc_result, alloc_int, alloc_float  # noqa: E402,F401
from binding import (ORACLE_DIR, ORACLE_PATH, ORACLE_TESTS, ORACLE_TIE, TIE_LOWEST, TIE_PHILOX, OracleDpOut, OracleGameOut,  # noqa: E402,F401
                     OracleParams, build_oracle, oracle, oracle_callback)


def build_product():
    subprocess.run(["make", "-s", "-C", PKG], check=True)
