#!/usr/bin/env python
"""One bounded launch of k_default_policy for ncu: 65,536 games (a B200 keeps 75,776 resident at 128 registers per thread),
50 searches per stone, the first 10 stones of every game."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "alphazero-connect6_b200", "python"))
from pyconnect6 import _capi as capi  # noqa: E402

L = capi.product()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rec = (capi.Record * n)()
st = capi.DpStats()
assert L.c6_default_policy_play(50, 0xC6, 0, n, 10, 0, rec, None, None, C.byref(st)) == 0, L.c6_last_error()
print("k_default_policy:", n, "games x 10 stones,", st.searches, "searches,", round(st.kernel_ms, 1), "ms")
