#!/usr/bin/env python
"""Generates the constant tables of the deterministic f64 log / exp used by the Dirichlet-noise sampler.

  python scripts/gen_tables.py

Writes the SAME literals (C99 hex floats: exact) into
  alphazero-connect6_b200/csrc/c6_tables.cuh   (engine)
  oracle/c6_tables.hpp                         (CPU oracle, test infrastructure)
so that both sides evaluate bit-identical arithmetic.  Values:
  LOG_TAB[i] = {1/c_i, ln c_i}, c_i = 1 + (2i+1)/128, i = 0..63   (mantissa interval centres)
  EXP_TAB[j] = 2^(j/32), j = 0..31
and of the f32 log2 / exp2 of the default (alpha < 1) log-space Gamma sampler:
  LG2F_TAB[i] = {r_i, -log2 r_i}, r_i = f32(1/c_i), c_i = the f32 with bits 0x3F340000 + i * 2^18, i = 0..32
                (mantissa interval centres over [sqrt(1/2), sqrt(2)); c_19 = 1 exactly, so r_19 = 1, log = 0)
  LG2F_POLY     log2(1 + r) = r (c1 + r (c2 + r (c3 + r c4))), c_n = (-1)^(n+1) / (n ln 2)
  EX2F_POLY     2^f on [-1/2, 1/2], degree 6, constant term exactly 1 (weighted least squares on Chebyshev nodes)
"""
import math
import os
from fractions import Fraction

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ln_exact(fr, terms=80):
    """ln of a Fraction in (0.5, 2.5) to far beyond double precision: 2 atanh((x-1)/(x+1))."""
    y = (fr - 1) / (fr + 1)
    y2 = y * y
    s, t = Fraction(0), y
    for k in range(terms):
        s += t / (2 * k + 1)
        t *= y2
    return 2 * s


def to_double(fr):
    return fr.numerator / fr.denominator  # correctly rounded by Python's int/int true division


def pow2_frac(j, n=32):
    """2^(j/n) correctly rounded: integer n-th root of 2^j * 2^(n*200), then rounded once to double."""
    num = 1 << (j + n * 200)
    lo, hi = 0, 1 << 260
    while lo < hi:  # floor(num ** (1/n))
        mid = (lo + hi + 1) // 2
        if mid ** n <= num:
            lo = mid
        else:
            hi = mid - 1
    return to_double(Fraction(2 * lo + 1, 2 << 200))  # midpoint of [lo, lo+1): error << half an ulp


log_rows = []
for i in range(64):
    c = Fraction(128 + 2 * i + 1, 128)
    log_rows.append((to_double(1 / c), to_double(ln_exact(c))))
exp_rows = [pow2_frac(j) for j in range(32)]
for i, (a, b) in enumerate(log_rows):
    c = 1 + (2 * i + 1) / 128
    assert abs(a * c - 1) < 1e-15 and abs(b - math.log(c)) < 1e-16
for j, v in enumerate(exp_rows):
    assert abs(v - 2 ** (j / 32)) <= 2.3e-16 * v

# ---- f32 tables -------------------------------------------------------------------------------------
import struct

import numpy as np


def f32_of_fraction(fr):
    """Fraction -> correctly rounded f32 (as a Python float holding an f32 value)."""
    if fr == 0:
        return 0.0
    d = to_double(fr)  # 53-bit rounding first; the f32 neighbours are then compared exactly
    f = struct.unpack("<f", struct.pack("<f", d))[0]
    bits = struct.unpack("<I", struct.pack("<f", f))[0]
    best = None
    for b in (bits - 1, bits, bits + 1):
        v = struct.unpack("<f", struct.pack("<I", b))[0]
        err = abs(Fraction(v) - fr)
        if best is None or err < best[0]:
            best = (err, v)
    return best[1]


LN2 = ln_exact(Fraction(2))
lg2f_rows = []
for i in range(33):
    c = Fraction(struct.unpack("<f", struct.pack("<I", 0x3F340000 + i * 0x40000))[0])
    r = f32_of_fraction(1 / c)
    lg = f32_of_fraction(-ln_exact(Fraction(r)) / LN2)
    lg2f_rows.append((r, lg))
assert lg2f_rows[19] == (1.0, 0.0)
lg2f_poly = [f32_of_fraction(Fraction((-1) ** (n + 1), n) / LN2) for n in (1, 2, 3, 4)]

_n = 4000
_x = np.cos(np.pi * (np.arange(_n) + 0.5) / _n) * 0.5
_y = 2.0 ** _x
_w = np.abs(_x) / _y  # relative error of 1 + x g(x) against 2^x
_coef, *_ = np.linalg.lstsq(np.vander(_x, 6, increasing=True) * _w[:, None], ((_y - 1) / _x) * _w, rcond=None)
ex2f_poly = [1.0] + [float(np.float32(v)) for v in _coef]


# fast acceptance thresholds of the log2-space Gamma sampler: with the top byte v of V known, 1 - V > (255 - v)/256 whatever the
# other 24 bits are, so a draw X = 2^L with 2^L <= (255 - v)/256 is accepted for sure.  T[v] = log2((255 - v)/256) - 2^-20 rounded
# DOWN to f32 (the margin covers the 1.2e-7 relative error of ex2_det in the exact test the full loop applies); T[255] = -inf.
gthr = []
for v in range(256):
    if v == 255:
        gthr.append(-float(2.0 ** 127))  # "never": below every log2 the sampler can produce
        continue
    t = math.log2((255 - v) / 256) - 2.0 ** -20  # double precision is ample next to the margin
    f = float(np.float32(t))
    if f > t:
        f = float(np.nextafter(np.float32(f), np.float32(-np.inf)))
    gthr.append(f)
assert all(2.0 ** gthr[v] < (255 - v) / 256 for v in range(255)) and gthr[0] < 0


def fhex(v):
    return float(v).hex() + "f"


gthr_init = ",\n".join("    " + ", ".join(fhex(v) for v in gthr[i:i + 4]) for i in range(0, 256, 4))
lg2f_init = ",\n".join(f"    {{{fhex(a)}, {fhex(b)}}}" for a, b in lg2f_rows)
lg2f_poly_init = ", ".join(fhex(v) for v in lg2f_poly)
ex2f_poly_init = ", ".join(fhex(v) for v in ex2f_poly)

BODY = "// GENERATED by scripts/gen_tables.py — do not edit.  C99 hex floats (exact).\n"
log_init = ",\n".join(f"    {{{a.hex()}, {b.hex()}}}" for a, b in log_rows)
exp_init = ",\n".join(f"    {v.hex()}" for v in exp_rows)

with open(os.path.join(ROOT, "alphazero-connect6_b200", "csrc", "c6_tables.cuh"), "w") as f:
    f.write(BODY + "#pragma once\nnamespace c6 {\n"
            "// {1/c_i, ln c_i}, c_i = 1 + (2i+1)/128\n"
            f"static __device__ __constant__ double2 LOG_TAB_C[64] = {{\n{log_init}}};\n"
            f"static const double LOG_TAB_H[64][2] = {{\n{log_init}}};\n"
            "// 2^(j/32)\n"
            f"static __device__ __constant__ double EXP_TAB_C[32] = {{\n{exp_init}}};\n"
            f"static const double EXP_TAB_H[32] = {{\n{exp_init}}};\n"
            "// f32: {r_i, -log2 r_i}, r_i = f32(1/c_i), c_i = bits 0x3F340000 + i * 2^18\n"
            f"static __device__ __constant__ float2 LG2F_TAB_C[33] = {{\n{lg2f_init}}};\n"
            f"static __device__ __constant__ float LG2F_POLY_C[4] = {{{lg2f_poly_init}}};\n"
            f"static __device__ __constant__ float EX2F_POLY_C[7] = {{{ex2f_poly_init}}};\n"
            "// fast acceptance thresholds of the Gamma sampler by the top byte of V: log2((255 - v)/256) - 2^-20, rounded down\n"
            f"static __device__ __constant__ float GAMMA_THR_C[256] = {{\n{gthr_init}}};\n"
            "}  // namespace c6\n")
with open(os.path.join(ROOT, "oracle", "c6_tables.hpp"), "w") as f:
    f.write("// ORACLE — TEST INFRASTRUCTURE ONLY.\n" + BODY + "#pragma once\nnamespace c6o {\n"
            "// {1/c_i, ln c_i}, c_i = 1 + (2i+1)/128\n"
            f"static const double LOG_TABLE[64][2] = {{\n{log_init}}};\n"
            "// 2^(j/32)\n"
            f"static const double EXP_TABLE[32] = {{\n{exp_init}}};\n"
            "// f32: {r_i, -log2 r_i}, r_i = f32(1/c_i), c_i = bits 0x3F340000 + i * 2^18\n"
            f"static const float LG2F_TABLE[33][2] = {{\n{lg2f_init}}};\n"
            f"static const float LG2F_POLY[4] = {{{lg2f_poly_init}}};\n"
            f"static const float EX2F_POLY[7] = {{{ex2f_poly_init}}};\n"
            "// fast acceptance thresholds of the Gamma sampler by the top byte of V: log2((255 - v)/256) - 2^-20, rounded down\n"
            f"static const float GAMMA_THR[256] = {{\n{gthr_init}}};\n"
            "}  // namespace c6o\n")
print("tables written")
