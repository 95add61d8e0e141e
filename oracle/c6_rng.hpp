// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by, or
// executed from the shipped library; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may use it.
//
// Deterministic random sources for the places where the reference uses rand 0.6's
// `thread_rng()` (unseedable, unpinned by any reference test — SURVEY.md §8c "parity unpinned"):
//   * Dirichlet noise        — alphazero_policy/mod.rs:262,283
//   * RandomEvaluator        — alphazero_policy/mod.rs:137-169
// The published algorithms restated here:
//   * Philox4x32-10 (Salmon et al., SC'11 "Parallel random numbers: as easy as 1, 2, 3").
//   * rand 0.6 `Dirichlet::new_with_param(alpha, n)` = n independent Gamma(alpha, 1) draws
//     normalised by their sum, in f64 (rand 0.6.x is a caret dependency, no lockfile:
//     LC/Cargo.toml:12).  rand's own Gamma uses Marsaglia–Tsang (+ U^(1/shape) for shape < 1) on
//     ziggurat normals from an unseedable thread_rng, so no stream can be matched; what is
//     contractual is the distribution.  The exact samplers used here are the ones the engine
//     runs: Ahrens–Dieter GS (1974) for shape < 1, Marsaglia–Tsang (2000) on polar-method
//     normals for shape >= 1 — only + - * / sqrt fma and the fixed-polynomial log/exp below,
//     hence bit-identical on x86-64 and on the GPU (no implicit FMA contraction on either side).
// Everything in this header must stay bit-for-bit equal to the device twin in
// alphazero-connect6_b200/csrc/c6_rng.cuh; tests/test_rng_parity.py checks that on the GPU.
#pragma once
#include <cstdint>
#include <cstring>

#include "c6_tables.hpp"

namespace c6o {

// ----------------------------------------------------------------------------------------------
// Philox4x32-10
// ----------------------------------------------------------------------------------------------
struct Philox4 {
    uint32_t v[4];
};

inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                             uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0, c1 = n1, c2 = n2, c3 = n3;
        k0 += W0, k1 += W1;
    }
    return Philox4{{c0, c1, c2, c3}};
}

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Key domains (key0 = seed ^ domain, key1 = game id).
constexpr uint32_t DOMAIN_NOISE = 0x4E4F4953u;  // "NOIS"
constexpr uint32_t DOMAIN_EVAL = 0x4556414Cu;   // "EVAL"
constexpr uint32_t DOMAIN_PLAYOUT = 0x504C4159u;  // "PLAY"
constexpr uint32_t DOMAIN_BOARDS = 0x42524453u;   // "BRDS"

// u32 -> f32 uniform on [-1, 1): 24-bit integer, exact in f32 (SURVEY B.4 style).
inline float u32_to_pm1(uint32_t x) { return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f; }
// two u32 -> f64 uniform on (0, 1): the top 52 bits become the mantissa of a double in [1, 2);
// subtracting 1 and adding 2^-53 is exact, so the result is (2k + 1) / 2^53, never 0 or 1.
inline double bits_to_double(uint64_t b);
inline double u64_to_open01(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    double one_two = bits_to_double(0x3FF0000000000000ull | (x >> 12));
    return (one_two - 1.0) + 1.1102230246251565e-16;
}

// ----------------------------------------------------------------------------------------------
// Table + fixed-polynomial log / exp in f64 (only IEEE + - * fma and bit moves: same bits on CPU and GPU).
// ----------------------------------------------------------------------------------------------
inline double bits_to_double(uint64_t b) {
    double d;
    std::memcpy(&d, &b, 8);
    return d;
}
inline uint64_t double_to_bits(double d) {
    uint64_t b;
    std::memcpy(&b, &d, 8);
    return b;
}

inline double fma1(double a, double b, double c) { return __builtin_fma(a, b, c); }  // one rounding (FMA3 / libm)

// Horner evaluation with one fused multiply-add per coefficient, highest degree first.
template <int N>
inline double horner_fma(const double (&coef)[N], double t) {
    double acc = coef[0];
    for (int i = 1; i < N; ++i) acc = fma1(acc, t, coef[i]);
    return acc;
}

// ln x for positive finite x, division-free (tables: c6_tables.hpp).  x = 2^e * m with m in [1, 2);
// the top six mantissa bits pick the interval centre c; r = m / c - 1 through the tabulated 1/c;
// ln x = e ln2 + ln c + log1p(r) with log1p(r) = r + r^2 (-1/2 + r/3 - ... - r^6/8).
inline double det_log(double x) {
    int e = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) {  // subnormal: scale by 2^54
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)((b >> 52) & 0x7FF) - 1023;
    double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);  // [1,2)
    const double* row = LOG_TABLE[(b >> 46) & 63];
    double r = fma1(m, row[0], -1.0);
    static const double C[7] = {-1.0 / 8.0, 1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, -1.0 / 4.0, 1.0 / 3.0, -0.5};
    double log1p_r = fma1(horner_fma(C, r), r * r, r);
    double ed = (double)e;
    return fma1(ed, 0.6931471803691238, row[1] + fma1(ed, 1.9082149292705877e-10, log1p_r));
}

// e^x.  k = floor(x * 32/ln2 + 1/2) = 32 q + j; r = x - k ln2/32 (two-part ln2); e^x = 2^q * 2^(j/32) *
// (1 + expm1(r)) with expm1 to r^6; the 2^q scaling is done in two exact steps so that subnormal
// results round once.
inline double det_exp(double x) {
    if (x < -745.2) return 0.0;
    if (x > 709.0) x = 709.0;
    double kd = (double)(int64_t)__builtin_floor(fma1(x, 46.16624130844683, 0.5));
    double r = fma1(-kd, 5.9631716539705866e-12, fma1(-kd, 0.02166084938653512, x));
    static const double C[6] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};
    double expm1_r = horner_fma(C, r) * r;
    int k = (int)kd;
    double tj = EXP_TABLE[k & 31];
    double y = fma1(tj, expm1_r, tj);
    int q = k >> 5;  // arithmetic shift: floor division by 32
    int q1 = q / 2, q2 = q - q1;
    double s1 = bits_to_double((uint64_t)(q1 + 1023) << 52);
    double s2 = bits_to_double((uint64_t)(q2 + 1023) << 52);
    return (y * s1) * s2;
}

}  // namespace c6o

#include <cmath>
namespace c6o {
inline double det_sqrt(double x) { return std::sqrt(x); }  // IEEE sqrt, correctly rounded

// ----------------------------------------------------------------------------------------------
// Gamma(shape, 1), draw (a, b, c) of the stream keyed (k0, k1).  Trial t of the rejection loop
// reads Philox block(s) of counter (a, b, c, t) — two 53-bit uniforms per block.
//   shape <  1 : Ahrens & Dieter (1974) algorithm GS.  With b = (e + shape)/e and P = b*U:
//                  P <= 1 : X = P^(1/shape), accept iff V <= e^(-X)   (cubic lower bound tried first)
//                  P >  1 : X = -ln((b - P)/shape), accept iff V <= X^(shape-1)
//   shape >= 1 : Marsaglia & Tsang (2000), normal deviate from the Marsaglia polar method
//                (blocks 2t and 2t+1 of trial t).
// ----------------------------------------------------------------------------------------------
struct Uniform2 {
    double first, second;
};
inline Uniform2 uniforms(uint32_t a, uint32_t b, uint32_t c, uint32_t block, uint32_t k0, uint32_t k1) {
    Philox4 r = philox4x32_10(a, b, c, block, k0, k1);
    return Uniform2{u64_to_open01(r.v[0], r.v[1]), u64_to_open01(r.v[2], r.v[3])};
}

inline double gamma_small_shape(double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1) {
    const double inv_shape = 1.0 / shape;
    const double bb = (2.718281828459045 + shape) / 2.718281828459045;
    double x = 0.0;
    for (uint32_t t = 0; t < 64; ++t) {
        Uniform2 uv = uniforms(a, b, c, t, k0, k1);
        double p = bb * uv.first;
        if (p <= 1.0) {
            x = det_exp(det_log(p) * inv_shape);
            // squeeze first (e^-x >= 1 - x + x^2/2 - x^3/6 for x >= 0), then the exact test
            double lower = fma1(x, fma1(x, fma1(x, -1.0 / 6.0, 0.5), -1.0), 1.0);
            if (uv.second <= lower || uv.second <= det_exp(-x)) return x;
        } else {
            x = -det_log((bb - p) * inv_shape);
            if (uv.second <= det_exp((shape - 1.0) * det_log(x))) return x;
        }
    }
    return x;  // 64 rejections in a row: probability < 1e-100
}

inline double gamma_large_shape(double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1) {
    const double d = shape - 1.0 / 3.0;
    const double cc = 1.0 / det_sqrt(9.0 * d);
    double out = d;
    for (uint32_t t = 0; t < 64; ++t) {
        Uniform2 pq = uniforms(a, b, c, 2 * t, k0, k1);
        double p = 2.0 * pq.first - 1.0, q = 2.0 * pq.second - 1.0;
        double s = p * p + q * q;
        if (s >= 1.0 || s == 0.0) continue;
        double x = p * det_sqrt(-2.0 * det_log(s) / s);
        double vc = 1.0 + cc * x;
        if (vc <= 0.0) continue;
        double v = vc * vc * vc;
        double u = uniforms(a, b, c, 2 * t + 1, k0, k1).first;
        double x2 = x * x;
        out = d * v;
        if (u < 1.0 - 0.0331 * x2 * x2) return out;
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) return out;
    }
    return out;
}

inline double gamma_sample(double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1) {
    return shape < 1.0 ? gamma_small_shape(shape, a, b, c, k0, k1) : gamma_large_shape(shape, a, b, c, k0, k1);
}

}  // namespace c6o
