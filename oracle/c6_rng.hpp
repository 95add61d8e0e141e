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

// ----------------------------------------------------------------------------------------------
// The sampler the search uses since round 2: log2 of a Gamma(shape, 1) draw, in f32.
//
// The reference normalises the Gamma draws by their sum in f64 and casts the result to f32
// (alphazero_policy/mod.rs:262-267).  With alpha = 0.03 almost every draw is astronomically small
// (X = P^(1/alpha)), so the engine keeps L = log2 X instead of X — no underflow for any alpha — and
// normalises like a softmax: noise_k = 2^(L_k - Lmax) / sum_j 2^(L_j - Lmax).  Same distribution
// (Ahrens–Dieter GS is still the exact rejection sampler; only the arithmetic is f32 where the
// reference's is f64 rounded to f32 at the end), one third of the instructions on the GPU.
//   * uniforms: see gamma_uniforms — trial 0 of four draws shares one Philox block (24 bits of U and the
//     top byte of V per draw; the rest of V from a block of its own, read only when that byte does not
//     decide); trial t >= 1 from words (0, 1) of block (a, b, k, t) as 32-bit uniforms; complements
//     1 - U, 1 - V from the complemented integers (exact before the conversion to f32);
//   * trial 0 is first tried on the hot branch with a sufficient acceptance test (2^L <= 2^floor(L)
//     (1 + frac L)); only if that fails the full rejection loop runs (from trial 0 again).
// lg2_det / ex2_det: table + polynomial, explicit fmaf only — bit-identical on x86-64 and the GPU.
// ----------------------------------------------------------------------------------------------
inline float u2f(uint32_t b) {
    float f;
    std::memcpy(&f, &b, 4);
    return f;
}
inline uint32_t f2u(float f) {
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return b;
}
inline float fmaf1(float a, float b, float c) { return __builtin_fmaf(a, b, c); }

// log2 of a positive normal f32.  x = 2^e m with m in [sqrt(1/2), sqrt(2)); the interval centre c nearest to m
// (33 centres, one of them exactly 1) gives r = m / c - 1 through the tabulated f32 reciprocal, |r| <= 2^-6;
// log2 x = e - log2(1/c) + log2(1 + r).
inline float lg2_det(float x) {
    const uint32_t bits = f2u(x);
    const int32_t e = (int32_t)(bits - 0x3F3504F3u) >> 23;
    const uint32_t mb = bits - ((uint32_t)e << 23);
    const float* row = LG2F_TABLE[(mb + 0x20000u - 0x3F340000u) >> 18];
    const float r = fmaf1(u2f(mb), row[0], -1.0f);
    float p = fmaf1(r, LG2F_POLY[3], LG2F_POLY[2]);
    p = fmaf1(r, p, LG2F_POLY[1]);
    p = fmaf1(r, p, LG2F_POLY[0]);
    return (float)e + fmaf1(r, p, row[1]);
}
// 2^t for t < 127; 0 below -125.  t = n + f, n = rint(t), 2^f by a degree-6 polynomial, scaled through the exponent.
inline float ex2_det(float t) {
    if (t < -125.0f) return 0.0f;
    const float n = __builtin_rintf(t);  // round-half-even (default rounding mode), as CUDA's rintf
    const float f = t - n;
    float p = EX2F_POLY[6];
    for (int i = 5; i >= 0; --i) p = fmaf1(p, f, EX2F_POLY[i]);
    return u2f(f2u(p) + ((uint32_t)(int32_t)n << 23));
}

struct GammaF {  // per-run constants of the shape < 1 sampler, rounded to f32 once on the host
    float b, inv_a, c1, am1;  // (e + a)/e, 1/a, b/a, a - 1
};
inline GammaF gamma_f(double shape) {
    const double b = (2.718281828459045 + shape) / 2.718281828459045;
    return GammaF{(float)b, (float)(1.0 / shape), (float)(b / shape), (float)(shape - 1.0)};
}
inline float unit_of(uint32_t w) { return fmaf1((float)w, 0x1p-32f, 0x1p-33f); }    // (w + 1/2) 2^-32, rounded once
inline float unit24_of(uint32_t w) { return fmaf1((float)w, 0x1p-24f, 0x1p-25f); }  // (w + 1/2) 2^-24 for a 24-bit w

// The uniforms of trial `trial` of draw k.  Trial 0 takes ONE word of Philox block (a, b, kb, 0), kb = (k >> 7) << 5 | (k & 31),
// word (k >> 5) & 3 — four draws per block: its top 24 bits are U (all the f32 conversion keeps of a uniform anyway), its low 8
// bits the top byte of V, which decides almost every acceptance; the other 24 bits of V come from word 0 of block
// (a, b, k, 0xFFFFFFFF) and are only looked at when the first byte does not decide.  Trial t >= 1 takes words 0, 1 of block
// (a, b, k, t) as 32-bit uniforms.
struct GammaUniforms {
    float U, Uc, V, Vc;  // U, 1 - U, V, 1 - V
};
inline uint32_t gamma_word0(uint32_t a, uint32_t b, uint32_t k, uint32_t k0, uint32_t k1) {
    return philox4x32_10(a, b, ((k >> 7) << 5) | (k & 31u), 0u, k0, k1).v[(k >> 5) & 3u];
}
inline GammaUniforms gamma_uniforms(uint32_t a, uint32_t b, uint32_t k, uint32_t trial, uint32_t k0, uint32_t k1) {
    if (trial == 0) {
        const uint32_t w = gamma_word0(a, b, k, k0, k1);
        const uint32_t ext = philox4x32_10(a, b, k, 0xFFFFFFFFu, k0, k1).v[0];
        const uint32_t v = ((w & 0xFFu) << 24) | (ext >> 8);
        return GammaUniforms{unit24_of(w >> 8), unit24_of(0xFFFFFFu - (w >> 8)), unit_of(v), unit_of(~v)};
    }
    const Philox4 r = philox4x32_10(a, b, k, trial, k0, k1);
    return GammaUniforms{unit_of(r.v[0]), unit_of(~r.v[0]), unit_of(r.v[1]), unit_of(~r.v[1])};
}

// hot branch of trial 0 on the draw's word alone; false = undecided (the full loop decides).  The test only has to be
// SUFFICIENT for the loop's trial-0 acceptance (then both return the same L): 1 - V >= (255 - v8)/256 whatever the other
// bits of V are, and the loop accepts when 1 - V >= d with d = 1 - e^-X <= X = 2^L; GAMMA_THR[v8] = log2((255 - v8)/256)
// - 2^-20 (rounded down; the margin covers the rounding of ex2_det and of d), GAMMA_THR[255] = -3e38 (never).
inline bool gamma_log2_fast(const GammaF& g, uint32_t w, float& L) {
    const float P = g.b * unit24_of(w >> 8);
    L = 0.0f;
    if (P > 1.0f) return false;
    L = lg2_det(P) * g.inv_a;
    return L <= GAMMA_THR[w & 0xFFu];
}
inline float gamma_log2_full(const GammaF& g, uint32_t a, uint32_t b, uint32_t k, uint32_t k0, uint32_t k1) {
    float L = 0.0f;
    for (uint32_t t = 0; t < 64; ++t) {
        const GammaUniforms q = gamma_uniforms(a, b, k, t, k0, k1);
        const float P = g.b * q.U;
        if (P <= 1.0f) {  // X = P^(1/a); accept iff V <= e^-X, i.e. 1 - V >= 1 - e^-X
            L = lg2_det(P) * g.inv_a;
            if (L < -40.0f) return L;  // 1 - e^-X < 2^-40 < the smallest 1 - V
            const float X = ex2_det(L);
            float d;
            if (X < 0x1p-6f)
                d = X * fmaf1(X, fmaf1(X, fmaf1(X, -1.0f / 24.0f, 1.0f / 6.0f), -0.5f), 1.0f);  // X - X^2/2 + X^3/6 - X^4/24
            else
                d = 1.0f - ex2_det(X * -1.44269504f);
            if (q.Vc >= d) return L;
        } else {  // X = -ln((b - P)/a) = -ln((b/a)(1 - U)); accept iff V <= X^(a-1)
            const float X = lg2_det(g.c1 * q.Uc) * -0.693147182f;
            L = lg2_det(X > 0x1p-100f ? X : 0x1p-100f);
            if (lg2_det(q.V) <= g.am1 * L) return L;
        }
    }
    return L;
}
// log2 of draw (a, b, k) of the stream keyed (k0, k1)
inline float gamma_log2_sample(double shape, uint32_t a, uint32_t b, uint32_t k, uint32_t k0, uint32_t k1) {
    if (shape >= 1.0) return (float)(det_log(gamma_large_shape(shape, a, b, k, k0, k1)) * 1.4426950408889634);
    const GammaF g = gamma_f(shape);
    float L;
    if (gamma_log2_fast(g, gamma_word0(a, b, k, k0, k1), L)) return L;
    return gamma_log2_full(g, a, b, k, k0, k1);
}

// Exact ties of the PUCT maximum are broken uniformly (alphazero_policy/mod.rs:291-295): tie number
// floor(u n / 2^32) of the n tied children in row-major order, u = word 0 of Philox block (sim, depth, 0, 0)
// keyed (seed ^ TIE, game).
constexpr uint32_t DOMAIN_TIE = 0x54494542u;  // "TIEB"
inline uint32_t tie_pick(uint32_t sim, uint32_t depth, uint32_t seed, uint32_t game, uint32_t n) {
    const Philox4 r = philox4x32_10(sim, depth, 0u, 0u, seed ^ DOMAIN_TIE, game);
    return (uint32_t)(((uint64_t)r.v[0] * (uint64_t)n) >> 32);
}

}  // namespace c6o
