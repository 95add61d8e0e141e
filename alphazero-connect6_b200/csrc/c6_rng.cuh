// Counter-based random sources and fixed-polynomial f64 math for the engine (device side).
//
// Replaces the reference's `thread_rng()` uses on the hot path:
//   Dirichlet noise at every select step   LC/policy/alphazero_policy/mod.rs:262,283
//   RandomEvaluator                        LC/policy/alphazero_policy/mod.rs:137-169
// with Philox4x32-10 streams keyed by (seed, game) and counted by (simulation, depth, child), so a
// game's search is a pure function of its id.  Dirichlet(alpha; C) = C independent Gamma(alpha, 1)
// draws normalised by their sum, as rand 0.6 builds it.  The Gamma draws use exact rejection samplers
// chosen for the GPU: Ahrens–Dieter GS for alpha < 1 (the default 0.03), kept in log2 space and in f32
// since round 2 (the r01 captures had the f64 sampler at half of all instructions: Philox 18 %, f64
// log / exp 22 %, diverged slow path 12 %; a draw is now one half of a Philox block, a 14-instruction
// log2 and an integer-built bound), Marsaglia–Tsang on polar normals in f64 for alpha >= 1.  log2 /
// exp2 (and the f64 log / exp) are tables + fixed polynomials over IEEE + - * fma only, and this
// translation unit is compiled with -fmad=false, so every value is reproducible bit for bit on a CPU
// (tests/ compare against the oracle's independent restatement).  rand's own stream is unseedable
// (thread_rng) and unpinned by any reference test, so only the distribution is contractual
// (tests/test_oracle.py: Kolmogorov–Smirnov against the Gamma CDF, moments).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "c6_tables.cuh"

namespace c6 {

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;  // one IMAD.WIDE each
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c0 = n0, c1 = (uint32_t)p1, c2 = n2, c3 = (uint32_t)p0;
        k0 += W0, k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

constexpr uint32_t DOMAIN_NOISE = 0x4E4F4953u;
constexpr uint32_t DOMAIN_EVAL = 0x4556414Cu;
constexpr uint32_t DOMAIN_PLAYOUT = 0x504C4159u;
constexpr uint32_t DOMAIN_BOARDS = 0x42524453u;

__host__ __device__ __forceinline__ float u32_to_pm1(uint32_t x) {
    return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f;
}
#ifdef __CUDA_ARCH__
__device__ __forceinline__ double bits_to_double(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t double_to_bits(double d) { return (uint64_t)__double_as_longlong(d); }
#else
inline double bits_to_double(uint64_t b) {
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
}
inline uint64_t double_to_bits(double d) {
    uint64_t b;
    __builtin_memcpy(&b, &d, 8);
    return b;
}
#endif

// 52 random bits -> uniform on (0,1): mantissa of a double in [1,2), minus 1, plus half a step, all exact
// (no 64-bit integer -> double conversion, which is slow on the GPU).
__host__ __device__ __forceinline__ double u64_to_open01(uint32_t lo, uint32_t hi) {
    const uint64_t x = ((uint64_t)hi << 32) | lo;
    return (bits_to_double(0x3FF0000000000000ull | (x >> 12)) - 1.0) + 1.1102230246251565e-16;
}

#ifdef __CUDACC__
// IEEE fused multiply-add, explicit (the translation unit is compiled with -fmad=false, so nothing else
// contracts): one rounding, identical on sm_100a and x86-64 FMA3 (the oracle uses __builtin_fma / __builtin_fmaf).
__device__ __forceinline__ double det_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float det_fmaf(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// Table of the f32 log2 below, staged in shared memory by every kernel that samples noise (per-lane indices
// would serialise in the constant cache).  The f64 log / exp of the alpha >= 1 sampler (not the default, out of
// line) read their tables from the constant bank.
#ifndef C6_SQRT_TAB
#define C6_SQRT_TAB 0
#endif
#ifndef C6_GAMMA_THR
#define C6_GAMMA_THR 1  // 1: acceptance threshold of the hot branch from a 256-entry table; 0: computed (2^floor(L) (1 + frac L))
#endif
struct MathTab {
    float2 lg2t[33];  // {r_i, -log2 r_i}, r_i = f32(1/c_i), c_i = bits 0x3F340000 + i 2^18
#if C6_GAMMA_THR
    float thr[256];  // log2((255 - v)/256) - 2^-20 rounded down, v = top byte of V (v = 255: never)
#endif
#if C6_SQRT_TAB
    float sqrt_tab[1024];  // sqrtf(i) (IEEE, correctly rounded: the value the PUCT pass would compute)
#endif
};
__device__ __forceinline__ void load_math_tab(MathTab& T) {
    if (threadIdx.x < 33) T.lg2t[threadIdx.x] = LG2F_TAB_C[threadIdx.x];  // blocks have at least 64 threads
#if C6_GAMMA_THR
    for (int i = threadIdx.x; i < 256; i += blockDim.x) T.thr[i] = GAMMA_THR_C[i];
#endif
#if C6_SQRT_TAB
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) T.sqrt_tab[i] = sqrtf((float)i);
#endif
}

// ------------------------------------------------------------------------------------------------
// f32 log2 / exp2: table + polynomial over explicit fmaf, bit-identical to oracle/c6_rng.hpp.
// ------------------------------------------------------------------------------------------------
// log2 of a positive normal f32.  x = 2^e m, m in [sqrt(1/2), sqrt(2)); the interval centre c nearest to m (33
// centres, one of them exactly 1) gives r = m / c - 1 through the tabulated reciprocal, |r| <= 2^-6;
// log2 x = e - log2(1/c) + log2(1 + r), log2(1 + r) to r^4.  Absolute error ~2^-25 + 2^-24 |log2 x|.
__device__ __forceinline__ float lg2_det(const MathTab& T, float x) {
    const uint32_t bits = __float_as_uint(x);
    const int e = (int)(bits - 0x3F3504F3u) >> 23;
    const uint32_t mb = bits - ((uint32_t)e << 23);
    const float2 t = T.lg2t[(mb + 0x20000u - 0x3F340000u) >> 18];
    const float r = det_fmaf(__uint_as_float(mb), t.x, -1.0f);
    float p = det_fmaf(r, LG2F_POLY_C[3], LG2F_POLY_C[2]);
    p = det_fmaf(r, p, LG2F_POLY_C[1]);
    p = det_fmaf(r, p, LG2F_POLY_C[0]);
    return (float)e + det_fmaf(r, p, t.y);
}
// 2^t for t < 127; 0 below -125.  t = n + f, n = rint(t), 2^f by a degree-6 polynomial (relative error ~1.3 ulp),
// scaled through the exponent field.
__device__ __forceinline__ float ex2_det(float t) {
    const float n = rintf(t);
    const float f = t - n;
    float p = EX2F_POLY_C[6];
#pragma unroll
    for (int i = 5; i >= 0; --i) p = det_fmaf(p, f, EX2F_POLY_C[i]);
    const float y = __uint_as_float(__float_as_uint(p) + ((uint32_t)(int)n << 23));
    return t < -125.0f ? 0.0f : y;
}

// natural log of a positive finite double, division-free: x = 2^e m, m in [1,2) falls in one of 64
// intervals with centre c_i; r = m/c_i - 1 (|r| < 2^-7) by one fma with the tabulated reciprocal;
// ln x = e ln2 + ln c_i + log1p(r), log1p as a degree-8 alternating series.  Absolute error ~1e-16.
// (alpha >= 1 sampler and the parity probes only.)
static __device__ __noinline__ double det_log(double x) {
    int e = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) {
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)((b >> 52) & 0x7FF) - 1023;
    const double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    const double2 t = LOG_TAB_C[(uint32_t)(b >> 46) & 63u];
    const double r = det_fma(m, t.x, -1.0);
    double p = -1.0 / 8.0;
    p = det_fma(p, r, 1.0 / 7.0);
    p = det_fma(p, r, -1.0 / 6.0);
    p = det_fma(p, r, 1.0 / 5.0);
    p = det_fma(p, r, -1.0 / 4.0);
    p = det_fma(p, r, 1.0 / 3.0);
    p = det_fma(p, r, -0.5);
    p = det_fma(p, r * r, r);  // r + r^2 (-1/2 + r/3 - ...)
    const double ed = (double)e;
    return det_fma(ed, 0.6931471803691238, t.y + det_fma(ed, 1.9082149292705877e-10, p));
}

// e^x: x = (32 q + j) ln2/32 + r, |r| <= ln2/64; e^x = 2^q 2^(j/32) (1 + expm1(r)), expm1 to r^6;
// scaled by 2^q in two exact steps (subnormal results round once).  (Parity probes only.)
static __device__ __noinline__ double det_exp(double x) {
    if (x < -745.2) return 0.0;
    if (x > 709.0) x = 709.0;
    const double kd = (double)__double2ll_rd(det_fma(x, 46.16624130844683, 0.5));  // floor(x 32/ln2 + 1/2)
    const double r = det_fma(-kd, 5.9631716539705866e-12, det_fma(-kd, 0.02166084938653512, x));
    double p = 1.0 / 720.0;
    p = det_fma(p, r, 1.0 / 120.0);
    p = det_fma(p, r, 1.0 / 24.0);
    p = det_fma(p, r, 1.0 / 6.0);
    p = det_fma(p, r, 0.5);
    p = det_fma(p, r, 1.0);
    p = p * r;  // expm1(r)
    const int k = (int)kd;
    const double tj = EXP_TAB_C[k & 31];
    const double y = det_fma(tj, p, tj);
    const int q = k >> 5;
    const int q1 = q / 2, q2 = q - q1;
    const double s1 = bits_to_double((uint64_t)(q1 + 1023) << 52);
    const double s2 = bits_to_double((uint64_t)(q2 + 1023) << 52);
    return (y * s1) * s2;
}

__device__ __forceinline__ double det_sqrt(double x) { return __dsqrt_rn(x); }
#endif  // __CUDACC__

// ------------------------------------------------------------------------------------------------
// log2 of a Gamma(shape, 1) draw, f32 (twin: oracle/c6_rng.hpp "The sampler the search uses").  The noise of a
// select step is normalised like a softmax, noise_k = 2^(L_k - Lmax) / sum_j 2^(L_j - Lmax), so a draw never
// has to exist as a number (with alpha = 0.03 almost all of them underflow even f64).
// ------------------------------------------------------------------------------------------------
struct GammaParams {  // per-run constants, computed once on the host (IEEE ops only)
    double shape;
    float b, inv_a, c1, am1;  // shape < 1: (e + a)/e, 1/a, b/a, a - 1, each rounded to f32 once
};
inline GammaParams gamma_params(double shape) {
    GammaParams g;
    g.shape = shape;
    const double b = (2.718281828459045 + shape) / 2.718281828459045;
    g.b = (float)b, g.inv_a = (float)(1.0 / shape), g.c1 = (float)(b / shape), g.am1 = (float)(shape - 1.0);
    return g;
}
constexpr uint32_t DOMAIN_TIE = 0x54494542u;
#ifdef __CUDACC__
__device__ __forceinline__ float unit_of(uint32_t w) { return det_fmaf((float)w, 0x1p-32f, 0x1p-33f); }    // (w + 1/2) 2^-32
__device__ __forceinline__ float unit24_of(uint32_t w) { return det_fmaf((float)w, 0x1p-24f, 0x1p-25f); }  // (w + 1/2) 2^-24, 24-bit w

// shape < 1 (the reference default is 0.03): Ahrens–Dieter GS rejection sampler — exact for every 0 < shape < 1,
// acceptance > 98% at 0.03.  With P = b U:
//   P <= 1 : X = P^(1/shape),          accept iff V <= e^-X      (1 - V >= 1 - e^-X)
//   P >  1 : X = -ln((b - P)/shape),   accept iff V <= X^(shape-1)
// Uniforms (twin: oracle/c6_rng.hpp gamma_uniforms): trial 0 of draw k is ONE word of Philox block (a, b, kb, 0),
// kb = (k >> 7) << 5 | (k & 31), word (k >> 5) & 3 — four draws per block: 24 bits of U (all the f32 conversion keeps) and the
// top byte of V; V's other 24 bits come from block (a, b, k, 0xFFFFFFFF) and are read by the full loop only.
// gamma_log2_fast tries trial 0 on the hot branch with a sufficient test and reports failure (~6% of the draws);
// gamma_log2_full is the whole loop.  Any sufficient test gives the same draws (a draw the test passes is the draw the loop
// accepts at trial 0): with v8 the top byte of V, 1 - V >= (255 - v8)/256, and the loop accepts when 1 - V >= d,
// d = 1 - e^-X <= X = 2^L.  Tabulated: L <= log2((255 - v8)/256) - 2^-20 (the margin covers ex2_det's and d's rounding);
// computed (C6_GAMMA_THR=0): (255 - v8)/256 >= 2^floor(L) (1 + frac L) >= X.
__device__ __forceinline__ bool gamma_log2_fast(const MathTab& T, const GammaParams& gp, uint32_t w, float& L) {
    const float P = gp.b * unit24_of(w >> 8);
    L = lg2_det(T, fminf(P, 1.0f)) * gp.inv_a;
#if C6_GAMMA_THR
    return P <= 1.0f && L <= T.thr[w & 0xFFu];
#endif
    const float Lc = fmaxf(L, -64.0f);
    const float kf = floorf(Lc);
    const float xub = __uint_as_float(__float_as_uint(1.0f + (Lc - kf)) + ((uint32_t)(int)kf << 23));
    return P <= 1.0f && (float)(255u - (w & 0xFFu)) * 0x1p-8f >= xub;
}
__device__ __forceinline__ uint32_t gamma_word0(uint32_t a, uint32_t b, uint32_t k, uint32_t k0, uint32_t k1) {
    const Philox4 r = philox4x32_10(a, b, ((k >> 7) << 5) | (k & 31u), 0u, k0, k1);
    const uint32_t h = (k >> 5) & 3u;
    return h == 0 ? r.x : h == 1 ? r.y : h == 2 ? r.z : r.w;
}
static __device__ __noinline__ float gamma_log2_full(const MathTab& T, float gb, float inv_a, float c1, float am1, uint32_t a,
                                                      uint32_t b, uint32_t k, uint32_t k0, uint32_t k1) {
    float L = 0.0f;
    for (uint32_t trial = 0; trial < 64u; ++trial) {
        float U, Uc, V, Vc;
        if (trial == 0) {
            const uint32_t w = gamma_word0(a, b, k, k0, k1);
            const uint32_t v = ((w & 0xFFu) << 24) | (philox4x32_10(a, b, k, 0xFFFFFFFFu, k0, k1).x >> 8);
            U = unit24_of(w >> 8), Uc = unit24_of(0xFFFFFFu - (w >> 8)), V = unit_of(v), Vc = unit_of(~v);
        } else {
            const Philox4 r = philox4x32_10(a, b, k, trial, k0, k1);
            U = unit_of(r.x), Uc = unit_of(~r.x), V = unit_of(r.y), Vc = unit_of(~r.y);
        }
        const float P = gb * U;
        if (P <= 1.0f) {
            L = lg2_det(T, P) * inv_a;
            if (L < -40.0f) break;  // 1 - e^-X < 2^-40 < the smallest 1 - V
            const float X = ex2_det(L);
            float d;
            if (X < 0x1p-6f)
                d = X * det_fmaf(X, det_fmaf(X, det_fmaf(X, -1.0f / 24.0f, 1.0f / 6.0f), -0.5f), 1.0f);
            else
                d = 1.0f - ex2_det(X * -1.44269504f);
            if (Vc >= d) break;
        } else {
            const float X = lg2_det(T, c1 * Uc) * -0.693147182f;
            L = lg2_det(T, fmaxf(X, 0x1p-100f));
            if (lg2_det(T, V) <= am1 * L) break;
        }
    }
    return L;
}

// shape >= 1: Marsaglia–Tsang (2000) on normals from the Marsaglia polar method, in f64; trial t reads blocks
// 2t (polar pair) and 2t+1 (acceptance uniform).  Not on the default path; kept out of line.  Returns the draw.
static __device__ __noinline__ double gamma_large(double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1) {
    const double d = shape - 1.0 / 3.0;
    const double cc = 1.0 / det_sqrt(9.0 * d);
    double out = d;
    for (uint32_t trial = 0; trial < 64u; ++trial) {
        const Philox4 r = philox4x32_10(a, b, c, 2u * trial, k0, k1);
        const double p = 2.0 * u64_to_open01(r.x, r.y) - 1.0, q = 2.0 * u64_to_open01(r.z, r.w) - 1.0;
        const double s = p * p + q * q;
        if (s >= 1.0 || s == 0.0) continue;
        const double x = p * det_sqrt(-2.0 * det_log(s) / s);
        const double vc = 1.0 + cc * x;
        if (vc <= 0.0) continue;
        const double v = vc * vc * vc;
        const Philox4 r2 = philox4x32_10(a, b, c, 2u * trial + 1u, k0, k1);
        const double u = u64_to_open01(r2.x, r2.y);
        const double x2 = x * x;
        out = d * v;
        if (u < 1.0 - 0.0331 * x2 * x2) break;
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) break;
    }
    return out;
}
__device__ __forceinline__ float gamma_large_log2(double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1) {
    return (float)(det_log(gamma_large(shape, a, b, c, k0, k1)) * 1.4426950408889634);
}

// uniform pick among n exactly tied children (alphazero_policy/mod.rs:291-295): floor(u n / 2^32)
__device__ __forceinline__ uint32_t tie_pick(uint32_t sim, uint32_t depth, uint32_t seed, uint32_t game, uint32_t n) {
    return __umulhi(philox4x32_10(sim, depth, 0u, 0u, seed ^ DOMAIN_TIE, game).x, n);
}
#endif  // __CUDACC__

}  // namespace c6
