// Counter-based random sources and fixed-polynomial f64 math for the engine (device side).
//
// Replaces the reference's `thread_rng()` uses on the hot path:
//   Dirichlet noise at every select step   LC/policy/alphazero_policy/mod.rs:262,283
//   RandomEvaluator                        LC/policy/alphazero_policy/mod.rs:137-169
// with Philox4x32-10 streams keyed by (seed, game) and counted by (simulation, depth, child), so a
// game's search is a pure function of its id.  Dirichlet(alpha; C) = C independent Gamma(alpha, 1)
// draws normalised by their sum in f64, as rand 0.6 builds it.  The Gamma draws use exact rejection
// samplers chosen for the GPU: Ahrens–Dieter GS for alpha < 1 (the default 0.03; ~175 instructions a
// draw against ~700 for Marsaglia–Tsang + boost — the first ncu profile had the sampler at two thirds
// of all instructions), Marsaglia–Tsang on polar normals for alpha >= 1.  log/exp are fixed
// polynomials over IEEE + - * / fma only, and this translation unit is compiled with -fmad=false, so
// every value is reproducible bit for bit on a CPU (tests/ compare against the oracle's independent
// restatement).  rand's own stream is unseedable (thread_rng) and unpinned by any reference test, so
// only the distribution is contractual.
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
// contracts): one rounding, identical on sm_100a and x86-64 FMA3 (the oracle uses __builtin_fma).
__device__ __forceinline__ double det_fma(double a, double b, double c) { return __fma_rn(a, b, c); }

// Tables of the log / exp below, staged in shared memory by every kernel that samples noise (per-lane
// indices would serialise in the constant cache).
struct MathTab {
    double2 logt[64];  // {1/c_i, ln c_i}, c_i = 1 + (2i+1)/128
    double expt[32];   // 2^(j/32)
};
__device__ __forceinline__ void load_math_tab(MathTab& T) {
    if (threadIdx.x < 64) T.logt[threadIdx.x] = LOG_TAB_C[threadIdx.x];  // blocks have at least 64 threads
    if (threadIdx.x < 32) T.expt[threadIdx.x] = EXP_TAB_C[threadIdx.x];
}

// Coefficients whose low word is not zero cannot be DFMA immediates; as literals the compiler rebuilds each of
// them with two moves per use (r01g capture: ~30 of the 168 instructions of a draw), from the constant bank they
// are operands.  Same bits either way.
#ifndef C6_CBANK_COEF
#define C6_CBANK_COEF 1
#endif
static __device__ __constant__ double MATH_K[13] = {1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, 1.0 / 3.0, 0.6931471803691238,
                                                    1.9082149292705877e-10, 46.16624130844683, 5.9631716539705866e-12,
                                                    0.02166084938653512, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0};
#if C6_CBANK_COEF
#define C6_K(i, v) MATH_K[i]
#else
#define C6_K(i, v) (v)
#endif

// natural log of a positive finite double, division-free: x = 2^e m, m in [1,2) falls in one of 64
// intervals with centre c_i; r = m/c_i - 1 (|r| < 2^-7) by one fma with the tabulated reciprocal;
// ln x = e ln2 + ln c_i + log1p(r), log1p as a degree-8 alternating series.  Absolute error ~1e-16.
// NORMAL: the caller guarantees a normal (not subnormal) argument, which skips the rescaling test; same value.
template <bool NORMAL = false>
__device__ __forceinline__ double det_log(const MathTab& T, double x) {
    int e = 0;
    uint64_t b = double_to_bits(x);
    if (!NORMAL && (b >> 52) == 0) {
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)((b >> 52) & 0x7FF) - 1023;
    const double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    const double2 t = T.logt[(uint32_t)(b >> 46) & 63u];
    const double r = det_fma(m, t.x, -1.0);
    double p = -1.0 / 8.0;
    p = det_fma(p, r, C6_K(0, 1.0 / 7.0));
    p = det_fma(p, r, C6_K(1, -1.0 / 6.0));
    p = det_fma(p, r, C6_K(2, 1.0 / 5.0));
    p = det_fma(p, r, -1.0 / 4.0);
    p = det_fma(p, r, C6_K(3, 1.0 / 3.0));
    p = det_fma(p, r, -0.5);
    p = det_fma(p, r * r, r);  // r + r^2 (-1/2 + r/3 - ...)
    const double ed = (double)e;
    return det_fma(ed, C6_K(4, 0.6931471803691238), t.y + det_fma(ed, C6_K(5, 1.9082149292705877e-10), p));
}

// e^x: x = (32 q + j) ln2/32 + r, |r| <= ln2/64; e^x = 2^q 2^(j/32) (1 + expm1(r)), expm1 to r^6;
// scaled by 2^q in two exact steps (subnormal results round once).
// NONPOS: the caller guarantees x <= 0, which skips the overflow clamp; same value.
template <bool NONPOS = false>
__device__ __forceinline__ double det_exp(const MathTab& T, double x) {
    if (x < -745.2) return 0.0;
    if (!NONPOS && x > 709.0) x = 709.0;
    const double kd = (double)__double2ll_rd(det_fma(x, C6_K(6, 46.16624130844683), 0.5));  // floor(x 32/ln2 + 1/2)
    const double r = det_fma(-kd, C6_K(7, 5.9631716539705866e-12), det_fma(-kd, C6_K(8, 0.02166084938653512), x));
    double p = C6_K(9, 1.0 / 720.0);
    p = det_fma(p, r, C6_K(10, 1.0 / 120.0));
    p = det_fma(p, r, C6_K(11, 1.0 / 24.0));
    p = det_fma(p, r, C6_K(12, 1.0 / 6.0));
    p = det_fma(p, r, 0.5);
    p = det_fma(p, r, 1.0);
    p = p * r;  // expm1(r)
    const int k = (int)kd;
    const double tj = T.expt[k & 31];
    const double y = det_fma(tj, p, tj);
    const int q = k >> 5;
    const int q1 = q / 2, q2 = q - q1;
    const double s1 = bits_to_double((uint64_t)(q1 + 1023) << 52);
    const double s2 = bits_to_double((uint64_t)(q2 + 1023) << 52);
    return (y * s1) * s2;
}

__device__ __forceinline__ double det_sqrt(double x) { return __dsqrt_rn(x); }

// ------------------------------------------------------------------------------------------------
// Gamma(shape, 1) draw number (a, b, c) of the stream keyed (k0, k1): Philox block `trial` of counter
// (a, b, c, trial) feeds trial number `trial` of a rejection sampler, so a draw is a pure function of
// its coordinates and costs one Philox call per trial.
// ------------------------------------------------------------------------------------------------
#endif  // __CUDACC__
struct GammaParams {  // per-run constants, computed once on the host (IEEE ops only)
    double shape, inv_shape, b;  // b = (e + shape) / e
};
inline GammaParams gamma_params(double shape) {
    GammaParams g;
    g.shape = shape;
    g.inv_shape = 1.0 / shape;
    g.b = (2.718281828459045 + shape) / 2.718281828459045;
    return g;
}
#ifdef __CUDACC__

// shape < 1 (the reference default is 0.03): Ahrens–Dieter GS rejection sampler — exact for every
// 0 < shape < 1, acceptance > 98% at 0.03.  With P = b U:
//   P <= 1 : X = P^(1/shape),           accept iff V <= e^-X   (squeeze: e^-X >= 1 - X + X^2/2 - X^3/6)
//   P >  1 : X = -ln((b - P)/shape),    accept iff V <= X^(shape-1)
// gamma_small_fast tries trial 0 on the hot branch only and reports failure; gamma_small is the full loop.
// (split in two so that a caller can have the Philox block of its next draw in flight under the f64 chain of this one)
__device__ __forceinline__ bool gamma_small_fast(const MathTab& T, const GammaParams& gp, const Philox4& r, double& x) {
    const double u = u64_to_open01(r.x, r.y), v = u64_to_open01(r.z, r.w);
    const double p = gp.b * u;
    x = 0.0;
    if (p > 1.0) return false;
    // p = b u with u >= 2^-53: normal; ln p <= 0
    x = det_exp<true>(T, det_log<true>(T, p) * gp.inv_shape);
    return v <= det_fma(x, det_fma(x, det_fma(x, -1.0 / 6.0, 0.5), -1.0), 1.0);
}
__device__ __forceinline__ bool gamma_small_fast(const MathTab& T, const GammaParams& gp, uint32_t a, uint32_t b, uint32_t c,
                                                 uint32_t k0, uint32_t k1, double& x) {
    return gamma_small_fast(T, gp, philox4x32_10(a, b, c, 0u, k0, k1), x);
}
static __device__ __noinline__ double gamma_small(const MathTab& T, double shape, double inv_shape, double bb, uint32_t a, uint32_t b,
                                           uint32_t c, uint32_t k0, uint32_t k1) {
    double x = 0.0;
    for (uint32_t trial = 0; trial < 64u; ++trial) {
        const Philox4 r = philox4x32_10(a, b, c, trial, k0, k1);
        const double u = u64_to_open01(r.x, r.y), v = u64_to_open01(r.z, r.w);
        const double p = bb * u;
        if (p <= 1.0) {
            x = det_exp(T, det_log(T, p) * inv_shape);
            if (v <= det_fma(x, det_fma(x, det_fma(x, -1.0 / 6.0, 0.5), -1.0), 1.0)) break;
            if (v <= det_exp(T, -x)) break;
        } else {
            x = -det_log(T, (bb - p) * inv_shape);
            if (v <= det_exp(T, (shape - 1.0) * det_log(T, x))) break;
        }
    }
    return x;
}

// shape >= 1: Marsaglia–Tsang (2000) on normals from the Marsaglia polar method; trial t reads blocks
// 2t (polar pair) and 2t+1 (acceptance uniform).  Not on the default path; kept out of line.
static __device__ __noinline__ double gamma_large(const MathTab& T, double shape, uint32_t a, uint32_t b, uint32_t c, uint32_t k0,
                                           uint32_t k1) {
    const double d = shape - 1.0 / 3.0;
    const double cc = 1.0 / det_sqrt(9.0 * d);
    double out = d;
    for (uint32_t trial = 0; trial < 64u; ++trial) {
        const Philox4 r = philox4x32_10(a, b, c, 2u * trial, k0, k1);
        const double p = 2.0 * u64_to_open01(r.x, r.y) - 1.0, q = 2.0 * u64_to_open01(r.z, r.w) - 1.0;
        const double s = p * p + q * q;
        if (s >= 1.0 || s == 0.0) continue;
        const double x = p * det_sqrt(-2.0 * det_log(T, s) / s);
        const double vc = 1.0 + cc * x;
        if (vc <= 0.0) continue;
        const double v = vc * vc * vc;
        const Philox4 r2 = philox4x32_10(a, b, c, 2u * trial + 1u, k0, k1);
        const double u = u64_to_open01(r2.x, r2.y);
        const double x2 = x * x;
        out = d * v;
        if (u < 1.0 - 0.0331 * x2 * x2) break;
        if (det_log(T, u) < 0.5 * x2 + d * (1.0 - v + det_log(T, v))) break;
    }
    return out;
}
#endif  // __CUDACC__

}  // namespace c6
