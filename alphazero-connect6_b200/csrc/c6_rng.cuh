// Counter-based random sources and fixed-polynomial f64 math for the engine (device side).
//
// Replaces the reference's `thread_rng()` uses on the hot path:
//   Dirichlet noise at every select step   LC/policy/alphazero_policy/mod.rs:262,283
//   RandomEvaluator                        LC/policy/alphazero_policy/mod.rs:137-169
// with Philox4x32-10 streams keyed by (seed, game) and counted by (simulation, depth, child), so a
// game's search is a pure function of its id.  The Gamma sampler is rand 0.6's algorithm
// (Marsaglia–Tsang + U^(1/alpha) boost for alpha < 1) on normals from the Marsaglia polar method;
// log/exp are fixed polynomials over IEEE + - * / only, and this translation unit is compiled with
// -fmad=false, so every value is reproducible bit for bit on a CPU (tests/ compare against the
// oracle's independent restatement).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c6 {

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                         uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0, c1 = lo1, c2 = n2, c3 = lo0;
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
__host__ __device__ __forceinline__ double u64_to_open01(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
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

__host__ __device__ inline double det_log(double x) {
    int e = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) {
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)((b >> 52) & 0x7FF) - 1023;
    double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    double p = 1.0 / 23.0;
    p = p * z + 1.0 / 21.0;
    p = p * z + 1.0 / 19.0;
    p = p * z + 1.0 / 17.0;
    p = p * z + 1.0 / 15.0;
    p = p * z + 1.0 / 13.0;
    p = p * z + 1.0 / 11.0;
    p = p * z + 1.0 / 9.0;
    p = p * z + 1.0 / 7.0;
    p = p * z + 1.0 / 5.0;
    p = p * z + 1.0 / 3.0;
    p = p * z + 1.0;
    double lm = 2.0 * s * p;
    return (double)e * 0.6931471803691238 + ((double)e * 1.9082149292705877e-10 + lm);
}

__host__ __device__ __forceinline__ double det_floor(double x) {
    double t = (double)(long long)x;
    return (t > x) ? t - 1.0 : t;
}

__host__ __device__ inline double det_exp(double x) {
    if (x < -745.2) return 0.0;
    if (x > 709.0) x = 709.0;
    double kd = det_floor(x * 1.4426950408889634 + 0.5);
    double r = (x - kd * 0.6931471803691238) - kd * 1.9082149292705877e-10;
    double p = 1.0 / 6227020800.0;
    p = p * r + 1.0 / 479001600.0;
    p = p * r + 1.0 / 39916800.0;
    p = p * r + 1.0 / 3628800.0;
    p = p * r + 1.0 / 362880.0;
    p = p * r + 1.0 / 40320.0;
    p = p * r + 1.0 / 5040.0;
    p = p * r + 1.0 / 720.0;
    p = p * r + 1.0 / 120.0;
    p = p * r + 1.0 / 24.0;
    p = p * r + 1.0 / 6.0;
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r + 1.0;
    int k = (int)kd;
    int k1 = k / 2, k2 = k - k1;
    double s1 = bits_to_double((uint64_t)(k1 + 1023) << 52);
    double s2 = bits_to_double((uint64_t)(k2 + 1023) << 52);
    return (p * s1) * s2;
}

__host__ __device__ __forceinline__ double det_sqrt(double x) {
#ifdef __CUDA_ARCH__
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}

struct PhiloxStream {
    uint32_t c0, c1, c2, blk, k0, k1;
    uint32_t b0, b1, b2, b3;
    int pos;
    __host__ __device__ PhiloxStream(uint32_t a, uint32_t b, uint32_t c, uint32_t key0, uint32_t key1)
        : c0(a), c1(b), c2(c), blk(0), k0(key0), k1(key1), b0(0), b1(0), b2(0), b3(0), pos(4) {}
    __host__ __device__ double next_open01() {
        if (pos >= 4) {
            Philox4 r = philox4x32_10(c0, c1, c2, blk++, k0, k1);
            b0 = r.x, b1 = r.y, b2 = r.z, b3 = r.w;
            pos = 0;
        }
        double u = (pos == 0) ? u64_to_open01(b0, b1) : u64_to_open01(b2, b3);
        pos += 2;
        return u;
    }
};

__host__ __device__ inline double normal_polar(PhiloxStream& st) {
    for (;;) {
        double a = 2.0 * st.next_open01() - 1.0;
        double b = 2.0 * st.next_open01() - 1.0;
        double s = a * a + b * b;
        if (s >= 1.0 || s == 0.0) continue;
        double f = det_sqrt(-2.0 * det_log(s) / s);
        return a * f;
    }
}

__host__ __device__ inline double gamma_large(PhiloxStream& st, double shape) {
    double d = shape - 1.0 / 3.0;
    double c = 1.0 / det_sqrt(9.0 * d);
    for (;;) {
        double x = normal_polar(st);
        double vc = 1.0 + c * x;
        if (vc <= 0.0) continue;
        double v = vc * vc * vc;
        double u = st.next_open01();
        double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
        if (det_log(u) < 0.5 * x2 + d * (1.0 - v + det_log(v))) return d * v;
    }
}

__host__ __device__ inline double gamma_sample(PhiloxStream& st, double shape) {
    if (shape < 1.0) {
        double g = gamma_large(st, shape + 1.0);
        double u = st.next_open01();
        return g * det_exp(det_log(u) / shape);
    }
    return gamma_large(st, shape);
}

// One Gamma(shape) draw from the stream with counter (a, b, c).  Kept out of line: eight inlined
// copies per select step thrash the instruction cache (ncu: stall_no_instruction dominated).
__host__ __device__ __noinline__ inline double gamma_draw(uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1,
                                                         double shape) {
    PhiloxStream st(a, b, c, k0, k1);
    return gamma_sample(st, shape);
}

}  // namespace c6
