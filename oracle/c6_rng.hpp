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
//   * rand 0.6 `Gamma` for shape < 1 (GammaSmallShape): Marsaglia–Tsang (2000) sample of
//     Gamma(shape+1) times U^(1/shape); `Dirichlet` = n independent Gamma(alpha,1) normalised
//     by their sum, in f64 (rand 0.6.x is a caret dependency, no lockfile: LC/Cargo.toml:12).
//     rand draws its normal deviates from a ziggurat table; here they come from the Marsaglia
//     polar method so that only + - * / sqrt and the fixed-polynomial log/exp below are needed
//     (bit-identical on x86-64 and on the GPU when neither side contracts FMAs).
// Everything in this header must stay bit-for-bit equal to the device twin in
// alphazero-connect6_b200/csrc/c6_rng.cuh; tests/test_rng_parity.py checks that on the GPU.
#pragma once
#include <cstdint>
#include <cstring>

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
// two u32 -> f64 uniform on (0, 1): 53 bits, never 0 or 1.
inline double u64_to_open01(uint32_t lo, uint32_t hi) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// ----------------------------------------------------------------------------------------------
// Fixed-polynomial log / exp in f64 (only IEEE + - * / and bit moves: same bits on CPU and GPU).
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

inline double det_log(double x) {
    // x > 0, finite.
    int e = 0;
    uint64_t b = double_to_bits(x);
    if ((b >> 52) == 0) {  // subnormal: scale by 2^54
        x = x * 18014398509481984.0;
        b = double_to_bits(x);
        e = -54;
    }
    e += (int)((b >> 52) & 0x7FF) - 1023;
    double m = bits_to_double((b & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);  // [1,2)
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    // 2*atanh(s) = 2s(1 + z/3 + z^2/5 + ...), |s| <= 0.1716
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

inline double det_floor(double x) {
    // |x| < 2^31 here.  Truncate then fix up; avoids relying on libm/device floor.
    double t = (double)(int64_t)x;
    return (t > x) ? t - 1.0 : t;
}

inline double det_exp(double x) {
    if (x < -745.2) return 0.0;
    if (x > 709.0) x = 709.0;
    double kd = det_floor(x * 1.4426950408889634 + 0.5);
    double r = (x - kd * 0.6931471803691238) - kd * 1.9082149292705877e-10;  // |r| <= 0.3466
    double p = 1.0 / 6227020800.0;  // 1/13!
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
    // scale by 2^k in two exact-power steps so that subnormal results round once, identically.
    int k1 = k / 2, k2 = k - k1;
    double s1 = bits_to_double((uint64_t)(k1 + 1023) << 52);
    double s2 = bits_to_double((uint64_t)(k2 + 1023) << 52);
    return (p * s1) * s2;
}

// ----------------------------------------------------------------------------------------------
// Gamma(alpha < 1 or >= 1, scale 1) from a counter-based stream.
// Stream = Philox blocks (c0,c1,c2, blk=0,1,2,...) consumed two u32 at a time.
// ----------------------------------------------------------------------------------------------
struct PhiloxStream {
    uint32_t c0, c1, c2, blk, k0, k1;
    uint32_t buf[4];
    int pos;
    PhiloxStream(uint32_t a, uint32_t b, uint32_t c, uint32_t key0, uint32_t key1)
        : c0(a), c1(b), c2(c), blk(0), k0(key0), k1(key1), pos(4) {}
    double next_open01() {
        if (pos >= 4) {
            Philox4 r = philox4x32_10(c0, c1, c2, blk++, k0, k1);
            buf[0] = r.v[0], buf[1] = r.v[1], buf[2] = r.v[2], buf[3] = r.v[3];
            pos = 0;
        }
        double u = u64_to_open01(buf[pos], buf[pos + 1]);
        pos += 2;
        return u;
    }
};

inline double det_sqrt(double x);  // IEEE sqrt; defined below via <cmath> (correctly rounded)

inline double normal_polar(PhiloxStream& st) {
    for (;;) {
        double a = 2.0 * st.next_open01() - 1.0;
        double b = 2.0 * st.next_open01() - 1.0;
        double s = a * a + b * b;
        if (s >= 1.0 || s == 0.0) continue;
        double f = det_sqrt(-2.0 * det_log(s) / s);
        return a * f;
    }
}

// Marsaglia–Tsang for shape >= 1 (rand 0.6 GammaLargeShape).
inline double gamma_large(PhiloxStream& st, double shape) {
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

// rand 0.6 Gamma::new(shape, 1.0): shape < 1 -> GammaSmallShape, else large (shape == 1 is
// Exp(1) in rand; Marsaglia–Tsang is used here for it too — same distribution).
inline double gamma_sample(PhiloxStream& st, double shape) {
    if (shape < 1.0) {
        double g = gamma_large(st, shape + 1.0);
        double u = st.next_open01();
        return g * det_exp(det_log(u) / shape);
    }
    return gamma_large(st, shape);
}

}  // namespace c6o

#include <cmath>
namespace c6o {
inline double det_sqrt(double x) { return std::sqrt(x); }
}  // namespace c6o
