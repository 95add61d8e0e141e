// Bit-packed Connect6 board helpers shared by the kernels and the host-side game mirror.
//   search()           LC/game/search_winner/mod.rs:201-246   -> win_check (shift/AND run detection)
//   turn rule          LC/game/game_impl/mod.rs:125-131,158-167 -> turn_of
//   augment_way8 maps  LC/policy/alphazero_policy/augment.rs:7-73 -> sym_fwd / sym_inv
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace c6 {

// ------------------------------------------------------------------------------------------------
// Board helpers.  Bit 16*row+col; word j holds rows 2j and 2j+1; lane l of a warp owns bit l of
// every word, i.e. cells (2j + (l>>4), l&15) for j = 0..7.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t valid_mask(int j) { return j < 7 ? 0x7FFF7FFFu : 0x00007FFFu; }
// colour (-1 Black / +1 White) that places the next stone on a board with `stones` stones
// (LC/game/game_impl/mod.rs:125-131,158-167: Black 1 stone, then 2 each).
__host__ __device__ __forceinline__ int turn_of(int stones) {
    return stones == 0 ? -1 : ((((stones - 1) >> 1) & 1) ? -1 : 1);
}
__host__ __device__ __forceinline__ int cell_of_bit(uint32_t bit) { return (int)(bit >> 4) * 15 + (int)(bit & 15); }
__host__ __device__ __forceinline__ uint32_t bit_of_cell(int cell) { return (uint32_t)((cell / 15) * 16 + cell % 15); }

template <int S>
__host__ __device__ __forceinline__ void shl256(const uint32_t (&b)[8], uint32_t (&o)[8]) {
    constexpr int ws = S >> 5, bs = S & 31;
#pragma unroll
    for (int w = 7; w >= 0; --w) {
        uint32_t hi = (w - ws >= 0) ? b[(w - ws >= 0) ? (w - ws) : 0] : 0u;
        uint32_t lo = (w - ws - 1 >= 0) ? b[(w - ws - 1 >= 0) ? (w - ws - 1) : 0] : 0u;
        o[w] = bs ? ((hi << bs) | (lo >> ((32 - bs) & 31))) : hi;
    }
}
// bits at cells that END a run of >= 6 in stride D (the cells where search() returns,
// LC/game/search_winner/mod.rs:213-217): x3 = b & b<<D & b<<2D (three in a row ending here), six = x3 & x3<<3D.
// Three shifts and two 3-input logic operations (LOP3) per word — the shift/AND ladder over pairs (b & b<<D, then <<2D, <<4D)
// costs three of each; on the GPU both kinds issue on the ALU pipe, which is what bounds k_win_check.
// Column 15 of every row is never occupied, so a run cannot continue across a row end in any of the four strides.
template <int D>
__host__ __device__ __forceinline__ void run6(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
    uint32_t t1[8], t2[8], x3[8];
    shl256<D>(b, t1);
    shl256<2 * D>(b, t2);
#pragma unroll
    for (int w = 0; w < 8; ++w) x3[w] = b[w] & t1[w] & t2[w];
    shl256<3 * D>(x3, t1);
#pragma unroll
    for (int w = 0; w < 8; ++w) acc[w] |= x3[w] & t1[w];
}
// stride 1 (along a row): bits 15 and 31 of every word are the unoccupied column, so nothing a run could use crosses a word
// boundary and the shifts are plain 32-bit ones — which the compiler may issue as multiplies on the FMA pipe (IMAD.SHL).
template <>
__host__ __device__ __forceinline__ void run6<1>(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t x3 = b[w] & (b[w] << 1) & (b[w] << 2);
        acc[w] |= x3 & (x3 << 3);
    }
}
// stride 16 (down a column): 32 and 64 bits are whole words — register renaming, no instruction
template <>
__host__ __device__ __forceinline__ void run6<16>(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
    uint32_t t[8], m2[8];
    shl256<16>(b, t);
#pragma unroll
    for (int w = 0; w < 8; ++w) m2[w] = b[w] & t[w];
#pragma unroll
    for (int w = 2; w < 8; ++w) acc[w] |= m2[w] & m2[w - 1] & m2[w - 2];  // m2 & (m2 << 32) & (m2 << 64); words 0, 1 have no room for six rows
}
__host__ __device__ __forceinline__ int first_bit256(const uint32_t (&a)[8]) {
    int r = 256;
#pragma unroll
    for (int w = 7; w >= 0; --w)
        if (a[w]) {
#ifdef __CUDA_ARCH__
            r = w * 32 + (__ffs((int)a[w]) - 1);
#else
            r = w * 32 + __builtin_ctz(a[w]);
#endif
        }
    return r;
}
// All four strides of one colour, sharing the board moved one row down (s = b << 16).  A move by k rows and c columns is a
// 16-bit move (k odd: the view s) and / or a move by whole words (k / 2: register renaming) followed by a plain 32-bit shift
// by |c| inside each word: what such a shift drops at a word boundary, and what it carries across the row boundary in the
// middle of a word, only ever lands on cells whose run would have to pass the unoccupied column 15 — the AND with the board
// clears them (checked against a cell-by-cell scan, tests/cpp/win_check_scan.cpp).  Against run6<17> / run6<15>
// this trades two funnel shifts per word (ALU pipe, like the LOP3s) for plain shifts, which the FMA pipe can take
// (IMAD.SHL for left shifts; right shifts as IMAD.HI when C6_SHR_MULHI).
#ifndef C6_SHR_MULHI
#define C6_SHR_MULHI 1
#endif
template <int K>
__host__ __device__ __forceinline__ uint32_t shr_plain(uint32_t x) {
#if defined(__CUDA_ARCH__) && C6_SHR_MULHI
    return __umulhi(x, 1u << (32 - K));
#else
    return x >> K;
#endif
}
__host__ __device__ __forceinline__ void run6_all(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
    uint32_t s[8], x[8], sx[8];
    shl256<16>(b, s);
    // stride 16 (down a column): pairs, then three pairs two rows apart
#pragma unroll
    for (int w = 0; w < 8; ++w) x[w] = b[w] & s[w];
#pragma unroll
    for (int w = 2; w < 8; ++w) acc[w] |= x[w] & x[w - 1] & x[w - 2];
    // stride 1 (along a row)
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t x3 = b[w] & (b[w] << 1) & (b[w] << 2);
        acc[w] |= x3 & (x3 << 3);
    }
    // stride 17 (one row down, one column right): b << 17 = s << 1, b << 34 = one word up << 2, x << 51 = (x << 16) one word up << 3
    x[0] = 0;
#pragma unroll
    for (int w = 1; w < 8; ++w) x[w] = b[w] & (s[w] << 1) & (b[w - 1] << 2);
    shl256<16>(x, sx);
#pragma unroll
    for (int w = 2; w < 8; ++w) acc[w] |= x[w] & (sx[w - 1] << 3);
    // stride 15 (one row down, one column left): the same with right shifts
#pragma unroll
    for (int w = 1; w < 8; ++w) x[w] = b[w] & shr_plain<1>(s[w]) & shr_plain<2>(b[w - 1]);
    shl256<16>(x, sx);
#pragma unroll
    for (int w = 2; w < 8; ++w) acc[w] |= x[w] & shr_plain<3>(sx[w - 1]);
}
// position of the n-th (0-based) set bit of x, n < popc(x): binary search on the popcounts of the lower halves, five branch-free
// steps.  (`__fns` does the same in a data-dependent software loop, and eight predicated copies of it — one per board word —
// left 4 to 8 of a warp's 32 lanes active in every loop that picks a random empty cell.)
__host__ __device__ __forceinline__ uint32_t nth_set_bit(uint32_t x, uint32_t n) {
    uint32_t pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const uint32_t lo = x & ((1u << w) - 1u);
#ifdef __CUDA_ARCH__
        const uint32_t c = (uint32_t)__popc(lo);
#else
        const uint32_t c = (uint32_t)__builtin_popcount(lo);
#endif
        const bool up = n >= c;
        n -= up ? c : 0u, pos += up ? (uint32_t)w : 0u, x = up ? (x >> w) : lo;
    }
    return pos;
}
// bit position of the pick-th (0-based) empty cell in row-major order, pick < number of empty cells
__host__ __device__ __forceinline__ uint32_t pick_empty_cell(const uint32_t (&B)[8], const uint32_t (&W)[8], uint32_t pick) {
    uint32_t word = 0, mask = 0;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = ~(B[j] | W[j]) & valid_mask(j);
#ifdef __CUDA_ARCH__
        const uint32_t c = (uint32_t)__popc(e);
#else
        const uint32_t c = (uint32_t)__builtin_popcount(e);
#endif
        const bool here = !found && pick < c;
        if (here) word = (uint32_t)j, mask = e;
        found = found || here;
        if (!found) pick -= c;
    }
    return 32u * word + nth_set_bit(mask, pick);
}

// search(): winner = colour whose first run-end cell comes first in row-major order
__host__ __device__ inline int win_check(const uint32_t (&B)[8], const uint32_t (&W)[8]) {
    uint32_t eb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ew[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    run6<1>(B, eb), run6<16>(B, eb), run6<17>(B, eb), run6<15>(B, eb);
    run6<1>(W, ew), run6<16>(W, ew), run6<17>(W, ew), run6<15>(W, ew);
    int fb = first_bit256(eb), fw = first_bit256(ew);
    if (fb == 256 && fw == 256) return 0;
    return fb < fw ? -1 : 1;
}

// The same decision with one scan (config-3a kernel): since a cell holds one colour, the winner is the colour owning the
// lowest set bit of (black run ends | white run ends) — one pass over eight words instead of two.
__host__ __device__ inline int win_check_fast(const uint32_t (&B)[8], const uint32_t (&W)[8]) {
    uint32_t eb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ew[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    run6_all(B, eb), run6_all(W, ew);
    uint32_t first = 0, first_b = 0;
#pragma unroll
    for (int w = 7; w >= 0; --w) {
        const uint32_t e = eb[w] | ew[w];
        if (e) first = e, first_b = eb[w];
    }
    if (!first) return 0;
    return (first_b & (first & (0u - first))) ? -1 : 1;
}

// symmetry s of augment_way8 (LC/policy/alphazero_policy/augment.rs:60-73): where the content of
// original cell (r,c) lands in plane s.
__host__ __device__ __forceinline__ void sym_fwd(int s, int r, int c, int& rr, int& cc) {
    switch (s) {
        case 0: rr = 14 - c, cc = r; break;
        case 1: rr = 14 - c, cc = 14 - r; break;
        case 2: rr = 14 - r, cc = 14 - c; break;
        case 3: rr = 14 - r, cc = c; break;
        case 4: rr = c, cc = 14 - r; break;
        case 5: rr = c, cc = r; break;
        case 6: rr = r, cc = c; break;
        default: rr = r, cc = 14 - c; break;
    }
}
// inverse: which original cell is shown at (rr,cc) of plane s
__host__ __device__ __forceinline__ void sym_inv(int s, int rr, int cc, int& r, int& c) {
    switch (s) {
        case 0: r = cc, c = 14 - rr; break;
        case 1: r = 14 - cc, c = 14 - rr; break;
        case 2: r = 14 - rr, c = 14 - cc; break;
        case 3: r = 14 - rr, c = cc; break;
        case 4: r = 14 - cc, c = rr; break;
        case 5: r = cc, c = rr; break;
        case 6: r = rr, c = cc; break;
        default: r = rr, c = 14 - cc; break;
    }
}

}  // namespace c6
