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
// LC/game/search_winner/mod.rs:213-217)
template <int D>
__host__ __device__ __forceinline__ void run6(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
    uint32_t t[8], m2[8], m4[8];
    shl256<D>(b, t);
#pragma unroll
    for (int w = 0; w < 8; ++w) m2[w] = b[w] & t[w];
    shl256<2 * D>(m2, t);
#pragma unroll
    for (int w = 0; w < 8; ++w) m4[w] = m2[w] & t[w];
    shl256<4 * D>(m2, t);
#pragma unroll
    for (int w = 0; w < 8; ++w) acc[w] |= m4[w] & t[w];
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
// search(): winner = colour whose first run-end cell comes first in row-major order
__host__ __device__ inline int win_check(const uint32_t (&B)[8], const uint32_t (&W)[8]) {
    uint32_t eb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ew[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    run6<1>(B, eb), run6<16>(B, eb), run6<17>(B, eb), run6<15>(B, eb);
    run6<1>(W, ew), run6<16>(W, ew), run6<17>(W, ew), run6<15>(W, ew);
    int fb = first_bit256(eb), fw = first_bit256(ew);
    if (fb == 256 && fw == 256) return 0;
    return fb < fw ? -1 : 1;
}

// The same decision with fewer instructions (config-3a kernel; 420 ALU instructions per board made k_win_check as much issue-
// as bandwidth-bound, profiles/r01p_micro.jsonl): the vertical direction shifts by 16 / 32 / 64 bits, and the last two are whole
// words — register renaming, no instruction; and since a cell holds one colour, the winner is the colour owning the lowest set
// bit of (black run ends | white run ends): one scan over eight words instead of two.
__host__ __device__ __forceinline__ void run6_vertical(const uint32_t (&b)[8], uint32_t (&acc)[8]) {
    uint32_t t[8], m2[8];
    shl256<16>(b, t);
#pragma unroll
    for (int w = 0; w < 8; ++w) m2[w] = b[w] & t[w];
#pragma unroll
    for (int w = 2; w < 8; ++w) acc[w] |= m2[w] & m2[w - 1] & m2[w - 2];  // m2 & (m2 << 32) & (m2 << 64); words 0, 1 have no room for six rows
}
__host__ __device__ inline int win_check_fast(const uint32_t (&B)[8], const uint32_t (&W)[8]) {
    uint32_t eb[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ew[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    run6<1>(B, eb), run6_vertical(B, eb), run6<17>(B, eb), run6<15>(B, eb);
    run6<1>(W, ew), run6_vertical(W, ew), run6<17>(W, ew), run6<15>(W, ew);
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
