// Stand-alone kernels of the path (BASELINE.json configs[2]): bit-packed win detection over a batch
// of boards and random playouts from the empty board.
//   search()            LC/game/search_winner/mod.rs:201-246
//   rollout loop        LC/policy/default_policy/mod.rs:186-203
#include "c6_board.cuh"
#include "c6_engine.cuh"
#include "c6_rng.cuh"

namespace c6 {

// Two boards per thread (boards i and i + n/2: both request streams stay coalesced, and the two independent shift/AND chains
// interleave): 64 B in per board (two 256-bit loads), 1 B out.  65 algorithmic bytes per board.
//
// What bounds it (profiles/r02o_prof_secondary.md): not HBM and not issue slots (63 % busy) but the ALU pipe (76 % busy,
// `math_pipe_throttle` the top stall): every LOP3 and every funnel shift (SHF) of the 256-bit run detection is an ALU-pipe
// instruction and that pipe takes one warp instruction every other cycle.  c6_board.cuh's run6_all keeps it to 168 of them per
// board (2 LOP3 per word, stride and colour; the plain in-word shifts of its row-view formulation issue on the FMA pipe as
// IMAD.SHL / IMAD.HI): 0.73 of the copy bandwidth, from 0.56 with 272 (pair ladder, funnel shifts everywhere).  Doing ALL
// shifts as 32 x 32 -> 64-bit multiplies (r02f A/B) was slower: IMAD.WIDE is half rate.
__device__ __forceinline__ void load_board(const uint32_t* p, uint32_t (&B)[8], uint32_t (&W)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(B[0]), "=r"(B[1]), "=r"(B[2]), "=r"(B[3]), "=r"(B[4]), "=r"(B[5]), "=r"(B[6]), "=r"(B[7])
                 : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(W[0]), "=r"(W[1]), "=r"(W[2]), "=r"(W[3]), "=r"(W[4]), "=r"(W[5]), "=r"(W[6]), "=r"(W[7])
                 : "l"(p + 8));
}
__global__ void __launch_bounds__(256) k_win_check(const uint32_t* __restrict__ boards, int8_t* __restrict__ winners,
                                                   long long n) {
    const long long half = (n + 1) / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    const long long i2 = i + half;
    uint32_t B0[8], W0[8], B1[8], W1[8];
    load_board(boards + i * 16, B0, W0);
    if (i2 < n) load_board(boards + i2 * 16, B1, W1);
    winners[i] = (int8_t)win_check_fast(B0, W0);
    if (i2 < n) winners[i2] = (int8_t)win_check_fast(B1, W1);
}

// Synthetic boards: every cell iid uniform over {Black, None, White} (the reference's own random-board
// helper, LC/policy/alphazero_policy/augment_tests.rs:126-129), Philox keyed (seed^BRDS, 0),
// counter (board_lo, board_hi, block, 0), four cells per block.
__global__ void __launch_bounds__(256) k_random_boards(uint32_t* __restrict__ boards, long long n, uint32_t seed) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0}, W[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int blk = 0; blk < 57; ++blk) {
        Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), (uint32_t)blk, 0,
                                  seed ^ DOMAIN_BOARDS, 0);
        uint32_t v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int cell = blk * 4 + t;
            if (cell < CELLS) {
                uint32_t pick = (uint32_t)(((unsigned long long)v[t] * 3ull) >> 32);
                uint32_t bit = bit_of_cell(cell);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t m = ((int)(bit >> 5) == j) ? (1u << (bit & 31)) : 0u;
                    if (pick == 0) B[j] |= m;
                    if (pick == 2) W[j] |= m;
                }
            }
        }
    }
    uint4* out = reinterpret_cast<uint4*>(boards + i * 16);
    out[0] = make_uint4(B[0], B[1], B[2], B[3]);
    out[1] = make_uint4(B[4], B[5], B[6], B[7]);
    out[2] = make_uint4(W[0], W[1], W[2], W[3]);
    out[3] = make_uint4(W[4], W[5], W[6], W[7]);
}

// One playout per thread, board in registers.  Stone n takes the ((u32 * remaining) >> 32)-th empty
// cell in row-major order (IteratorRandom-free twin of `possible.choose`, default_policy/mod.rs:190-199),
// u32 = lane n%4 of Philox(counter (playout_lo, playout_hi, n/4, 0), key (seed^PLAY, 0)).
// Only the colour that just moved can newly own a six, so checking that colour after every stone
// is search() exactly.
__global__ void __launch_bounds__(128) k_playouts(long long first, long long n, uint32_t seed, int8_t* __restrict__ winner,
                                                  uint8_t* __restrict__ stones, uint32_t* __restrict__ boards) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long id = (unsigned long long)(first + i);
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0}, W[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int s = 0, win = 0;
    Philox4 r{0, 0, 0, 0};
    while (s < CELLS) {
        if ((s & 3) == 0) r = philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)(s >> 2), 0, seed ^ DOMAIN_PLAYOUT, 0);
        uint32_t u = (s & 3) == 0 ? r.x : (s & 3) == 1 ? r.y : (s & 3) == 2 ? r.z : r.w;
        uint32_t pick = (uint32_t)(((unsigned long long)u * (unsigned long long)(CELLS - s)) >> 32);
        const uint32_t bit = pick_empty_cell(B, W, pick);
        const int turn = turn_of(s);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t m = ((int)(bit >> 5) == j) ? (1u << (bit & 31)) : 0u;
            if (turn < 0) B[j] |= m; else W[j] |= m;
        }
        s += 1;
        uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (turn < 0) {
            run6_all(B, acc);
        } else {
            run6_all(W, acc);
        }
        uint32_t any = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) any |= acc[j];
        if (any) {
            win = turn;
            break;
        }
    }
    winner[i] = (int8_t)win;
    stones[i] = (uint8_t)s;
    if (boards) {
        uint4* out = reinterpret_cast<uint4*>(boards + i * 16);
        out[0] = make_uint4(B[0], B[1], B[2], B[3]);
        out[1] = make_uint4(B[4], B[5], B[6], B[7]);
        out[2] = make_uint4(W[0], W[1], W[2], W[3]);
        out[3] = make_uint4(W[4], W[5], W[6], W[7]);
    }
}

// Parity probes for the noise sampler (tests only): log2 of Gamma draw (7, 9, k = i) of the stream keyed (k0, k1) exactly as
// a select step computes it (hot branch of trial 0 first, then the full loop), the f32 log2 / exp2 and the f64 log / exp on
// given inputs, so that every bit can be compared with the oracle.
__global__ void k_probe_gamma(GammaParams gp, uint32_t k0, uint32_t k1, float* out, int n) {
    __shared__ MathTab mt;
    load_math_tab(mt);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = (uint32_t)i;
    float l;
    if (gp.shape < 1.0) {
        if (!gamma_log2_fast(mt, gp, gamma_word0(7u, 9u, k, k0, k1), l)) l = gamma_log2_full(mt, gp.b, gp.inv_a, gp.c1, gp.am1, 7u, 9u, k, k0, k1);
    } else {
        l = gamma_large_log2(gp.shape, 7u, 9u, k, k0, k1);
    }
    out[i] = l;
}
__global__ void k_probe_math(const double* xs, double* logs, double* exps, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    logs[i] = xs[i] > 0.0 ? det_log(xs[i]) : 0.0;
    exps[i] = det_exp(xs[i]);
}
__global__ void k_probe_math32(const float* xs, float* lg, float* ex, int n) {
    __shared__ MathTab mt;
    load_math_tab(mt);
    __syncthreads();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lg[i] = xs[i] > 0.f ? lg2_det(mt, xs[i]) : 0.f;
    ex[i] = ex2_det(xs[i]);
}
__global__ void k_probe_tie(uint32_t seed, uint32_t game, uint32_t count, uint32_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = tie_pick((uint32_t)i, (uint32_t)(i % 7), seed, game, count);
}
void launch_probe_gamma(double shape, uint32_t k0, uint32_t k1, float* d_out, int n, cudaStream_t s) {
    k_probe_gamma<<<(n + 127) / 128, 128, 0, s>>>(gamma_params(shape), k0, k1, d_out, n);
}
void launch_probe_math(const double* d_xs, double* d_logs, double* d_exps, int n, cudaStream_t s) {
    k_probe_math<<<(n + 127) / 128, 128, 0, s>>>(d_xs, d_logs, d_exps, n);
}
void launch_probe_math32(const float* d_xs, float* d_lg, float* d_ex, int n, cudaStream_t s) {
    k_probe_math32<<<(n + 127) / 128, 128, 0, s>>>(d_xs, d_lg, d_ex, n);
}
void launch_probe_tie(uint32_t seed, uint32_t game, uint32_t count, uint32_t* d_out, int n, cudaStream_t s) {
    k_probe_tie<<<(n + 127) / 128, 128, 0, s>>>(seed, game, count, d_out, n);
}

void launch_win_check(const uint32_t* d_boards, int8_t* d_winners, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    k_win_check<<<(unsigned)(((n + 1) / 2 + 255) / 256), 256, 0, s>>>(d_boards, d_winners, (long long)n);
}
void launch_random_boards(uint32_t* d_boards, int64_t n, uint32_t seed, cudaStream_t s) {
    if (n <= 0) return;
    k_random_boards<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_boards, (long long)n, seed);
}
void launch_playouts(int64_t first, int64_t n, uint32_t seed, int8_t* d_winner, uint8_t* d_stones, uint32_t* d_boards,
                     cudaStream_t s) {
    if (n <= 0) return;
    k_playouts<<<(unsigned)((n + 127) / 128), 128, 0, s>>>((long long)first, (long long)n, seed, d_winner, d_stones, d_boards);
}

}  // namespace c6
