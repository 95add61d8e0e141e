// Replay-buffer feed: packed game records -> training tuples in HBM.
//
// The consumers right behind self-play store one (winner, player, board BEFORE the stone, position) sample per stone:
//   AlphaZero/py_weighted/buffer.py:22-31        Buffer.push_game
//   AlphaZero/cpp_weighted/ReplayBuffer.hpp:22-65 push_game / sample (float32 tensors of winners, players, boards[225], poses)
// through 912-byte Path objects (and the C++ wrapper's Path copy keeps only 225 BYTES of the board, connect6.hpp:133,141).
// Here the 240-byte records stay packed until the trainer wants tensors; one warp per game replays its stones and writes the
// rows straight into the trainer's float32 tensors: 900 B of coalesced stores per sample, nothing else.
#include <cstdint>
#include <cuda_runtime.h>

#include "c6_board.cuh"
#include "c6_engine.cuh"

namespace c6 {

// row_offset[g] = number of samples of the games before g (exclusive prefix sum of n_moves); one warp per game.
__global__ void __launch_bounds__(128) k_expand_records(const GameRecord* rec, const long long* row_offset, int n_games, float* winners,
                                                        float* players, float* boards, float* positions) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_games) return;
    const GameRecord& r = rec[g];
    const int m = r.n_moves;
    const long long row0 = row_offset[g];
    const float win = (float)r.winner;
    float cell[8];  // cells lane, lane + 32, ... of the board before the current stone
#pragma unroll
    for (int j = 0; j < 8; ++j) cell[j] = 0.f;
    for (int s = 0; s < m; ++s) {
        float* out = boards + (row0 + s) * CELLS;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (32 * j + lane < CELLS) out[32 * j + lane] = cell[j];
        const int pos = r.moves[s];
        const float turn = (float)turn_of(s);
        if (lane == 0) winners[row0 + s] = win, players[row0 + s] = turn, positions[row0 + s] = (float)pos;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (32 * j + lane == pos) cell[j] = turn;
    }
}

void launch_expand_records(const GameRecord* d_rec, const long long* d_row_offset, int n_games, float* d_winners, float* d_players,
                           float* d_boards, float* d_positions, cudaStream_t s) {
    if (n_games <= 0) return;
    k_expand_records<<<(n_games * 32 + 127) / 128, 128, 0, s>>>(d_rec, d_row_offset, n_games, d_winners, d_players, d_boards, d_positions);
}

}  // namespace c6
