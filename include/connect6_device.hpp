// Opt-in extension of the reference's C++ API: AlphaZero self-play with an evaluator that reads the leaf batch where it
// already is — in HBM — instead of through two host copies per wave.
//
//   #include "connect6.hpp"          // the reference's own header, unchanged (Connect6/cppbind/headers/connect6.hpp)
//   #include "connect6_device.hpp"   // this file
//
//   void evaluate(int player, float* d_values, float* d_policies, int len, void* stream, void* user);
//   std::vector<Connect6::GameResult> games = Connect6::self_play(evaluate, user, Connect6::Param().NumGameThread(4096));
//
// Same contract as Connect6::Callback (connect6.hpp:14; what AlphaZero/cpp_weighted/main.cpp:16-37 implements) with DEVICE
// pointers: on entry `d_policies` holds len x 15 x 15 boards (-1 black, 0, +1 white), to be overwritten with the priors;
// `d_values` (len, pre-zeroed) takes the values; `player` is the side to move; `stream` is the cudaStream_t the engine's
// kernels for this batch are ordered on — enqueue the evaluation there or synchronise before returning.  The call arrives
// from the calling thread, at most twice per simulation wave (once per side to move), with 8 x (number of leaves) boards.
// Results come back through the reference's own GameResult / Path classes.
#pragma once
#ifndef C6_OMIT_REFERENCE_ABI
#define C6_OMIT_REFERENCE_ABI  // connect6.hpp has already declared cpp_play / cpp_self_play / cpp_play_with / test_*
#endif
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "connect6_b200.h"

namespace Connect6 {

using DeviceCallback = c6_device_callback;

struct DeviceOptions {
    int device = 0;              // CUDA ordinal
    unsigned seed = 0;           // Philox key of the noise and the tie-break; 0 = from the OS, like the reference's thread_rng
    int max_concurrent = 8192;   // games resident on the GPU at once (more are queued and refill finished slots)
    int log2_records = 18;       // node-record ring capacity per game
};

inline std::vector<GameResult> self_play(DeviceCallback callback, void* user, const Param& param,
                                         const DeviceOptions& opt = DeviceOptions()) {
    namespace FFI = Connect6_RustFFI;
    const int n = param.num_game_thread;
    if (n <= 0) return {};
    c6_config cfg;
    c6_default_config(&cfg);
    cfg.num_simulation = param.num_simulation, cfg.epsilon = param.epsilon, cfg.dirichlet_alpha = param.dirichlet_alpha;
    cfg.c_puct = param.c_puct, cfg.device = opt.device, cfg.log2_records = opt.log2_records;
    cfg.seed = opt.seed ? opt.seed : std::random_device{}();
    cfg.slots = n < opt.max_concurrent ? n : opt.max_concurrent;
    cfg.evaluator = C6_EVAL_DEVICE_CALLBACK;
    c6_engine* engine = c6_engine_create(&cfg);
    if (!engine) throw std::runtime_error(std::string("Connect6::self_play: ") + c6_last_error());
    std::vector<c6_record> records(static_cast<size_t>(n));
    const bool ok = c6_engine_start(engine, 0, n) == 0 && c6_engine_run_device_callback(engine, callback, user) == 0 &&
                    c6_engine_records(engine, records.data(), n) == 0;
    const std::string err = ok ? "" : c6_last_error();
    c6_engine_destroy(engine);
    if (!ok) throw std::runtime_error("Connect6::self_play: " + err);

    // packed records -> the ABI's PlayResult / Path arrays (board BEFORE the stone, agent_impl/mod.rs:139-143), handed to the
    // reference's own GameResult::from_vec, which takes ownership exactly as it does for cpp_self_play
    FFI::Vec vec;
    vec.len = n;
    vec.vec = FFI::allocator<FFI::PlayResult>(n);
    for (int g = 0; g < n; ++g) {
        const c6_record& r = records[static_cast<size_t>(g)];
        FFI::PlayResult& out = vec.vec[g];
        out.winner = r.winner, out.len = r.n_moves, out.paths = FFI::allocator<FFI::Path>(r.n_moves);
        int board[BOARD_SIZE][BOARD_SIZE] = {};
        for (int m = 0; m < r.n_moves; ++m) {
            const int turn = m == 0 ? -1 : ((((m - 1) / 2) % 2 == 0) ? 1 : -1);  // Black one stone, then two each
            FFI::Path& p = out.paths[m];
            p.turn = turn, p.row = r.moves[m] / BOARD_SIZE, p.col = r.moves[m] % BOARD_SIZE;
            std::memcpy(p.board, board, sizeof board);
            board[p.row][p.col] = turn;
        }
    }
    return GameResult::from_vec(vec);
}

}  // namespace Connect6
