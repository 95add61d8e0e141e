/* C-ABI of libcppconnect6 (B200 engine).
 *
 * Part 1 is, symbol for symbol and struct for struct, what the reference's C++ header binds
 * (reference: Connect6/cppbind/headers/connect6.hpp:10-93; Rust side
 * Connect6/cppbind/libcppconnect6/src/lib.rs:19-169, cppbind/rawobj/mod.rs:29-51,
 * cppbind/ffi_test/mod.rs).  The reference's own connect6.hpp compiles and links against this
 * library unchanged.
 *
 * Part 2 (c6_*) is the extended interface: explicit engine handles, device-resident evaluators,
 * packed game records, statistics and the config-3 micro-benchmarks.  Plain pointers and sizes
 * only; no C++ or torch types cross this boundary.
 *
 * There is no CPU fallback: every search runs on the GPU, and calls fail loudly (message on
 * stderr, abort() for the reference entry points — as the reference's .unwrap() does,
 * lib.rs:40,107,168 — or a non-zero return + c6_last_error() for c6_*) when no CUDA device or
 * not enough device memory is available.
 */
#ifndef CONNECT6_B200_H
#define CONNECT6_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* Part 1 — reference ABI                                                                      */
/* ------------------------------------------------------------------------------------------ */
#define C6_BOARD_SIZE 15
#define C6_BOARD_CAPACITY 225

/* connect6.hpp:14 — `policies` is in/out: on entry len x 225 boards (-1 black, 0, +1 white), on
 * exit len x 225 priors; `values` (len, pre-zeroed) out; player = side to move (-1 / +1).       */
typedef void (*c6_callback)(int player, float* values, float* policies, int len);
/* connect6.hpp:15 — position pre-set to {-1,-1}; leaving position[0] == -1 ends the game.       */
typedef void (*c6_policy_callback)(float* board, int* position);

/* connect6.hpp:26-42 (x86-64: 912 B, 24 B, 16 B) */
typedef struct c6_path {
    int turn;
    int board[C6_BOARD_SIZE][C6_BOARD_SIZE];
    int row;
    int col;
} c6_path;
typedef struct c6_play_result {
    int winner;
    c6_path* paths;
    int len;
} c6_play_result;
typedef struct c6_vec {
    c6_play_result* vec;
    int len;
} c6_vec;
typedef struct c6_vec_int {
    int* vec;
    int len;
} c6_vec_int;
typedef struct c6_vec_float {
    float* vec;
    int len;
} c6_vec_float;

/* connect6.hpp:17-23 — results are allocated by the caller's allocator (`new T[n]`), never by us */
typedef c6_path* (*c6_alloc_path)(int size);
typedef c6_play_result* (*c6_alloc_result)(int size);
typedef int* (*c6_alloc_int)(int size);
typedef float* (*c6_alloc_float)(int size);

#ifndef C6_OMIT_REFERENCE_ABI /* define it when the reference's own connect6.hpp (which declares these) is included too */
/* connect6.hpp:44-48 / lib.rs:19-58 — games driven by a user policy callback (host game loop). */
c6_vec cpp_play(c6_policy_callback callback, c6_alloc_path alloc_path, c6_alloc_result alloc_result, bool debug,
                int num_game_thread);
/* connect6.hpp:50-58 / lib.rs:74-126 — AlphaZero self-play; num_game_thread games run concurrently
 * on the GPU, `callback` is invoked from the calling thread at most twice per simulation wave.  */
c6_vec cpp_self_play(c6_callback callback, c6_alloc_path alloc_path, c6_alloc_result alloc_result, int num_simulation,
                     float epsilon, double dirichlet_alpha, float c_puct, bool debug, int num_game_thread);
/* connect6.hpp:60-65 / lib.rs:139-169 — AlphaZero (Black, GPU) against stdin moves "aB" (White). */
c6_play_result cpp_play_with(c6_callback callback, c6_alloc_path alloc_path, int num_simulation, float epsilon,
                             double dirichlet_alpha, float c_puct);

/* connect6.hpp:70-91 / ffi_test/mod.rs — ABI echo helpers used by the reference's Catch2 suite. */
c6_path test_new_raw_path(void);
c6_path test_with_raw_path(void);
c6_path test_echo_raw_path(int turn, int* board, int row, int col);
c6_play_result test_with_raw_play_result(c6_alloc_path allocator);
c6_play_result test_echo_raw_play_result(int winner, c6_path* path, int len, c6_alloc_path allocator);
c6_vec_int test_with_raw_vec(c6_alloc_int allocator);
c6_vec_int test_echo_raw_vec(int* ptr, int len, c6_alloc_int allocator);
c6_vec_float test_echo_cppeval(int turn, int* boards, int len, c6_callback callback, c6_alloc_float allocator);
c6_vec_int test_cpp_policy(float* board, c6_policy_callback callback, c6_alloc_int allocator);
#endif /* C6_OMIT_REFERENCE_ABI */

/* ------------------------------------------------------------------------------------------ */
/* Part 2 — engine interface                                                                   */
/* ------------------------------------------------------------------------------------------ */
enum { C6_EVAL_CALLBACK = 0, C6_EVAL_PHILOX_RANDOM = 1, C6_EVAL_DET_HASH = 2, C6_EVAL_NEG_HASH = 3, C6_EVAL_DEVICE_CALLBACK = 4 };

/* connect6.hpp:14 with DEVICE pointers: the leaf batch stays in HBM (float[len][15][15] boards in `d_policies`,
 * overwritten with the priors; `d_values` pre-zeroed).  `stream` is the cudaStream_t the engine's kernels of this
 * pipeline group are ordered on: enqueue the evaluation there (or synchronise before returning).  This is the
 * zero-copy path for a network that lives on the same GPU (AlphaZero/cpp_weighted/main.cpp:16-37 copies through
 * the host twice per leaf). */
typedef void (*c6_device_callback)(int player, float* d_values, float* d_policies, int len, void* stream, void* user);

/* HyperParameter (alphazero_policy/mod.rs:90-125) + engine knobs */
typedef struct c6_config {
    int32_t num_simulation;  /* 800 */
    float epsilon;           /* 0.25 */
    double dirichlet_alpha;  /* 0.03 */
    float c_puct;            /* 1 */
    uint32_t seed;           /* Philox key of the noise / random evaluator */
    int32_t device;          /* CUDA ordinal */
    int32_t slots;           /* games resident on the GPU at once */
    int32_t log2_records;    /* node-record ring capacity per game (18 = 262,144) */
    int32_t evaluator;       /* C6_EVAL_* */
    int32_t trace;           /* keep root child visits per move */
    int32_t max_moves;       /* stop each game after this many stones (0 = play out) */
    int32_t virtual_loss;    /* 1 = the reference's strictly sequential simulations (default).  K in 2..8, opt-in and NOT the
                                reference's search: waves of up to K descents per game whose visits are applied at once (virtual
                                loss) and whose leaves are evaluated in one batch and expanded afterwards */
    int32_t tie_break;       /* exact PUCT ties: 1 = uniform among the tied children as alphazero_policy/mod.rs:291-295 does
                                (Philox keyed (seed, game), counted by (simulation, depth)); 0 = lowest row-major index */
} c6_config;

/* 240 B per finished game */
typedef struct c6_record {
    int8_t winner;
    uint8_t n_moves;
    uint8_t finished;
    uint8_t pad;
    uint32_t game_id;
    uint8_t moves[232]; /* row*15+col per stone */
} c6_record;

typedef struct c6_stats {
    uint64_t sims, sum_ck, sum_cleaf, sum_depth, evals, gc_copied, games_done;
    uint32_t peak_rec, peak_lnk;
    uint64_t kernel_launches;
    uint64_t bytes_per_game;
    uint64_t degraded; /* searches restarted from a fresh root because a game's node pool was full (summed over resident slots) */
} c6_stats;

typedef struct c6_engine c6_engine;

void c6_default_config(c6_config* cfg);
const char* c6_last_error(void);
int c6_device_count(void);
/* NULL on failure (see c6_last_error) */
c6_engine* c6_engine_create(const c6_config* cfg);
void c6_engine_destroy(c6_engine* e);
/* every function below returns 0 on success */
int c6_engine_start(c6_engine* e, uint32_t first_game_id, int64_t total_games);
int c6_engine_step_device(c6_engine* e);                  /* one move for every resident game, async */
int c6_engine_wave_callback(c6_engine* e, c6_callback cb); /* one simulation for every resident game */
int c6_engine_waves_callback(c6_engine* e, c6_callback cb, int waves); /* `waves` simulations per game, pipelined */
/* a stone chosen outside the engine for the game in `slot` (Game::set rules; the node of the new board is
 * reused if the table has it, as AlphaZero::init does, alphazero_policy/mod.rs:302-306) */
int c6_engine_external_move(c6_engine* e, int slot, int row, int col);
int c6_engine_waves_device_callback(c6_engine* e, c6_device_callback cb, void* user, int waves);
int c6_engine_run_device_callback(c6_engine* e, c6_device_callback cb, void* user); /* until all games are finished */
int c6_engine_run_device(c6_engine* e);                   /* until all games are finished */
int c6_engine_run_callback(c6_engine* e, c6_callback cb);
int c6_engine_sync(c6_engine* e);
/* callable from inside the evaluator callback: c6_engine_run_* return non-zero ("cancelled") at the next wave */
void c6_engine_cancel(c6_engine* e);
int c6_engine_timed_steps(c6_engine* e, int steps, float* milliseconds);
int c6_engine_timed_waves(c6_engine* e, c6_callback cb, int waves, float* milliseconds, double* h2d_bytes,
                          double* d2h_bytes);
int c6_engine_stats(c6_engine* e, c6_stats* out);
/* callback path, cumulative: seconds the calling thread was blocked on the GPU / spent inside the user Callback */
int c6_engine_host_times(c6_engine* e, double* wait_seconds, double* callback_seconds);
/* device path, cumulative over c6_engine_timed_steps: milliseconds inside the fused search kernel (one CUDA event
 * pair per launch on the engine stream) and the number of its launches */
int c6_engine_kernel_times(c6_engine* e, double* sims_kernel_ms, long long* sims_kernel_launches);
int64_t c6_engine_games_done(c6_engine* e);
int c6_engine_records(c6_engine* e, c6_record* out, int64_t n); /* n = total_games */
/* child_visits: total_games x 225 x 225 (i32, -1 = occupied / no entry), root_n / root_w: total_games x 225 */
int c6_engine_trace(c6_engine* e, int32_t* child_visits, int32_t* root_n, float* root_w);
/* the same for one game (index = game id - first_game_id): child_visits 225 x 225, root_n / root_w 225 */
int c6_engine_trace_game(c6_engine* e, int64_t game, int32_t* child_visits, int32_t* root_n, float* root_w);
/* node-pool invariant check of every resident game; returns the number of violations (-1 on error);
 * detail (may be NULL): 16 x {game<<32|code, a, b, c} */
long long c6_engine_validate(c6_engine* e, unsigned long long* detail);
/* raw per-slot state for tests: what = 0 game state, 1 expanded boards, 2 expanded keys, 3 expanded record ids */
long long c6_engine_debug_read(c6_engine* e, int what, int slot, void* out, long long max_bytes);

/* one-shot convenience: n_games of self-play with a device evaluator, packed records out */
int c6_self_play_device(const c6_config* cfg, uint32_t first_game_id, int64_t n_games, c6_record* out, c6_stats* stats);

/* the same, sharded by game id over several GPUs of the box (one host thread + engine each, no exchange) */
int c6_self_play_device_sharded(const c6_config* cfg, const int* devices, int n_devices, uint32_t first_game_id,
                                int64_t n_games, c6_record* out, c6_stats* stats);

/* config 3a: win detection (search(), game/search_winner/mod.rs:201-246) on bit-packed boards.
 * boards: n x 16 u32 (words 0-7 black, 8-15 white; bit 16*row+col). Host pointers. */
int c6_win_check_host(const uint32_t* boards, int8_t* winners, int64_t n, int device);
/* config 3a bench: n random boards generated on the device (cells iid over {B, none, W}), checked
 * `iters` times; returns the average kernel time; optional copy-out for parity checks. */
int c6_bench_win_check(int64_t n, uint32_t seed, int iters, int device, float* ms_per_iter, uint32_t* boards_out,
                       int8_t* winners_out, int64_t n_out);
/* config 3b: random playouts from the empty board (default_policy/mod.rs:186-203 rollout) */
int c6_bench_playouts(int64_t first, int64_t n, uint32_t seed, int iters, int device, float* ms_per_iter,
                      int8_t* winners_out, uint8_t* stones_out, uint32_t* boards_out, int64_t n_out,
                      uint64_t* total_stones);

/* DefaultPolicy — the reference's pure MCTS with random rollouts (policy/default_policy/mod.rs:94-271) — under
 * Agent::play (agent/agent_impl/mod.rs:121-178) for n_games games on the GPU, `num_iter` searches per stone
 * (DefaultPolicy::new() uses 50, mod.rs:104).  One policy instance per game plays both colours, as
 * `Agent::new(&mut policy).play()` does.  Random choices: Philox streams keyed (seed, game id).
 * root_visit / root_black_win (may be NULL): n_games x 225 i32, the root node's counters when stone m was chosen.
 * Tree memory is per RESIDENT game (one thread each, as many as the GPU keeps resident or the free memory allows; C6_DP_SLOTS
 * caps it): a thread whose game is over takes the next game id, so n_games may be far larger; results depend on
 * (seed, game id) only. */
typedef struct c6_dp_stats {
    uint64_t searches, rollouts, rollout_stones, select_steps, children_scanned, replaced, nodes, errors;
    float kernel_ms;
} c6_dp_stats;
int c6_default_policy_play(int num_iter, uint32_t seed, uint32_t first_game_id, int64_t n_games, int max_moves, int device,
                           c6_record* out, int32_t* root_visit, int32_t* root_black_win, c6_dp_stats* stats);

/* Replay-buffer feed (AlphaZero/cpp_weighted/ReplayBuffer.hpp:22-65, py_weighted/buffer.py:22-31): one sample per stone =
 * (winner, player, board BEFORE the stone as 225 floats, position row*15+col), written as the four float32 tensors the
 * reference trainer's sample() builds.  Device form: every pointer is a device pointer, d_row_offset[g] = number of stones of
 * the games before g, `stream` a cudaStream_t (0 = default).  Host form: host pointers, arrays sized for sum(n_moves) rows. */
int c6_expand_records_device(const c6_record* d_records, const long long* d_row_offset, int n_games, float* d_winners, float* d_players,
                             float* d_boards, float* d_positions, void* stream);
int c6_expand_records_host(const c6_record* records, int n_games, int device, float* winners, float* players, float* boards,
                           float* positions, long long* n_rows);

/* parity probes of the Dirichlet-noise sampler (tests): log2 of Gamma(shape) draw (7, 9, k = i) of the stream keyed
 * (k0, k1) exactly as a select step computes it; the engine's deterministic f32 log2 / exp2 and f64 log / exp on xs
 * (logs only where xs > 0); the uniform tie-break pick (sim = i, depth = i % 7) among `count` tied children */
int c6_probe_gamma(double shape, uint32_t k0, uint32_t k1, float* out, int n, int device);
int c6_probe_math32(const float* xs, float* lg2s, float* ex2s, int n, int device);
int c6_probe_math(const double* xs, double* logs, double* exps, int n, int device);
int c6_probe_tie(uint32_t seed, uint32_t game, uint32_t count, uint32_t* out, int n, int device);

/* pinned host<->device copy bandwidth in GB/s: H2D alone, D2H alone, both at once (per direction) */
int c6_probe_pcie(int64_t bytes, int iters, int device, double* h2d_gbs, double* d2h_gbs, double* both_gbs);

/* sample host evaluator (random values/policies), the twin of Connect6/sample/cpp/main.cpp:6-28 (write-only, like it) */
void c6_sample_random_callback(int player, float* values, float* policies, int len);
/* a second sample evaluator that READS every board before overwriting it (value = mean stone colour, priors f(cell)) */
void c6_sample_reading_callback(int player, float* values, float* policies, int len);
/* pin the calling thread (and threads it creates later) to the CPUs of the GPU's NUMA node; returns the node or -1
 * (topology not exposed: nothing changed) */
int c6_bind_host_to_device_numa(int device);
/* host cores this process may keep busy (hardware threads / LOCAL_WORLD_SIZE) and the polling budget (microseconds) a
 * host-side wait gets before it blocks (C6_SPIN_US overrides) */
int c6_host_cores_per_rank(void);
int c6_host_spin_budget_us(void);
/* callback path, cumulative: waits of the calling thread that outlasted the polling budget and slept in the driver */
long long c6_engine_host_blocked_waits(c6_engine* e);
/* Destroyed engines leave their device and pinned blocks (up to C6_ALLOC_CACHE_MB, default 256 MB in total, never a block
 * above half of that) to the next engine of the process, so that a run of short jobs does not pay the driver's
 * allocation round trips each time; this frees them now. */
void c6_release_cached_memory(void);

#ifdef __cplusplus
}
#endif
#endif /* CONNECT6_B200_H */
