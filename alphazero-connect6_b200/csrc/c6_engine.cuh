// Batched Connect6 AlphaZero self-play engine for B200 (sm_100a) — host-visible declarations.
//
// Replaces, for thousands of concurrent games per GPU (one warp per game):
//   AlphaZero::{next,search,select,maximum_from,expand,update,policy}   LC/policy/alphazero_policy/mod.rs:218-489
//   augment_way8 / recover_way8                                          LC/policy/alphazero_policy/augment.rs:60-92
//   Simulate (board + possible list)                                     LC/policy/simulate/mod.rs:93-302
//   Game::set turn rule, search() win detector                           LC/game/game_impl/mod.rs:148-169, search_winner/mod.rs:201-246
//   Agent::play / AsyncAgent::run game loop and records                  LC/agent/agent_impl/mod.rs:121-178, async_agent/mod.rs:100-137
// (LC = /root/reference/Connect6/libconnect6/src).  See DESIGN.md for the HBM layout.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace c6 {

constexpr int BOARD_SIZE = 15;
constexpr int CELLS = 225;
constexpr uint32_t NIL = 0xFFFFFFFFu;  // Rec.link of an unexpanded node
constexpr uint32_t FWD = 0xFFFFFFFEu;  // Rec.link of a record already moved by the collector

// Node record, 16 B (visit N, value sum W, prior P, offset of the child-link run).
// q_value of the reference == W / N whenever N > 0 (alphazero_policy/mod.rs:77-79,386-408) and
// a node has children iff N >= 1 (:353-357), so neither q_value nor next_node.len() is stored.
struct __align__(16) Rec {
    int32_t n;
    float w;
    float p;
    uint32_t link;
};
// Key-table slot, 16 B; valid iff epoch == the game's current epoch (O(1) clear).  The table holds the
// EXPANDED nodes of a game only (one insert per simulation); children are shared through link runs
// (c6_search.cuh: expand_leaf), so there is no per-child board-key index.
struct __align__(16) Slot {
    uint64_t key;
    uint32_t idx;
    uint32_t epoch;
};

constexpr int VL_MAX = 8;  // most leaves one game may have waiting for their evaluation (Config::virtual_loss)

// A leaf waiting for its evaluation: between two launches on the callback paths, within a wave in multi-leaf mode.
struct PendingLeaf {
    uint32_t B[8], W[8];  // bit boards of the leaf
    uint64_t key;
    uint32_t rec;         // its record
    int32_t depth;        // length of the path that reached it (the path itself: Dev::path, one row per waiting leaf)
    int32_t stones;
    uint32_t last_bit;    // bit position of the last stone on the path (valid when depth >= 1)
    uint32_t sim_index;   // simulation index of the descent: keys the evaluator / noise streams of this leaf
    int32_t batch_pos;    // position of the leaf in its group's evaluator batch (callback paths)
};

enum Status : int32_t { ST_IDLE = 0, ST_ACTIVE = 1, ST_DONE = 2, ST_ERROR = 3 };
enum ErrCode : int32_t { ERR_NONE = 0, ERR_REC_FULL = 1, ERR_LNK_FULL = 2, ERR_EXP_FULL = 3 };
enum EvalKind : int32_t { EVAL_CALLBACK = 0, EVAL_PHILOX = 1, EVAL_DET = 2, EVAL_NEG = 3, EVAL_DEVICE_CALLBACK = 4 };

struct GameState {
    uint32_t B[8], W[8];  // root bitboards: bit 16*row+col, bit 15 of every row is a zero guard
    uint64_t key;         // Zobrist key of the root board
    uint32_t root;        // record index of the root
    int32_t stones;
    int32_t status;
    int32_t winner;
    int32_t err;
    uint32_t game_id;
    uint32_t epoch;
    uint32_t rec_head, rec_tail;  // ring counters (positions are taken & mask)
    uint32_t lnk_head, lnk_tail;
    uint32_t exp_count;
    uint32_t sim_index;       // simulations started in this game (Philox counter)
    int32_t sims_this_move;
    int32_t n_moves;
    // leaves waiting for their evaluation (one on the reference's sequential schedule, up to virtual_loss in multi-leaf mode)
    int32_t n_pend;
    int32_t pad0;
    PendingLeaf pend[VL_MAX];
    // bit s: some expanded node with s stones kept by the collector has no expanded parent left (its former
    // parent did not contain the new root); expansions that create children with s stones look those up by key
    uint32_t orph[8];
    // statistics (SURVEY §8d counters)
    unsigned long long sims, sum_ck, sum_cleaf, sum_depth, evals, gc_copied;
    uint32_t peak_rec, peak_lnk;
    uint32_t degraded;   // searches of this slot that ran out of pool space and were restarted from a fresh root
    uint32_t retried;    // the current stone's search has been restarted already
    uint8_t moves[232];
};

// One finished game, packed (SURVEY §7 "legacy record size": expand to 912 B Paths only at the C-ABI edge)
struct GameRecord {
    int8_t winner;
    uint8_t n_moves;
    uint8_t finished;
    uint8_t pad;
    uint32_t game_id;
    uint8_t moves[232];  // row*15+col per stone
};

struct Config {
    int32_t num_simulation = 800;
    float epsilon = 0.25f;
    double dirichlet_alpha = 0.03;
    float c_puct = 1.f;
    uint32_t seed = 0xC6;
    int32_t device = 0;
    int32_t slots = 1;            // concurrent games (G)
    int32_t log2_records = 18;    // node-record ring capacity per game
    int32_t evaluator = EVAL_CALLBACK;
    int32_t trace = 0;            // record root child visits per move (tests)
    int32_t max_moves = 0;        // stop games after this many stones (<=0: play to the end)
    int32_t virtual_loss = 1;     // 1 = the reference's strictly sequential simulations (mod.rs:468-470).  K > 1 (opt-in, NOT the
                                  // reference's search): a move is searched in waves of up to K descents — each applies its visits at
                                  // once (the virtual loss the next descent sees) — whose leaves are evaluated together and expanded
                                  // afterwards; K leaves per game and evaluator call.  Oracle twin: AlphaZero::search_wave.
    int32_t tie_break = 1;        // exact PUCT ties: 1 = uniform among the tied children like the reference (mod.rs:291-295),
                                  // drawn from Philox (seed, game; sim, depth); 0 = lowest row-major index (hand-checkable traces)
};

struct Stats {
    unsigned long long sims, sum_ck, sum_cleaf, sum_depth, evals, gc_copied;
    unsigned long long games_done;
    uint32_t peak_rec, peak_lnk;
    unsigned long long degraded;  // searches restarted from a fresh root because the game's node pool was full
};

using Callback = void (*)(int player, float* values, float* policies, int len);
// The same contract with DEVICE pointers: the leaf batch stays in HBM (API layout float[len][15][15] in
// `d_policies`, overwritten with the priors; `d_values` pre-zeroed), `stream` is the CUDA stream the engine's
// kernels of this group are ordered on — the evaluator enqueues its work there (or synchronises itself).
using DeviceCallback = void (*)(int player, float* d_values, float* d_policies, int len, void* stream, void* user);

struct DeviceArrays;  // opaque (engine .cu)

class Engine {
   public:
    explicit Engine(const Config& cfg);
    ~Engine();
    Engine(const Engine&) = delete;
    Engine& operator=(const Engine&) = delete;

    // Queue `total_games` games with ids first_id .. first_id+total-1; the first min(slots,total) start now.
    void start(uint32_t first_id, int64_t total_games);
    // Device-evaluator path: one move (num_simulation searches + move choice + collection) for every
    // active slot.  Asynchronous on the engine stream.
    void step_device();
    // Host-callback path: one wave = one simulation for every active slot (select kernel -> D2H ->
    // user Callback per side to move -> H2D -> expand kernel).
    void wave_callback(Callback cb);
    // `waves` simulations for every resident game, pipelined over the groups (see c6_engine.cu)
    void waves_callback(Callback cb, int waves);
    // device-pointer evaluator (engine created with EVAL_DEVICE_CALLBACK): no host copies of the batch
    void waves_device_callback(DeviceCallback cb, void* user, int waves);
    void run_device_callback_to_end(DeviceCallback cb, void* user);
    // A stone chosen outside the engine for the game in `slot` (rules as Game::set; throws "already set
    // position" like the reference's error text).  The node of the new board is reused if the table has it.
    void external_move(int slot, int row, int col);
    void sync();
    int64_t games_done();       // synchronises
    bool all_done();            // synchronises
    void run_device_to_end();
    void run_callback_to_end(Callback cb, bool debug);
    // k calls of step_device() bracketed by CUDA events on the engine stream; returns milliseconds.
    float timed_steps_device(int k);
    float timed_waves_callback(Callback cb, int k, double* h2d_bytes, double* d2h_bytes);
    Stats stats();              // synchronises
    // invariant check of every resident game's node pool; returns the number of violations
    // (detail: 16 x {game<<32|code, a, b, c}); synchronises
    long long validate(unsigned long long* detail);
    size_t debug_read(int what, int slot, void* out, size_t max_bytes);
    std::vector<GameRecord> records();   // synchronises; index = game id - first_id
    // trace (cfg.trace): root child visits [game][move][225], root N and W per move
    void trace(std::vector<int32_t>& child_visits, std::vector<int32_t>& root_n, std::vector<float>& root_w);
    // the same for one game (index = game id - first_id): 225 x 225 child visits, 225 root N, 225 root W
    void trace_game(int64_t game, int32_t* child_visits, int32_t* root_n, float* root_w);
    const Config& config() const { return cfg_; }
    cudaStream_t stream() const { return stream_; }
    size_t bytes_per_game() const;
    long long kernel_launches() const { return launches_; }
    double host_wait_seconds() const { return host_wait_s_; }
    double host_callback_seconds() const { return host_callback_s_; }
    long long host_blocked_waits() const { return host_blocked_waits_; }
    int host_spin_us() const { return spin_us_; }
    // timed_steps_device() brackets every launch of the fused search kernel with its own pair of CUDA events on
    // the engine stream: cumulative milliseconds and launch count (bench.py's roofline denominator)
    double sims_kernel_ms() const { return sims_kernel_ms_; }
    long long sims_kernel_launches() const { return sims_kernel_launches_; }
    GameState slot_state(int slot);   // synchronises (debug printing)
    // May be called from inside the evaluator callback (or another thread): the run_*_to_end loops stop
    // at the next wave / poll and throw "cancelled".
    void cancel() { cancel_.store(true); }

   private:
    void free_records();
    void group_issue(int p, size_t len, bool do_back, bool do_front);  // launching thread: every CUDA enqueue of a wave
    size_t group_eval(int p, Callback cb);                             // calling thread: wait for the leaves, run the user Callback(s)
    void waves_callback_threaded(Callback cb, int waves);
    DeviceCallback dcb_ = nullptr;
    void* dcb_user_ = nullptr;
    // Measured on the B200 box (profiles/e2e/e2e7.md): with five driver calls per group and wave the enqueues no
    // longer bound the wave, so the second thread buys nothing (C6_LAUNCHER=1 turns it on for slow hosts); leaf
    // boards stored by the kernel straight into mapped host memory reach ~30 GB/s against ~45 GB/s for the copy
    // engine, so the copy is the default (C6_ZEROCOPY=1 for the store variant).
    bool use_launcher_ = false;
    bool zero_copy_ = false;
    std::atomic<bool> abort_{false};
    Config cfg_;
    DeviceArrays* d_ = nullptr;
    cudaStream_t stream_ = nullptr;
    int64_t total_games_ = 0;
    uint32_t first_id_ = 0;
    long long launches_ = 0;
    double h2d_bytes_ = 0, d2h_bytes_ = 0;
    double sims_kernel_ms_ = 0;
    long long sims_kernel_launches_ = 0;
    std::vector<cudaEvent_t> kev_;  // event pairs of the timed region in flight
    bool time_kernels_ = false;
    double host_wait_s_ = 0, host_callback_s_ = 0;  // callback path: time blocked on the GPU / inside the user Callback
    long long host_blocked_waits_ = 0;              // waits that outlasted the polling budget and slept in the driver
    int spin_us_ = 200;                             // polling budget of a host-side wait (host_spin_budget_us())
    std::atomic<bool> cancel_{false};
};

// Stand-alone kernels of the path (config 3 microbenches), host launchers.
// boards: n x 16 u32 (8 words black, 8 words white); winners: n x i8.
// DefaultPolicy (pure MCTS with random rollouts) self-play, one thread per game (c6_default_policy.cu); returns kernel ms
float default_policy_play(int num_iter, uint32_t seed, uint32_t first_game_id, long long n_games, int max_moves, int device,
                          GameRecord* out, int32_t* root_visit, int32_t* root_bw, unsigned long long* stats);
void launch_win_check(const uint32_t* d_boards, int8_t* d_winners, int64_t n, cudaStream_t s);
void launch_random_boards(uint32_t* d_boards, int64_t n, uint32_t seed, cudaStream_t s);
// playouts: winner i8, stones u8 (and final boards if d_boards != nullptr)
void launch_playouts(int64_t first, int64_t n, uint32_t seed, int8_t* d_winner, uint8_t* d_stones,
                     uint32_t* d_boards, cudaStream_t s);

// replay feed (c6_replay.cu): packed records -> float32 training tensors in HBM, one warp per game
void launch_expand_records(const GameRecord* d_rec, const long long* d_row_offset, int n_games, float* d_winners, float* d_players,
                           float* d_boards, float* d_positions, cudaStream_t s);
void launch_probe_gamma(double shape, uint32_t k0, uint32_t k1, float* d_out, int n, cudaStream_t s);
void launch_probe_math(const double* d_xs, double* d_logs, double* d_exps, int n, cudaStream_t s);
void launch_probe_math32(const float* d_xs, float* d_lg, float* d_ex, int n, cudaStream_t s);
void launch_probe_tie(uint32_t seed, uint32_t game, uint32_t count, uint32_t* d_out, int n, cudaStream_t s);

const char* last_error();
// host cores this process may keep busy (hardware threads / ranks sharing the box) and the polling budget derived from it
int host_cores_per_rank();
int host_spin_budget_us();
// frees the device / pinned blocks kept from destroyed engines for reuse (C6_ALLOC_CACHE_MB, default 256 MB in total)
void release_cached_memory();

}  // namespace c6
