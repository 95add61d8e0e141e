// ORACLE — TEST INFRASTRUCTURE ONLY (see c6_rng.hpp header note).
// C interface over c6_oracle.hpp for ctypes (tests/, bench.py cpu_baseline / --impl reference,
// __graft_entry__.smoke()).  Nothing here is linked into the shipped library.
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>

#include "c6_oracle.hpp"
#include "c6_default_policy.hpp"

using namespace c6o;

namespace {
Board board_from_i8(const int8_t* b) {
    Board out;
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c) out[r][c] = player_from(b[r * BOARD_SIZE + c]);
    return out;
}
}  // namespace

extern "C" {

// ---- deterministic evaluators usable as the reference `Callback` (connect6.hpp:14) -------------
// SURVEY B.4.  Bit-identical to the engine's device twin (csrc/c6_rng.cuh: det_eval_*).
static inline uint64_t oriented_key(const float* board) {
    const Zobrist& z = zobrist();
    uint64_t k = 0;
    for (int i = 0; i < BOARD_CAPACITY; ++i) {
        if (board[i] < 0.f) k ^= z.z[0][i];
        if (board[i] > 0.f) k ^= z.z[1][i];
    }
    return k;
}
void det_eval_callback(int player, float* values, float* policies, int len) {
    for (int n = 0; n < len; ++n) {
        float* b = policies + (size_t)n * BOARD_CAPACITY;
        uint64_t h = splitmix64(oriented_key(b) ^ (player < 0 ? 0xB1ACull : 0x3417Eull));
        values[n] = (float)(uint32_t)(h >> 40) * (1.0f / 8388608.0f) - 1.0f;
        for (int i = 0; i < BOARD_CAPACITY; ++i)
            b[i] = (float)(uint32_t)(splitmix64(h + (uint64_t)i + 1) >> 40) * (1.0f / 16777216.0f);
    }
}
// priors in [-1,1) like RandomEvaluator (negative priors exercise the sqrt(S-n) quirk)
void neg_eval_callback(int player, float* values, float* policies, int len) {
    for (int n = 0; n < len; ++n) {
        float* b = policies + (size_t)n * BOARD_CAPACITY;
        uint64_t h = splitmix64(oriented_key(b) ^ (player < 0 ? 0xB1ACull : 0x3417Eull));
        values[n] = (float)(uint32_t)(h >> 40) * (1.0f / 8388608.0f) - 1.0f;
        for (int i = 0; i < BOARD_CAPACITY; ++i)
            b[i] = (float)(uint32_t)(splitmix64(h + (uint64_t)i + 1) >> 40) * (1.0f / 8388608.0f) - 1.0f;
    }
}
// a peaked evaluator (what a trained network looks like to the node pool): almost all of the prior on one empty
// cell chosen by the board hash, so that most of a stone's simulations go down one line and a large subtree survives the stone
void peaked_eval_callback(int player, float* values, float* policies, int len) {
    for (int n = 0; n < len; ++n) {
        float* b = policies + (size_t)n * BOARD_CAPACITY;
        const uint64_t h = splitmix64(oriented_key(b) ^ (player < 0 ? 0xB1ACull : 0x3417Eull));
        int empties[BOARD_CAPACITY], ne = 0;
        for (int i = 0; i < BOARD_CAPACITY; ++i)
            if (b[i] == 0.f) empties[ne++] = i;
        const int hot = ne ? empties[h % (uint64_t)ne] : 0;
        values[n] = (float)(uint32_t)(h >> 40) * (1.0f / 8388608.0f) * 0.1f - 0.05f;
        for (int i = 0; i < BOARD_CAPACITY; ++i) b[i] = i == hot ? 0.9f : 0.1f / 224.f;
    }
}
// SURVEY B.1 hand-derived trace evaluator
void const_eval_callback(int, float* values, float* policies, int len) {
    for (int n = 0; n < len; ++n) values[n] = 0.5f;
    for (size_t i = 0; i < (size_t)len * BOARD_CAPACITY; ++i) policies[i] = 0.25f;
}

// the RandomEvaluator twin on one leaf: values[8], policies[8][225] in the API layout (planes of augment_way8)
void oracle_random_evaluator(uint32_t seed, uint32_t game_id, uint32_t sim_index, const int8_t* board, float* values, float* policies) {
    PhiloxRandomEvaluator ev;
    EvalCtx ctx;
    ctx.seed = seed, ctx.game_id = game_id, ctx.sim_index = sim_index;
    const Board b = board_from_i8(board);
    std::vector<float> v;
    std::vector<FBoard> p;
    ev.eval(Player::Black, augment_way8(b), ctx, v, p);
    for (int s = 0; s < 8; ++s) {
        values[s] = v[s];
        std::memcpy(policies + (size_t)s * BOARD_CAPACITY, &p[s][0][0], sizeof(FBoard));
    }
}

// ---- small pieces -------------------------------------------------------------------------------
int oracle_search(const int8_t* board) { return (int)search(board_from_i8(board)); }

void oracle_augment_way8(const int8_t* board, int8_t* out) {
    auto v = augment_way8(board_from_i8(board));
    for (int s = 0; s < 8; ++s)
        for (int r = 0; r < BOARD_SIZE; ++r)
            for (int c = 0; c < BOARD_SIZE; ++c) out[s * BOARD_CAPACITY + r * BOARD_SIZE + c] = (int8_t)v[s][r][c];
}
void oracle_recover_way8(const float* in, float* out) {
    std::vector<FBoard> p(8);
    for (int s = 0; s < 8; ++s) std::memcpy(&p[s][0][0], in + s * BOARD_CAPACITY, sizeof(FBoard));
    FBoard r = recover_way8(p);
    std::memcpy(out, &r[0][0], sizeof(FBoard));
}
// kind: 0 rotate_left, 1 rotate_right, 2 flip_vertical, 3 flip_horizontal (augment.rs:8-51)
void oracle_transform(int kind, const int8_t* board, int8_t* out) {
    Board b = board_from_i8(board);
    if (kind == 0) rotate_left(b);
    else if (kind == 1) rotate_right(b);
    else if (kind == 2) flip_vertical(b);
    else flip_horizontal(b);
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c) out[r * BOARD_SIZE + c] = (int8_t)b[r][c];
}
// Game::set replay (game_impl/mod.rs:148-169): per position code 0 ok / 1 "invalid position" /
// 2 "already set position", the SetResult fields, and the game's turn / num_remain after the call.
void oracle_game_replay(const int* pos, int n, int* code, int* player, int* remain, int* turn_after, int* remain_after) {
    Game g;
    for (int i = 0; i < n; ++i) {
        SetResult sr;
        std::string err = g.set(Pos(pos[2 * i], pos[2 * i + 1]), &sr);
        code[i] = err.empty() ? 0 : (err == "invalid position" ? 1 : (err == "already set position" ? 2 : 3));
        player[i] = (int)sr.player, remain[i] = sr.num_remain;
        turn_after[i] = (int)g.turn, remain_after[i] = g.num_remain;
    }
}
uint64_t oracle_zobrist(int white, int cell) { return zobrist().z[white ? 1 : 0][cell]; }
void oracle_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    Philox4 r = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    std::memcpy(out, r.v, 16);
}
double oracle_det_log(double x) { return det_log(x); }
double oracle_det_exp(double x) { return det_exp(x); }
double oracle_gamma(uint32_t a, uint32_t b, uint32_t c, uint32_t k0, uint32_t k1, double shape) {
    return gamma_sample(shape, a, b, c, k0, k1);
}
// the search's sampler: log2 of Gamma draw (a, b, k) in f32; f32 log2 / exp2
float oracle_gamma_log2(uint32_t a, uint32_t b, uint32_t k, uint32_t k0, uint32_t k1, double shape) {
    return gamma_log2_sample(shape, a, b, k, k0, k1);
}
void oracle_gamma_log2_many(uint32_t a, uint32_t b, uint32_t k_first, int n, uint32_t k0, uint32_t k1, double shape, float* out) {
    for (int i = 0; i < n; ++i) out[i] = gamma_log2_sample(shape, a, b, k_first + (uint32_t)i, k0, k1);
}
float oracle_lg2(float x) { return lg2_det(x); }
float oracle_ex2(float x) { return ex2_det(x); }
void oracle_math32_many(const float* xs, int n, float* lg, float* ex) {
    for (int i = 0; i < n; ++i) lg[i] = xs[i] > 0.f ? lg2_det(xs[i]) : 0.f, ex[i] = ex2_det(xs[i]);
}
uint32_t oracle_tie_pick(uint32_t sim, uint32_t depth, uint32_t seed, uint32_t game, uint32_t n) { return tie_pick(sim, depth, seed, game, n); }

// Random playout (default_policy/mod.rs:186-203 rollout loop) from the empty board with the
// engine's Philox stream: stone n of playout p picks index (u32 * remaining) >> 32 of the
// row-major list of empty cells; counter (p_lo, p_hi, n/4, 0) lane n%4, key (seed^PLAY, 0).
void oracle_playout(uint32_t seed, uint64_t playout, int* winner, int* stones, int8_t* final_board) {
    Simulate sim;
    sim.possible = Simulate::possible_from(sim.board);
    int n = 0;
    Philox4 blk{};
    while (search(sim.board) == Player::None && !sim.possible.empty()) {
        if ((n & 3) == 0)
            blk = philox4x32_10((uint32_t)playout, (uint32_t)(playout >> 32), (uint32_t)(n >> 2), 0,
                                seed ^ DOMAIN_PLAYOUT, 0);
        uint32_t u = blk.v[n & 3];
        size_t pick = (size_t)(((uint64_t)u * (uint64_t)sim.possible.size()) >> 32);
        // `possible` stays in row-major order: simulate_in removes order-preservingly and the
        // rollout never rolls back (default_policy/mod.rs:190-199), so this is the k-th empty cell.
        Pos p = sim.possible[pick];
        sim.simulate_in(p.first, p.second);
        ++n;
    }
    *winner = (int)search(sim.board);
    *stones = n;
    if (final_board)
        for (int r = 0; r < BOARD_SIZE; ++r)
            for (int c = 0; c < BOARD_SIZE; ++c) final_board[r * BOARD_SIZE + c] = (int8_t)sim.board[r][c];
}

// ---- scripted game (agent_impl/tests.rs TestPolicy) ------------------------------------------
struct ScriptPolicy : Policy {
    const int* moves;
    int n, i = 0;
    ScriptPolicy(const int* m, int n_) : moves(m), n(n_) {}
    std::optional<Pos> next(const Game&) override {
        if (i >= n) return std::nullopt;
        Pos p(moves[2 * i], moves[2 * i + 1]);
        ++i;
        return p;
    }
};
// returns path length, or -1 on Err (err_out gets the description); paths: per entry turn + 225 cells + row + col (ints)
int oracle_script_game(const int* moves, int n, int* winner, int* paths, char* err_out) {
    ScriptPolicy pol(moves, n);
    PlayResult r = agent_play(pol);
    if (!r.error.empty()) {
        std::strncpy(err_out, r.error.c_str(), 63);
        return -1;
    }
    *winner = (int)r.winner;
    for (size_t k = 0; k < r.path.size(); ++k) {
        int* p = paths + k * (BOARD_CAPACITY + 3);
        p[0] = (int)r.path[k].turn;
        for (int i = 0; i < BOARD_CAPACITY; ++i) p[1 + i] = (int)r.path[k].board[i / BOARD_SIZE][i % BOARD_SIZE];
        p[1 + BOARD_CAPACITY] = r.path[k].pos.first;
        p[2 + BOARD_CAPACITY] = r.path[k].pos.second;
    }
    return (int)r.path.size();
}

// ---- self-play ----------------------------------------------------------------------------------
struct OracleParams {
    int32_t num_simulation;
    float epsilon;
    double dirichlet_alpha;
    float c_puct;
    uint32_t seed;
    int32_t noise_mode;  // 0 none (epsilon must be 0), 1 philox
    int32_t tie_mode;    // 0 lowest index, 1 random (splitmix), 2 Philox keyed (sim, depth): the engine default
    int32_t faithful;    // 0 lean, 1 reference cost model
    int32_t evaluator;   // 0 callback, 1 philox random evaluator
    int32_t max_moves;   // stop after this many stones (<=0: play to the end)
    int32_t virtual_loss;  // <= 1: the reference's sequential simulations; K > 1: waves of up to K leaves (AlphaZero::search_wave)
};

struct OracleGameOut {
    int32_t winner;
    int32_t n_moves;
    int32_t finished;     // 0 when cut by max_moves
    int32_t moves[BOARD_CAPACITY];             // row*15+col
    int32_t turns[BOARD_CAPACITY];             // -1/+1 side that placed the stone
    uint64_t sims, sum_ck, sum_cleaf, sum_depth, evals, max_nodes;
};

static void play_one(Callback cb, const OracleParams& p, uint32_t game_id, OracleGameOut* out,
                     int32_t* child_visits, int32_t* root_visit, float* root_qsum, const int* prefix = nullptr, int n_prefix = 0) {
    HyperParameter hp;
    hp.num_simulation = p.num_simulation, hp.epsilon = p.epsilon, hp.dirichlet_alpha = p.dirichlet_alpha,
    hp.c_puct = p.c_puct;
    std::unique_ptr<Evaluator> ev;
    if (p.evaluator == 1)
        ev.reset(new PhiloxRandomEvaluator());
    else
        ev.reset(new CppEval(cb));
    AlphaZeroPolicy pol(ev.get(), hp);
    pol.az.tie = p.tie_mode == 2 ? TieBreak::Philox : (p.tie_mode ? TieBreak::Random : TieBreak::LowestIndex);
    pol.az.noise = p.noise_mode ? NoiseMode::Philox : NoiseMode::None;
    pol.az.faithful = p.faithful != 0;
    pol.az.ctx.seed = p.seed;
    pol.az.ctx.game_id = game_id;
    pol.az.virtual_loss = p.virtual_loss;
    pol.trace = child_visits != nullptr;

    // Agent::play with an optional stone limit (bounded samples for timing)
    Game game;
    out->winner = 0, out->n_moves = 0, out->finished = 1;
    // stones placed by somebody else before the policy is asked for the first time (must not end the game)
    for (int i = 0; i < n_prefix; ++i) {
        out->moves[out->n_moves] = prefix[i];
        out->turns[out->n_moves] = (int)game.turn;
        out->n_moves += 1;
        if (!game.set(Pos(prefix[i] / BOARD_SIZE, prefix[i] % BOARD_SIZE)).empty() || game.is_game_end() != Player::None) std::abort();
    }
    for (;;) {
        if (p.max_moves > 0 && out->n_moves >= p.max_moves) {
            out->finished = 0;
            break;
        }
        auto pos = pol.next(game);
        if (!pos) break;
        out->moves[out->n_moves] = pos->first * BOARD_SIZE + pos->second;
        out->turns[out->n_moves] = (int)game.turn;
        out->n_moves += 1;
        std::string err = game.set(*pos);
        if (!err.empty()) std::abort();
        Player e = game.is_game_end();
        if (e != Player::None) {
            out->winner = (int)e;
            break;
        }
    }
    const SearchStats& s = pol.az.stats;
    out->sims = s.sims, out->sum_ck = s.sum_ck, out->sum_cleaf = s.sum_cleaf, out->sum_depth = s.sum_depth,
    out->evals = s.evals, out->max_nodes = s.max_nodes;
    if (child_visits) {
        // row = stones on the board when next() was called; the call on a full board (225 stones, returns None) has no row
        for (size_t m = 0; m < pol.child_visits.size() && m + n_prefix < (size_t)BOARD_CAPACITY; ++m) {
            std::memcpy(child_visits + (m + n_prefix) * BOARD_CAPACITY, pol.child_visits[m].data(), BOARD_CAPACITY * 4);
            root_visit[m + n_prefix] = pol.root_visit[m];
            root_qsum[m + n_prefix] = pol.root_qsum[m];
        }
    }
}

// One game.  child_visits: [225][225] i32 (row m = the root's children before stone m), may be NULL.
void oracle_self_play(Callback cb, const OracleParams* p, uint32_t game_id, OracleGameOut* out,
                      int32_t* child_visits, int32_t* root_visit, float* root_qsum) {
    play_one(cb, *p, game_id, out, child_visits, root_visit, root_qsum);
}

// The same from a given position: `prefix` = cells (row*15+col) of the stones already played, in order.
void oracle_self_play_from(Callback cb, const OracleParams* p, uint32_t game_id, const int* prefix, int n_prefix, OracleGameOut* out,
                           int32_t* child_visits, int32_t* root_visit, float* root_qsum) {
    play_one(cb, *p, game_id, out, child_visits, root_visit, root_qsum, prefix, n_prefix);
}

// n_games games on n_threads OS threads, one game per task like AsyncAgent::run
// (agent/async_agent/mod.rs:100-137).  Returns wall seconds; outs[n_games].
double oracle_self_play_many(Callback cb, const OracleParams* p, uint32_t first_game_id, int n_games,
                             int n_threads, OracleGameOut* outs) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::atomic<int> nextg{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            for (;;) {
                int g = nextg.fetch_add(1);
                if (g >= n_games) break;
                play_one(cb, *p, first_game_id + (uint32_t)g, &outs[g], nullptr, nullptr, nullptr);
            }
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// MultiPolicy(AlphaZero as Black, scripted White) under Agent::play — the shape of cpp_play_with
// (cppbind/libcppconnect6/src/lib.rs:140-169, policy/multi_policy/mod.rs:61-71) with the stdin policy
// replaced by a move list.  White running out of moves ends the game (next() -> None).
void oracle_play_with(Callback cb, const OracleParams* p, uint32_t game_id, const int* white_moves, int n_white,
                      OracleGameOut* out) {
    HyperParameter hp;
    hp.num_simulation = p->num_simulation, hp.epsilon = p->epsilon, hp.dirichlet_alpha = p->dirichlet_alpha, hp.c_puct = p->c_puct;
    std::unique_ptr<Evaluator> ev;
    if (p->evaluator == 1)
        ev.reset(new PhiloxRandomEvaluator());
    else
        ev.reset(new CppEval(cb));
    AlphaZeroPolicy black(ev.get(), hp);
    black.az.tie = p->tie_mode == 2 ? TieBreak::Philox : TieBreak::LowestIndex;
    black.az.noise = p->noise_mode ? NoiseMode::Philox : NoiseMode::None;
    black.az.faithful = p->faithful != 0;
    black.az.ctx.seed = p->seed, black.az.ctx.game_id = game_id;
    ScriptPolicy white(white_moves, n_white);
    struct Multi : Policy {
        Policy &b, &w;
        Multi(Policy& b_, Policy& w_) : b(b_), w(w_) {}
        std::optional<Pos> next(const Game& g) override { return g.turn == Player::Black ? b.next(g) : w.next(g); }
    } multi(black, white);
    PlayResult r = agent_play(multi);
    out->winner = (int)r.winner, out->n_moves = (int)r.path.size(), out->finished = r.error.empty() ? 1 : 0;
    for (size_t k = 0; k < r.path.size(); ++k) {
        out->moves[k] = r.path[k].pos.first * BOARD_SIZE + r.path[k].pos.second;
        out->turns[k] = (int)r.path[k].turn;
    }
    const SearchStats& st = black.az.stats;
    out->sims = st.sims, out->sum_ck = st.sum_ck, out->sum_cleaf = st.sum_cleaf, out->sum_depth = st.sum_depth, out->evals = st.evals,
    out->max_nodes = st.max_nodes;
}

// ---- DefaultPolicy (pure MCTS with random rollouts, default_policy/mod.rs) under Agent::play ----------------
struct OracleDpOut {
    int winner, n_moves;
    int moves[BOARD_CAPACITY];
    int32_t root_visit[BOARD_CAPACITY], root_black_win[BOARD_CAPACITY];  // root node after the searches of move m
    uint64_t searches, rollouts, rollout_stones, select_steps, children_scanned, replaced, nodes;
};
static void dp_play_one(int num_iter, uint32_t seed, uint32_t game_id, int max_moves, OracleDpOut* out) {
    DefaultPolicy pol(num_iter, seed, game_id);
    Game game;
    std::memset(out, 0, sizeof *out);
    for (;;) {
        if (max_moves > 0 && out->n_moves >= max_moves) break;
        auto pos = pol.next(game);
        if (!pos) break;
        const DpNode& root = pol.map.at(DefaultPolicy::hash(game.board));
        out->root_visit[out->n_moves] = root.visit, out->root_black_win[out->n_moves] = root.black_win;
        out->moves[out->n_moves++] = pos->first * BOARD_SIZE + pos->second;
        if (!game.set(*pos).empty()) std::abort();
        Player e = game.is_game_end();
        if (e != Player::None) {
            out->winner = (int)e;
            break;
        }
    }
    const DpStats& s = pol.stats;
    out->searches = s.searches, out->rollouts = s.rollouts, out->rollout_stones = s.rollout_stones;
    out->select_steps = s.select_steps, out->children_scanned = s.children_scanned, out->replaced = s.replaced;
    out->nodes = pol.map.size();
}
void oracle_default_policy_play(int num_iter, uint32_t seed, uint32_t game_id, int max_moves, OracleDpOut* out) {
    dp_play_one(num_iter, seed, game_id, max_moves, out);
}
// n_games on n_threads OS threads (one game per task like AsyncAgent::run); returns wall seconds
double oracle_default_policy_many(int num_iter, uint32_t seed, uint32_t first_game_id, int n_games, int max_moves, int n_threads,
                                  OracleDpOut* outs) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::atomic<int> nextg{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            for (;;) {
                int g = nextg.fetch_add(1);
                if (g >= n_games) break;
                dp_play_one(num_iter, seed, first_game_id + (uint32_t)g, max_moves, &outs[g]);
            }
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Batch forms for the BASELINE configs[2] sizes (16M boards, 1M playouts): bit-packed boards in the engine's layout
// (n x 16 u32: words 0-7 black, 8-15 white, bit 16*row+col), unpacked here and fed to the same search() / rollout.
static Board board_from_packed(const uint32_t* w) {
    Board b = empty_board();
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c) {
            const int bit = 16 * r + c;
            if ((w[bit >> 5] >> (bit & 31)) & 1u) b[r][c] = Player::Black;
            if ((w[8 + (bit >> 5)] >> (bit & 31)) & 1u) b[r][c] = Player::White;
        }
    return b;
}
void oracle_search_batch_packed(const uint32_t* boards, int64_t n, int8_t* winners, int n_threads) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([=]() {
            for (int64_t i = t; i < n; i += n_threads) winners[i] = (int8_t)search(board_from_packed(boards + i * 16));
        });
    for (auto& x : th) x.join();
}
void oracle_playout(uint32_t seed, uint64_t playout, int* winner, int* stones, int8_t* final_board);
void oracle_playout_batch(uint32_t seed, uint64_t first, int64_t n, int8_t* winners, uint8_t* stones, uint32_t* boards, int n_threads) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([=]() {
            for (int64_t i = t; i < n; i += n_threads) {
                int w, s;
                int8_t fb[BOARD_CAPACITY];
                oracle_playout(seed, first + (uint64_t)i, &w, &s, fb);
                winners[i] = (int8_t)w, stones[i] = (uint8_t)s;
                if (boards) {
                    uint32_t* o = boards + i * 16;
                    for (int j = 0; j < 16; ++j) o[j] = 0;
                    for (int c = 0; c < BOARD_CAPACITY; ++c) {
                        const int bit = 16 * (c / BOARD_SIZE) + c % BOARD_SIZE;
                        if (fb[c] < 0) o[bit >> 5] |= 1u << (bit & 31);
                        if (fb[c] > 0) o[8 + (bit >> 5)] |= 1u << (bit & 31);
                    }
                }
            }
        });
    for (auto& x : th) x.join();
}

int oracle_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

// 16M-board style batch win check (config 3a): boards as int8 [n][225]
void oracle_search_batch(const int8_t* boards, int64_t n, int8_t* winners, int n_threads) {
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([=]() {
            for (int64_t i = t; i < n; i += n_threads)
                winners[i] = (int8_t)search(board_from_i8(boards + i * BOARD_CAPACITY));
        });
    for (auto& x : th) x.join();
}
}
