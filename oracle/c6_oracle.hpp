// ORACLE — TEST INFRASTRUCTURE ONLY (see c6_rng.hpp header note).
//
// CPU restatement of the reference's AlphaZero self-play path, following the Rust sources
// function by function.  Paths are relative to /root/reference/Connect6/libconnect6/src/ :
//   game/player/mod.rs            Player
//   game/game_impl/mod.rs         Game (rules, turn order, error strings)
//   game/search_winner/mod.rs     search()  (row-major DP win detector)
//   policy/simulate/mod.rs        Simulate  (board + ordered `possible` list)
//   policy/alphazero_policy/augment.rs   rotate/flip/augment_way8/recover_way8
//   policy/alphazero_policy/mod.rs       Node, HyperParameter, Evaluator, AlphaZero
//   policy/default_policy/mod.rs:55-83   hash, diff_board
//   agent/agent_impl/mod.rs       Agent::play, Path, PlayResult
// The reference cannot be built here (Rust toolchain absent), so parity is pinned by porting the
// reference's own known-answer tests (oracle_tests.cpp) — see DESIGN.md "Oracle".
//
// Two seams replace the reference's `thread_rng()` uses, as SURVEY.md §0.3/§8c define parity:
//   TieBreak::Philox       the reference's uniform pick among exactly equal PUCT scores (mod.rs:291-295) drawn from a
//                          Philox stream keyed (seed, game) and counted by (simulation, depth) — the engine's default;
//   TieBreak::LowestIndex  the first of the tied children in row-major order (kept selectable: hand-derived traces)
//   NoiseMode::Philox      instead of rand's Dirichlet (mod.rs:262,283); NoiseMode::None only
//                          valid when epsilon == 0 (the draw is multiplied by 0: mod.rs:269).
// Two cost models, identical results:
//   faithful = true   whole-board SipHash-1-3 per lookup, full ~1.2 KB nodes, Vec `possible` — the
//                     reference's own data structures; used for CPU-baseline timing.
//   faithful = false  same algorithm with a cheaper 64-bit board key (incremental Zobrist) so that
//                     800-simulation parity cases finish in seconds.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <memory>
#include <optional>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "c6_rng.hpp"

namespace c6o {

// lib.rs:21-26
constexpr int BOARD_SIZE = 15;
constexpr int BOARD_CAPACITY = BOARD_SIZE * BOARD_SIZE;

// game/player/mod.rs:8-12
enum class Player : int8_t { Black = -1, None = 0, White = 1 };
inline Player switch_player(Player p) {  // player/mod.rs:27-33
    return p == Player::Black ? Player::White : (p == Player::White ? Player::Black : Player::None);
}
inline Player player_from(int v) {  // player/mod.rs:52-60
    return v == -1 ? Player::Black : (v == 1 ? Player::White : Player::None);
}

using Board = std::array<std::array<Player, BOARD_SIZE>, BOARD_SIZE>;
using FBoard = std::array<std::array<float, BOARD_SIZE>, BOARD_SIZE>;
using Pos = std::pair<int, int>;

inline Board empty_board() {
    Board b;
    for (auto& r : b) r.fill(Player::None);
    return b;
}

// ------------------------------------------------------------------------------------------
// game/search_winner/mod.rs:201-246 — first cell, in row-major order, at which a run in the
// direction order Right, Down, RightDown, LeftDown reaches >= 6 decides the winner.
// ------------------------------------------------------------------------------------------
struct Cumulative {
    int right = 0, down = 0, right_down = 0, left_down = 0;
};
struct Block {  // search_winner/mod.rs:83-141
    int flag = 0;
    Cumulative mem[2][BOARD_SIZE + 2];
    // as_tuple(): (prev, now) = (mem[flag], mem[1-flag])
    void update_row() {
        flag = 1 - flag;
        for (int i = 0; i < BOARD_SIZE + 2; ++i) mem[1 - flag][i] = Cumulative();
    }
};

inline Player search(const Board& table) {
    Block black, white;
    auto path_iter = [](Block& block, int col0) -> bool {
        int col = col0 + 1;  // one-indexed
        Cumulative* prev = block.mem[block.flag];
        Cumulative* now = block.mem[1 - block.flag];
        // Right
        int updated = now[col - 1].right + 1;
        if (updated >= 6) return true;
        now[col].right = updated;
        // Down
        updated = prev[col].down + 1;
        if (updated >= 6) return true;
        now[col].down = updated;
        // RightDown
        updated = prev[col - 1].right_down + 1;
        if (updated >= 6) return true;
        now[col].right_down = updated;
        // LeftDown
        updated = prev[col + 1].left_down + 1;
        if (updated >= 6) return true;
        now[col].left_down = updated;
        return false;
    };
    for (int row = 0; row < BOARD_SIZE; ++row) {
        black.update_row();
        white.update_row();
        for (int col = 0; col < BOARD_SIZE; ++col) {
            switch (table[row][col]) {
                case Player::None:
                    continue;
                case Player::Black:
                    if (path_iter(black, col)) return Player::Black;
                    break;
                case Player::White:
                    if (path_iter(white, col)) return Player::White;
                    break;
            }
        }
    }
    return Player::None;
}

// ------------------------------------------------------------------------------------------
// game/game_impl/mod.rs:117-253
// ------------------------------------------------------------------------------------------
struct SetResult {
    Player player = Player::None;
    int num_remain = 0;
    Pos position{0, 0};
};

struct Game {
    Player turn = Player::Black;
    int num_remain = 1;
    Board board = empty_board();

    // returns "" on success, else the reference's Error::description() string
    std::string set(Pos pos, SetResult* out = nullptr) {  // game_impl/mod.rs:148-169
        int row = pos.first, col = pos.second;
        if (row < 0 || col < 0 || row >= BOARD_SIZE || col >= BOARD_SIZE) return "invalid position";
        if (board[row][col] != Player::None) return "already set position";
        board[row][col] = turn;
        num_remain -= 1;
        if (out) *out = SetResult{turn, num_remain, pos};
        if (num_remain <= 0) {
            num_remain = 2;
            turn = switch_player(turn);
        }
        return "";
    }
    Player is_game_end() const { return search(board); }

    // game_impl/mod.rs:217-237
    std::string render() const {
        std::string s = "0 A B C D E F G H I J K L M N O\n";
        for (int i = 0; i < BOARD_SIZE; ++i) {
            s.push_back((char)(0x61 + i));
            s.push_back(' ');
            for (int j = 0; j < BOARD_SIZE; ++j)
                s += board[i][j] == Player::Black ? "X " : (board[i][j] == Player::White ? "O " : "_ ");
            s.push_back('\n');
        }
        return s;
    }
};

// ------------------------------------------------------------------------------------------
// policy/simulate/mod.rs:93-302 (value-semantics subset used by AlphaZero: the Rc<RefCell<Node>>
// sharing only matters for `simulate()`'s RAII undo, restated in expand() below).
// ------------------------------------------------------------------------------------------
struct Simulate {
    Player turn = Player::Black;
    int num_remain = 1;
    Board board = empty_board();
    std::vector<Pos> possible;

    static std::vector<Pos> possible_from(const Board& b) {  // simulate/mod.rs:41-68
        std::vector<Pos> v;
        for (int r = 0; r < BOARD_SIZE; ++r)
            for (int c = 0; c < BOARD_SIZE; ++c)
                if (b[r][c] == Player::None) v.emplace_back(r, c);
        return v;
    }
    static Simulate from_game(const Game& g) {  // :117-124
        Simulate s;
        s.turn = g.turn, s.num_remain = g.num_remain, s.board = g.board;
        s.possible = possible_from(g.board);
        return s;
    }
    Simulate deep_clone() const {  // :133-143 — `possible` rebuilt row-major
        Simulate s;
        s.turn = turn, s.num_remain = num_remain, s.board = board;
        s.possible = possible_from(board);
        return s;
    }
    bool validate(int row, int col) const {  // :171-181
        if (row < 0 || col < 0 || row >= BOARD_SIZE || col >= BOARD_SIZE) return false;
        return board[row][col] == Player::None;
    }
    Player next_turn() const { return num_remain <= 1 ? switch_player(turn) : turn; }  // :195-201
    void simulate_in(int row, int col) {  // :262-274 (order-preserving remove)
        auto it = std::find(possible.begin(), possible.end(), Pos(row, col));
        possible.erase(it);
        board[row][col] = turn;
        num_remain -= 1;
        if (num_remain <= 0) {
            num_remain = 2;
            turn = switch_player(turn);
        }
    }
    void rollback_in(int row, int col) {  // :290-301 (push to the END)
        board[row][col] = Player::None;
        possible.emplace_back(row, col);
        num_remain += 1;
        if (num_remain > 2) {
            num_remain = 1;
            turn = switch_player(turn);
        }
    }
};

// ------------------------------------------------------------------------------------------
// policy/alphazero_policy/augment.rs
// ------------------------------------------------------------------------------------------
template <class T>
using GBoard = std::array<std::array<T, BOARD_SIZE>, BOARD_SIZE>;

template <class T>
void rotate_left(GBoard<T>& b) {  // augment.rs:8-16
    GBoard<T> r{};
    for (int i = 0; i < BOARD_SIZE; ++i)
        for (int j = 0; j < BOARD_SIZE; ++j) r[BOARD_SIZE - j - 1][i] = b[i][j];
    b = r;
}
template <class T>
void rotate_right(GBoard<T>& b) {  // augment.rs:19-27
    GBoard<T> r{};
    for (int i = 0; i < BOARD_SIZE; ++i)
        for (int j = 0; j < BOARD_SIZE; ++j) r[j][BOARD_SIZE - i - 1] = b[i][j];
    b = r;
}
template <class T>
void flip_vertical(GBoard<T>& b) {  // augment.rs:30-38  (axis |)
    for (int i = 0; i < BOARD_SIZE; ++i)
        for (int j = 0; j < BOARD_SIZE / 2; ++j) std::swap(b[i][BOARD_SIZE - j - 1], b[i][j]);
}
template <class T>
void flip_horizontal(GBoard<T>& b) {  // augment.rs:41-51 (axis --)
    for (int i = 0; i < BOARD_SIZE; ++i)
        for (int j = 0; j < BOARD_SIZE / 2; ++j) std::swap(b[BOARD_SIZE - j - 1][i], b[j][i]);
}
template <class T>
void sum_board(GBoard<T>& a, const GBoard<T>& b) {  // augment.rs:54-63
    for (int i = 0; i < BOARD_SIZE; ++i)
        for (int j = 0; j < BOARD_SIZE; ++j) a[i][j] = a[i][j] + b[i][j];
}
inline std::vector<Board> augment_way8(const Board& board) {  // augment.rs:66-80
    std::vector<Board> v;
    v.reserve(8);
    Board b = board;
    for (int k = 0; k < 4; ++k) {
        rotate_left(b);
        v.push_back(b);
        Board c = b;
        flip_vertical(c);
        v.push_back(c);
    }
    return v;
}
inline FBoard recover_way8(std::vector<FBoard> probs) {  // augment.rs:83-100
    FBoard total{};
    for (int i = 0; i < 4; ++i) {
        flip_vertical(probs[i * 2 + 1]);
        FBoard flipped = probs[i * 2 + 1];
        sum_board(probs[i * 2], flipped);
        for (int k = 0; k < i + 1; ++k) rotate_right(probs[i * 2]);
        sum_board(total, probs[i * 2]);
    }
    for (auto& r : total)
        for (auto& x : r) x /= 8.f;
    return total;
}

// default_policy/mod.rs:74-83
inline std::optional<Pos> diff_board(const Board& a, const Board& b) {
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c)
            if (a[r][c] != b[r][c]) return Pos(r, c);
    return std::nullopt;
}

// ------------------------------------------------------------------------------------------
// Board keys.  alphazero_policy/mod.rs:456-460 hashes the whole board with std's DefaultHasher
// (SipHash-1-3, zero key; #[derive(Hash)] on a field-less enum writes the discriminant as an
// isize, i.e. 8 bytes per cell — SURVEY A.11(2)).  The value is only an identity key.
// ------------------------------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
inline uint64_t siphash13_board(const Board& board) {
    uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull, v2 = 0x6c7967656e657261ull,
             v3 = 0x7465646279746573ull;
    auto round = [&]() {
        v0 += v1, v1 = rotl64(v1, 13), v1 ^= v0, v0 = rotl64(v0, 32);
        v2 += v3, v3 = rotl64(v3, 16), v3 ^= v2;
        v0 += v3, v3 = rotl64(v3, 21), v3 ^= v0;
        v2 += v1, v1 = rotl64(v1, 17), v1 ^= v2, v2 = rotl64(v2, 32);
    };
    uint64_t n = 0;
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c) {
            uint64_t m = (uint64_t)(int64_t)(int8_t)board[r][c];  // isize discriminant
            v3 ^= m;
            round();
            v0 ^= m;
            n += 8;
        }
    uint64_t tail = (n & 0xff) << 56;
    v3 ^= tail;
    round();
    v0 ^= tail;
    v2 ^= 0xff;
    round(), round(), round();
    return v0 ^ v1 ^ v2 ^ v3;
}

struct Zobrist {
    uint64_t z[2][BOARD_CAPACITY];
    Zobrist() {
        uint64_t s = 0xC0221EC76B200ull;
        for (int p = 0; p < 2; ++p)
            for (int i = 0; i < BOARD_CAPACITY; ++i) {
                s = splitmix64(s);
                z[p][i] = s;
            }
    }
};
inline const Zobrist& zobrist() {
    static Zobrist z;
    return z;
}
inline uint64_t zobrist_board(const Board& b) {
    const Zobrist& z = zobrist();
    uint64_t k = 0;
    for (int r = 0; r < BOARD_SIZE; ++r)
        for (int c = 0; c < BOARD_SIZE; ++c)
            if (b[r][c] != Player::None) k ^= z.z[b[r][c] == Player::White][r * BOARD_SIZE + c];
    return k;
}

// ------------------------------------------------------------------------------------------
// alphazero_policy/mod.rs:25-80
// ------------------------------------------------------------------------------------------
struct Node {
    int32_t visit = 0;
    float value = 0.f, q_sum = 0.f, q_value = 0.f, n_prob = 0.f;
    std::unique_ptr<FBoard> prob;  // [[f32;15];15] in the reference; kept only in faithful mode (write-only there)
    size_t num_player = 0;
    Board board;
    std::vector<uint64_t> next_node;
    void recalc_q() { q_value = q_sum / (float)visit; }  // :77-79
};
inline Node node_new(const Board& b) {  // :43-59
    Node n;
    size_t cnt = 0;
    for (auto& r : b)
        for (auto x : r) cnt += (x != Player::None);
    n.num_player = cnt;
    n.board = b;
    return n;
}
inline Node node_new_with_num(const Board& b, size_t num) {  // :62-74
    Node n;
    n.num_player = num;
    n.board = b;
    return n;
}

struct HyperParameter {  // :90-125
    int32_t num_simulation = 800;
    float epsilon = 0.25f;
    double dirichlet_alpha = 0.03;
    float c_puct = 1.f;
    static HyperParameter light_weight() {
        HyperParameter p;
        p.num_simulation = 2;
        return p;
    }
};

// Context the deterministic random sources are keyed on (not in the reference).
struct EvalCtx {
    uint32_t seed = 0;
    uint32_t game_id = 0;
    uint32_t sim_index = 0;  // simulations started so far in this game (across moves)
};

// :128-134 — returns false for `None`
struct Evaluator {
    virtual ~Evaluator() {}
    virtual bool eval(Player turn, const std::vector<Board>& boards, const EvalCtx& ctx,
                      std::vector<float>& values, std::vector<FBoard>& policies) = 0;
};

// cppbind/libcppconnect6/src/cppbind/cppeval/mod.rs:8-55 + cpp_support.rs:10-18
using Callback = void (*)(int player, float* values, float* policies, int len);
struct CppEval : Evaluator {
    Callback cb;
    explicit CppEval(Callback c) : cb(c) {}
    bool eval(Player turn, const std::vector<Board>& boards, const EvalCtx&, std::vector<float>& values,
              std::vector<FBoard>& policies) override {
        int len = (int)boards.size();
        values.assign(len, 0.f);
        policies.resize(len);
        for (int n = 0; n < len; ++n)
            for (int i = 0; i < BOARD_SIZE; ++i)
                for (int j = 0; j < BOARD_SIZE; ++j) policies[n][i][j] = (float)(int)boards[n][i][j];
        cb((int)turn, values.data(), &policies[0][0][0], len);
        return true;
    }
};

// The engine's device RandomEvaluator twin (replaces thread_rng RandomEvaluator, mod.rs:137-169):
// every value / policy entry is iid U[-1,1) on a 24-bit grid from Philox keyed (seed^EVAL, game): values from blocks
// (sim, 225, half, 0), the eight draws of an EMPTY cell from the blocks of its rank pair (see eval below; occupied cells
// are masked by get_from and get none).  Plane s at oriented cell T_s(cell) holds draw number s of the original-frame
// `cell` (a fixed permutation of iid draws is iid).
struct PhiloxRandomEvaluator : Evaluator {
    bool eval(Player, const std::vector<Board>& boards, const EvalCtx& ctx, std::vector<float>& values,
              std::vector<FBoard>& policies) override;
};

enum class TieBreak { LowestIndex, Random, Philox };
enum class NoiseMode { None, Philox };

struct SearchStats {  // device-counter twins for the roofline's algorithmic-bytes formula
    uint64_t sims = 0, sum_ck = 0, sum_cleaf = 0, sum_depth = 0, evals = 0, max_nodes = 0;
};

// ------------------------------------------------------------------------------------------
// alphazero_policy/mod.rs:186-490
// ------------------------------------------------------------------------------------------
struct AlphaZero {
    std::unordered_map<uint64_t, Node> map;
    HyperParameter param;
    Evaluator* evaluator;
    TieBreak tie = TieBreak::LowestIndex;
    NoiseMode noise = NoiseMode::Philox;
    bool faithful = false;
    EvalCtx ctx;
    SearchStats stats;
    uint64_t tie_rng = 0x1234567ull;  // only for TieBreak::Random
    // NOT in the reference (its simulations are strictly sequential, mod.rs:468-470): with virtual_loss = K > 1 a move is searched
    // in waves of up to K descents whose leaves are expanded afterwards — the engine's opt-in multi-leaf mode (search_wave below)
    int virtual_loss = 1;
    // trace of the last next(): root child visits (row-major cells, -1 where occupied), root stats
    std::array<int32_t, BOARD_CAPACITY> last_child_visits{};
    int32_t last_root_visit = 0;
    float last_root_qsum = 0.f, last_root_value = 0.f;

    AlphaZero(Evaluator* e, HyperParameter p) : param(p), evaluator(e) {}

    uint64_t hash(const Board& b) const { return faithful ? siphash13_board(b) : zobrist_board(b); }

    // :218-234
    bool get_from(Player turn, const Board& board, float& value, FBoard& recovered) {
        std::vector<float> value_vec;
        std::vector<FBoard> policy_vec;
        if (!evaluator->eval(turn, augment_way8(board), ctx, value_vec, policy_vec)) return false;
        float sum = 0.f;
        for (float v : value_vec) sum = sum + v;
        value = sum / 8.f;
        recovered = recover_way8(policy_vec);
        for (int i = 0; i < BOARD_SIZE; ++i)
            for (int j = 0; j < BOARD_SIZE; ++j)
                recovered[i][j] *= (float)(int)(board[i][j] == Player::None);
        stats.evals += 1;
        return true;
    }

    // Dirichlet(alpha; n) twin of the engine's sampler (c6_rng.hpp: log2-space Gamma draws keyed (sim, depth, child)),
    // normalised like a softmax: w_k = 2^(L_k - Lmax), noise_k = w_k / sum w.  The sum is taken in the engine's lane
    // order (child k of the row-major child list belongs to lane k & 31: per-lane partial sums in k order, then a
    // butterfly).  f32 throughout: the reference casts its f64 draw to f32 right here (mod.rs:267).
    std::vector<float> dirichlet(int depth, size_t n) {
        std::vector<float> g(n);
        float lmax = 0.f;
        for (size_t k = 0; k < n; ++k) {
            g[k] = gamma_log2_sample(param.dirichlet_alpha, ctx.sim_index, (uint32_t)depth, (uint32_t)k, ctx.seed ^ DOMAIN_NOISE,
                                     ctx.game_id);
            if (k == 0 || g[k] > lmax) lmax = g[k];
        }
        float partial[32];
        for (auto& p : partial) p = 0.f;
        for (size_t k = 0; k < n; ++k) {
            g[k] = ex2_det(g[k] - lmax);
            partial[k & 31] = partial[k & 31] + g[k];
        }
        for (int off = 16; off >= 1; off >>= 1)
            for (int l = 0; l < off; ++l) partial[l] = partial[l] + partial[l + off];
        const float inv = 1.0f / partial[0];  // the largest draw contributes exactly 1: the sum is >= 1
        for (auto& x : g) x = x * inv;
        return g;
    }

    // :243-296
    std::optional<uint64_t> maximum_from(const Simulate& sim, int depth) {
        const Node& tree_node = map.at(hash(sim.board));
        std::vector<const Node*> child_nodes;
        child_nodes.reserve(tree_node.next_node.size());
        for (uint64_t h : tree_node.next_node) child_nodes.push_back(&map.at(h));
        if (child_nodes.empty()) return std::nullopt;
        stats.sum_ck += child_nodes.size();  // a single-child node counts C_k = 1
        if (child_nodes.size() == 1) return hash(child_nodes[0]->board);

        float epsilon = param.epsilon;
        double alpha = param.dirichlet_alpha;
        (void)alpha;
        std::vector<float> noise_v;
        if (noise == NoiseMode::Philox)
            noise_v = dirichlet(depth, child_nodes.size());
        else
            noise_v.assign(child_nodes.size(), 0.f);

        float c_puct = param.c_puct;
        int32_t vs = 0;
        for (auto* n : child_nodes) vs += n->visit;
        float visit_sum = (float)vs;
        auto puct = [&](const Node& node, float noise_f) {
            float visit = (float)node.visit;
            float prob = epsilon * noise_f + (1.f - epsilon) * node.n_prob;
            return prob * std::sqrt(visit_sum - visit) / (1.f + visit);
        };
        bool black = sim.turn == Player::Black;
        std::vector<float> probs(child_nodes.size());
        for (size_t k = 0; k < child_nodes.size(); ++k) {
            float q = child_nodes[k]->q_value;
            float un = black ? -q : q;
            probs[k] = un + c_puct * puct(*child_nodes[k], noise_v[k]);
        }
        float mx = probs[0];
        for (float p : probs)
            if (p > mx) mx = p;
        std::vector<size_t> ties;
        for (size_t k = 0; k < probs.size(); ++k)
            if (probs[k] == mx) ties.push_back(k);
        if (ties.empty()) return std::nullopt;  // NaN scores: the reference panics here
        size_t pick = ties[0];
        if (tie == TieBreak::Random) {
            tie_rng = splitmix64(tie_rng);
            pick = ties[tie_rng % ties.size()];
        } else if (tie == TieBreak::Philox && ties.size() > 1) {
            pick = ties[tie_pick(ctx.sim_index, (uint32_t)depth, ctx.seed, ctx.game_id, (uint32_t)ties.size())];
        }
        return hash(child_nodes[pick]->board);
    }

    void init(const Simulate& sim) {  // :302-306
        uint64_t h = hash(sim.board);
        if (!map.count(h)) map.emplace(h, node_new(sim.board));
    }

    std::optional<Pos> select(const Simulate& sim, int depth) {  // :313-323
        const Node& tree_node = map.at(hash(sim.board));
        auto hashed = maximum_from(sim, depth);
        if (!hashed) return std::nullopt;
        const Node& node = map.at(*hashed);
        return diff_board(node.board, tree_node.board);
    }

    void expand(const Simulate& sim) {  // :331-381
        Board board = sim.board;
        uint64_t parent_hashed = hash(board);
        {
            float cost = (float)(int)sim.turn;
            Node& node = map.at(parent_hashed);
            if (node.num_player == (size_t)BOARD_CAPACITY) {
                Player winner = search(sim.board);
                if (winner == sim.turn)
                    node.value = cost;
                else if (winner != Player::None)
                    node.value = -cost;
                return;
            }
        }
        float value;
        FBoard prob;
        if (!get_from(sim.turn, board, value, prob)) {
            std::fprintf(stderr, "alpha_zero::expand couldn't get value, prob from pyobject\n");
            std::abort();
        }
        size_t child_num;
        {
            Node& parent = map.at(parent_hashed);
            parent.value = value;
            if (faithful) parent.prob = std::make_unique<FBoard>(prob);
            parent.visit += 1;
            child_num = parent.num_player + 1;
        }
        std::vector<uint64_t> hashed_vec;
        std::vector<Pos> possible = sim.possible;  // snapshot (`sim.possible()` clones)
        Board child = board;
        // Incremental key for the non-faithful cost model (same value as hashing the whole board).
        uint64_t pk = faithful ? 0 : parent_hashed;
        const Zobrist& z = zobrist();
        for (auto [row, col] : possible) {
            // sim.simulate(row,col): place the stone, (RAII) take it back after the iteration
            child[row][col] = sim.turn;
            uint64_t hashed = faithful ? siphash13_board(child)
                                       : (pk ^ z.z[sim.turn == Player::White][row * BOARD_SIZE + col]);
            hashed_vec.push_back(hashed);
            auto it = map.find(hashed);
            if (it == map.end()) it = map.emplace(hashed, node_new_with_num(child, child_num)).first;
            it->second.n_prob = prob[row][col];
            child[row][col] = Player::None;
        }
        stats.sum_cleaf += possible.size();
        Node& parent = map.at(parent_hashed);
        for (uint64_t h : hashed_vec) parent.next_node.push_back(h);
    }

    void update(const Simulate& sim, const std::vector<Pos>& path) {  // :386-408
        if (path.empty()) return;
        auto [row, col] = path.back();
        float value = map.at(hash(sim.board)).value;
        {
            Simulate s = sim.deep_clone();
            s.rollback_in(row, col);
            map.at(hash(s.board)).q_sum += value;
        }
        Simulate s = sim.deep_clone();
        for (auto it = path.rbegin(); it != path.rend(); ++it) {
            s.rollback_in(it->first, it->second);
            Node& node = map.at(hash(s.board));
            node.visit += 1;
            node.recalc_q();
        }
    }

    void search_once(const Simulate& simulate0) {  // :411-426
        init(simulate0);
        Simulate simulate = simulate0.deep_clone();
        std::vector<Pos> path;
        while (auto pos = select(simulate, (int)path.size())) {
            path.push_back(*pos);
            simulate.simulate_in(pos->first, pos->second);
        }
        expand(simulate);
        update(simulate, path);
        stats.sims += 1;
        stats.sum_depth += path.size();
        ctx.sim_index += 1;
    }

    // Multi-leaf search, the twin of the engine's opt-in virtual-loss mode (Config::virtual_loss = K).  One wave:
    //   1. up to K descents, one after the other.  A descent applies the VISITS of update() (:400-406: every node on the path
    //      except the leaf gets visit += 1) at once — they are the virtual loss: the next descent of the wave sees them in
    //      sqrt(S - n) / (1 + n) and in q = q_sum / visit — and defers the VALUE (:394-398) until the leaf has been evaluated;
    //   2. a descent that ends in a leaf which is already waiting in this wave is discarded (it changes nothing, does not count as
    //      a simulation, and ends the wave); a full-board leaf has its value at once (:338-348);
    //   3. the waiting leaves are expanded in descent order, each with the evaluator / noise streams of its own simulation index,
    //      and each adds its value to its parent's q_sum.
    // With K = 1 this is search_once.  Every simulation index is used by one descent only (discarded ones consume theirs).
    void search_wave(const Simulate& simulate0, int& done) {
        struct Waiting {
            Simulate sim;
            std::vector<Pos> path;
            uint32_t sim_index;
            uint64_t hashed;
        };
        init(simulate0);
        std::vector<Waiting> waiting;
        for (int j = 0; j < virtual_loss && done < param.num_simulation; ++j) {
            Simulate simulate = simulate0.deep_clone();
            std::vector<Pos> path;
            while (auto pos = select(simulate, (int)path.size())) {
                path.push_back(*pos);
                simulate.simulate_in(pos->first, pos->second);
            }
            const uint64_t leaf = hash(simulate.board);
            const uint32_t index = ctx.sim_index;
            ctx.sim_index += 1;
            stats.sum_depth += path.size();
            bool dup = false;
            for (const Waiting& w : waiting) dup = dup || w.hashed == leaf;
            if (dup) break;
            {  // the visits of update(), now
                Simulate s = simulate.deep_clone();
                for (auto it = path.rbegin(); it != path.rend(); ++it) {
                    s.rollback_in(it->first, it->second);
                    Node& node = map.at(hash(s.board));
                    node.visit += 1;
                    node.recalc_q();
                }
            }
            done += 1;
            stats.sims += 1;
            if (map.at(leaf).num_player == (size_t)BOARD_CAPACITY) {
                expand(simulate);  // full board: sets the value from search(), no evaluator call
                add_value(simulate, path);
                continue;
            }
            waiting.push_back(Waiting{simulate, path, index, leaf});
        }
        for (const Waiting& w : waiting) {
            const uint32_t keep = ctx.sim_index;
            ctx.sim_index = w.sim_index;
            expand(w.sim);
            ctx.sim_index = keep;
            add_value(w.sim, w.path);
        }
    }
    void add_value(const Simulate& sim, const std::vector<Pos>& path) {  // the first half of update() (:394-398)
        if (path.empty()) return;
        const float value = map.at(hash(sim.board)).value;
        Simulate s = sim.deep_clone();
        s.rollback_in(path.back().first, path.back().second);
        Node& parent = map.at(hash(s.board));
        parent.q_sum += value;
        parent.recalc_q();
    }

    std::optional<Pos> policy(const Simulate& sim) {  // :433-452 — Rust max_by keeps the LAST maximum
        const Node& tree_node = map.at(hash(sim.board));
        std::vector<const Node*> child_node;
        for (uint64_t h : tree_node.next_node) child_node.push_back(&map.at(h));
        int32_t vs = 0;
        for (auto* n : child_node) vs += n->visit;
        float visit_sum = (float)vs;
        auto prob = [&](const Node& n) {
            float visit = (float)n.visit;
            return visit / (visit_sum - visit + 1.f);
        };
        // trace
        last_child_visits.fill(-1);
        for (auto* n : child_node) {
            Pos p = *diff_board(sim.board, n->board);
            last_child_visits[p.first * BOARD_SIZE + p.second] = n->visit;
        }
        last_root_visit = tree_node.visit;
        last_root_qsum = tree_node.q_sum;
        last_root_value = tree_node.value;
        if (child_node.empty()) return std::nullopt;
        const Node* best = child_node[0];
        for (size_t k = 1; k < child_node.size(); ++k)
            if (!(prob(*child_node[k]) < prob(*best))) best = child_node[k];  // >= : last max wins
        return diff_board(sim.board, best->board);
    }

    std::optional<Pos> next(const Game& game) {  // :464-489
        Simulate simulate = Simulate::from_game(game);
        init(simulate);
        if (virtual_loss <= 1) {
            for (int i = 0; i < param.num_simulation; ++i) search_once(simulate);
        } else {
            for (int done = 0; done < param.num_simulation;) search_wave(simulate, done);
        }
        auto res = policy(simulate);
        stats.max_nodes = std::max<uint64_t>(stats.max_nodes, map.size());
        // remove siblings
        Node node = std::move(map.at(hash(simulate.board)));
        size_t num_player = node.num_player;
        std::vector<uint64_t> sibling;
        for (auto& kv : map)
            if (kv.second.num_player == num_player) sibling.push_back(kv.first);
        for (uint64_t h : sibling) map.erase(h);
        uint64_t root_hash = hash(node.board);
        map.emplace(root_hash, std::move(node));
        if (!faithful && res) {
            // Memory valve for the lean cost model only (the reference never frees deeper nodes):
            // a node can be reached again only if its board is a superset of every later root,
            // so everything that does not contain (root + chosen stone) is dead.  Results are
            // unchanged; tests/test_oracle.py checks lean == faithful on small cases.
            Board nb = simulate.board;
            nb[res->first][res->second] = simulate.turn;
            for (auto it = map.begin(); it != map.end();) {
                const Board& b = it->second.board;
                bool sup = true;
                for (int r = 0; r < BOARD_SIZE && sup; ++r)
                    for (int c = 0; c < BOARD_SIZE; ++c)
                        if (nb[r][c] != Player::None && b[r][c] != nb[r][c]) {
                            sup = false;
                            break;
                        }
                if (!sup && it->first != root_hash)
                    it = map.erase(it);
                else
                    ++it;
            }
        }
        return res;
    }
};

inline bool PhiloxRandomEvaluator::eval(Player, const std::vector<Board>& boards, const EvalCtx& ctx,
                                        std::vector<float>& values, std::vector<FBoard>& policies) {
    // boards.size() == 8 (one leaf); plane order = augment_way8 order, so boards[6] (rotate_left four times) is the leaf itself.
    size_t len = boards.size();
    values.assign(len, 0.f);
    policies.assign(len, FBoard{});
    uint32_t k0 = ctx.seed ^ DOMAIN_EVAL, k1 = ctx.game_id;
    Philox4 va = philox4x32_10(ctx.sim_index, BOARD_CAPACITY, 0, 0, k0, k1);
    Philox4 vb = philox4x32_10(ctx.sim_index, BOARD_CAPACITY, 1, 0, k0, k1);
    for (int s = 0; s < 8 && s < (int)len; ++s) values[s] = u32_to_pm1(s < 4 ? va.v[s] : vb.v[s - 4]);
    // Original-frame draws, then carried to each plane with the same maps augment_way8 uses.  Only empty cells get draws (the
    // prior of an occupied cell is masked, mod.rs:227-232): the empty cell of row-major rank k takes eight 24-bit draws out of
    // the 384 bits of Philox blocks (sim, kb, 0..2, 0), kb = (k >> 6) << 5 | (k & 31) — draws 0..7 of that string for rank
    // kb + 64 (k >> 6) ... i.e. the lower rank of the pair (k, k + 32), draws 8..15 for the upper one.
    std::vector<FBoard> orig(8);
    for (auto& b : orig)
        for (auto& r : b) r.fill(0.f);
    if (len >= 7) {
        const Board& leaf = boards[6];
        uint32_t k = 0;
        for (int r = 0; r < BOARD_SIZE; ++r)
            for (int c = 0; c < BOARD_SIZE; ++c) {
                if (leaf[r][c] != Player::None) continue;
                const uint32_t kb = ((k >> 6) << 5) | (k & 31u), upper = (k >> 5) & 1u;
                uint32_t w[12];
                for (uint32_t j = 0; j < 3; ++j) {
                    Philox4 blk = philox4x32_10(ctx.sim_index, kb, j, 0, k0, k1);
                    for (int t = 0; t < 4; ++t) w[4 * j + t] = blk.v[t];
                }
                for (int s = 0; s < 8; ++s) {
                    const uint32_t bit = 24u * (8u * upper + (uint32_t)s), q = bit >> 5, sh = bit & 31u;
                    const uint64_t two = (uint64_t)w[q] | ((uint64_t)(q + 1 < 12 ? w[q + 1] : 0u) << 32);
                    const uint32_t x = (uint32_t)(two >> sh) & 0xFFFFFFu;
                    orig[s][r][c] = (float)x * (1.0f / 8388608.0f) - 1.0f;
                }
                k += 1;
            }
    }
    for (int k = 0; k < 4; ++k) {
        FBoard even = orig[2 * k], odd = orig[2 * k + 1];
        for (int t = 0; t <= k; ++t) {
            rotate_left(even);
            rotate_left(odd);
        }
        flip_vertical(odd);
        if (2 * k < (int)len) policies[2 * k] = even;
        if (2 * k + 1 < (int)len) policies[2 * k + 1] = odd;
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// agent/agent_impl/mod.rs:34-44,121-178
// ------------------------------------------------------------------------------------------
struct Path {
    Player turn;
    Board board;
    Pos pos;
};
struct PlayResult {
    Player winner = Player::None;
    std::vector<Path> path;
    std::string error;  // non-empty = Err(description)
};
struct Policy {  // policy/mod.rs:45-48
    virtual ~Policy() {}
    virtual std::optional<Pos> next(const Game& game) = 0;
};
struct AlphaZeroPolicy : Policy {
    AlphaZero az;
    // per-move trace (filled when `trace` is set)
    bool trace = false;
    std::vector<std::array<int32_t, BOARD_CAPACITY>> child_visits;
    std::vector<int32_t> root_visit;
    std::vector<float> root_qsum;
    AlphaZeroPolicy(Evaluator* e, HyperParameter p) : az(e, p) {}
    std::optional<Pos> next(const Game& game) override {
        auto r = az.next(game);
        if (trace) {
            child_visits.push_back(az.last_child_visits);
            root_visit.push_back(az.last_root_visit);
            root_qsum.push_back(az.last_root_qsum);
        }
        return r;
    }
};

inline PlayResult agent_play(Policy& policy, bool debug = false, FILE* out = stdout) {
    PlayResult res;
    Game game;
    for (;;) {
        if (debug) std::fputs(game.render().c_str(), out);
        auto pos = policy.next(game);
        if (!pos) break;
        res.path.push_back(Path{game.turn, game.board, *pos});
        SetResult sr;
        std::string err = game.set(*pos, &sr);
        if (!err.empty()) {
            res.error = err;
            return res;
        }
        if (debug)
            std::fprintf(out, "%s (%c, %c), remain %d, 0.0 elapsed\n",
                         sr.player == Player::Black ? "Black" : "White", (char)(sr.position.first + 0x61),
                         (char)(sr.position.second + 0x41), sr.num_remain);
        Player is_end = game.is_game_end();
        if (is_end != Player::None) {
            res.winner = is_end;
            break;
        }
    }
    if (debug) std::fputs(game.render().c_str(), out);
    return res;
}

}  // namespace c6o
