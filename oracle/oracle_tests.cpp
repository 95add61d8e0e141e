// ORACLE — TEST INFRASTRUCTURE ONLY.
// Ports of the reference's own known-answer tests onto the restatement in c6_oracle.hpp.  This is
// what "pins" the oracle (the Rust reference cannot be executed here).  Each case cites the
// reference test it restates, relative to /root/reference/Connect6/.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "c6_oracle.hpp"
#include "c6_default_policy.hpp"

using namespace c6o;

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                              \
    do {                                                                         \
        ++g_checks;                                                              \
        if (!(cond)) {                                                           \
            ++g_fail;                                                            \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);          \
        }                                                                        \
    } while (0)

constexpr Player B = Player::Black, W = Player::White, N = Player::None;

// libconnect6/src/game/player/tests.rs
static void test_player() {
    CHECK(switch_player(B) == W);
    CHECK(switch_player(W) == B);
    CHECK(switch_player(N) == N);
    CHECK(player_from(-1) == B && player_from(1) == W && player_from(0) == N && player_from(7) == N);
}

// libconnect6/src/game/search_winner/tests.rs:102-139
static void test_base_search() {
    Board t = empty_board();
    for (int i = 0; i < 5; ++i) t[0][i] = B;
    CHECK(search(t) == N);
    t[0][5] = B;
    CHECK(search(t) == B);
    for (int i = 0; i < 5; ++i) t[5 - i][0] = W;
    CHECK(search(t) == B);
    t[0][0] = W;
    CHECK(search(t) == W);
    for (int i = 0; i < 5; ++i) t[5 - i][5 - i] = B;
    CHECK(search(t) == W);
    t[0][0] = B;
    CHECK(search(t) == B);
    for (int i = 0; i < 5; ++i) t[5 - i][i] = W;
    CHECK(search(t) == B);
    t[0][5] = W;
    CHECK(search(t) == W);  // both colours own a six: scan order decides
}

// search_winner/tests.rs:141-182
static void test_boundary_search() {
    const int S = BOARD_SIZE;
    int bases[2] = {0, S - 1}, inits[3] = {0, 1, S - 6};
    for (int base : bases)
        for (int init : inits) {
            Board t = empty_board();
            for (int i = 0; i < 5; ++i) t[base][init + i] = W;
            CHECK(search(t) == N);
            t[base][init + 5] = W;
            CHECK(search(t) == W);
            Board u = empty_board();
            for (int i = 0; i < 5; ++i) u[init + i][base] = B;
            CHECK(search(u) == N);
            u[init + 5][base] = B;
            CHECK(search(u) == B);
        }
}

// search_winner/tests.rs:184-216
static void test_cross_search() {
    const int S = BOARD_SIZE;
    int inits[3] = {0, 1, S - 6};
    for (int init : inits) {
        Board t = empty_board();
        for (int i = 0; i < 5; ++i) t[init + i][init + i] = W;
        CHECK(search(t) == N);
        t[init + 5][init + 5] = W;
        CHECK(search(t) == W);
        Board u = empty_board();
        for (int i = 0; i < 5; ++i) u[init + i][S - 1 - init - i] = B;
        CHECK(search(u) == N);
        u[init + 5][S - init - 6] = B;
        CHECK(search(u) == B);
    }
}

// libconnect6/src/game/game_impl/tests.rs:4-52
static void test_game() {
    Game g;
    CHECK(g.turn == B && g.num_remain == 1 && g.board == empty_board());
    SetResult r;
    CHECK(g.set(Pos(0, 0), &r).empty());
    CHECK(r.player == B && r.num_remain == 0 && r.position == Pos(0, 0));
    CHECK(g.turn == W && g.num_remain == 2 && g.board[0][0] == B);
    CHECK(g.set(Pos(0, 0)) == "already set position");
    CHECK(g.set(Pos(BOARD_SIZE, BOARD_SIZE)) == "invalid position");
    CHECK(Game().is_game_end() == N);
}

// libconnect6/src/policy/simulate/tests.rs:10-138
static void test_simulate() {
    Simulate s = Simulate::from_game(Game());
    CHECK(s.turn == B && s.num_remain == 1 && s.board == empty_board() && (int)s.possible.size() == BOARD_CAPACITY);
    Simulate c = s.deep_clone();
    s.board[0][0] = B;
    CHECK(c.board[0][0] == N);
    CHECK(!s.validate(0, 0));
    CHECK(c.validate(0, 0));
    CHECK(!c.validate(17, 3) && !c.validate(3, 40));
    // next_turn (:72-81)
    Simulate n = Simulate::from_game(Game());
    CHECK(n.next_turn() == W);
    n.simulate_in(0, 0);
    CHECK(n.next_turn() == W);
    n.simulate_in(0, 1);
    CHECK(n.next_turn() == B);
    // simulate_in (:105-118)
    Simulate a = Simulate::from_game(Game());
    a.simulate_in(0, 0);
    CHECK(a.board[0][0] == B && a.turn == W && a.num_remain == 2);
    CHECK(std::find(a.possible.begin(), a.possible.end(), Pos(0, 0)) == a.possible.end());
    // rollback_in (:120-138)
    a.rollback_in(0, 0);
    CHECK(a.turn == B && a.num_remain == 1 && a.board == empty_board() && (int)a.possible.size() == BOARD_CAPACITY);
    CHECK(a.possible.back() == Pos(0, 0));  // pushed to the END (simulate/mod.rs:295)
}

// libconnect6/src/policy/alphazero_policy/augment_tests.rs:5-94
static Board corner_board() {
    const int S = BOARD_SIZE;
    Board b = empty_board();
    b[0][0] = B, b[0][1] = B, b[0][S - 1] = W, b[1][S - 1] = B;
    b[S - 1][S - 1] = W, b[S - 1][S - 2] = W, b[S - 1][0] = B, b[S - 2][0] = W;
    return b;
}
static void test_augment() {
    const int S = BOARD_SIZE;
    Board b = corner_board();
    rotate_left(b);
    CHECK(b[S - 1][0] == B && b[S - 2][0] == B && b[0][0] == W && b[0][1] == B);
    CHECK(b[0][S - 1] == W && b[1][S - 1] == W && b[S - 1][S - 1] == B && b[S - 1][S - 2] == W);
    b = corner_board();
    rotate_right(b);
    CHECK(b[S - 1][0] == W && b[S - 2][0] == W && b[0][0] == B && b[0][1] == W);
    CHECK(b[0][S - 1] == B && b[1][S - 1] == B && b[S - 1][S - 1] == W && b[S - 1][S - 2] == B);
    b = corner_board();
    flip_vertical(b);
    CHECK(b[S - 1][0] == W && b[S - 1][1] == W && b[0][0] == W && b[1][0] == B);
    CHECK(b[0][S - 1] == B && b[0][S - 2] == B && b[S - 1][S - 1] == B && b[S - 2][S - 1] == W);
    b = corner_board();
    flip_horizontal(b);
    CHECK(b[S - 1][0] == B && b[S - 1][1] == B && b[0][0] == B && b[1][0] == W);
    CHECK(b[0][S - 1] == W && b[0][S - 2] == W && b[S - 1][S - 1] == W && b[S - 2][S - 1] == B);
    // sum_board (:97-112)
    GBoard<int> b1{}, b2{};
    int mx = S * S;
    for (int i = 0; i < S; ++i)
        for (int j = 0; j < S; ++j) b1[i][j] = i * S + j, b2[i][j] = mx - b1[i][j];
    sum_board(b1, b2);
    for (auto& r : b1)
        for (int x : r) CHECK(x == mx);
    // augment -> recover round trip on random boards (:115-142)
    uint64_t s = 42;
    for (int rep = 0; rep < 20; ++rep) {
        Board rb;
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) {
                s = splitmix64(s);
                rb[i][j] = player_from((int)(s % 3) - 1);
            }
        auto aug = augment_way8(rb);
        CHECK(aug.size() == 8);
        for (int i = 0; i < 8; ++i)
            for (int j = i + 1; j < 8; ++j) CHECK(aug[i] != aug[j]);
        std::vector<FBoard> conv(8);
        for (int k = 0; k < 8; ++k)
            for (int i = 0; i < S; ++i)
                for (int j = 0; j < S; ++j) conv[k][i][j] = (float)(int)aug[k][i][j];
        FBoard rec = recover_way8(conv);
        for (int i = 0; i < S; ++i)
            for (int j = 0; j < S; ++j) CHECK(rec[i][j] == (float)(int)rb[i][j]);
    }
}

// libconnect6/src/agent/agent_impl/tests.rs:37-129
struct Script : Policy {
    std::vector<Pos> m;
    size_t i = 0;
    std::optional<Pos> next(const Game&) override { return i < m.size() ? std::optional<Pos>(m[i++]) : std::nullopt; }
};
static void test_agent() {
    {
        Script p;
        p.m = {{BOARD_SIZE, BOARD_SIZE}};
        CHECK(agent_play(p).error == "invalid position");
    }
    {
        Script p;
        p.m = {{0, 0}, {0, 0}};
        CHECK(agent_play(p).error == "already set position");
    }
    Script p;
    p.m = {{0, 0}, {0, 1}, {0, 2}, {1, 0}, {2, 0}, {0, 3}, {0, 4}, {3, 0}, {4, 0}, {0, 5}, {0, 6}};
    PlayResult r = agent_play(p);
    CHECK(r.error.empty());
    CHECK(r.winner == W);
    CHECK(r.path.size() == 11);
    Player turn = B;
    int num_remain = 1;
    Board board = empty_board();
    for (size_t k = 0; k < r.path.size() && k < p.m.size(); ++k) {
        CHECK(r.path[k].turn == turn);
        CHECK(r.path[k].board == board);
        CHECK(r.path[k].pos == p.m[k]);
        board[p.m[k].first][p.m[k].second] = turn;
        if (--num_remain <= 0) num_remain = 2, turn = switch_player(turn);
    }
}

// libconnect6/src/policy/alphazero_policy/tests.rs:7-204 (RandomEvaluator -> Philox twin)
static void test_alphazero_invariants(bool faithful) {
    PhiloxRandomEvaluator ev;
    // test_expand (:43-88)
    {
        AlphaZero pol(&ev, HyperParameter());
        pol.faithful = faithful;
        Simulate sim = Simulate::from_game(Game());
        pol.init(sim);
        while (auto pos = pol.select(sim, 0)) sim.simulate_in(pos->first, pos->second);
        pol.expand(sim);
        Simulate fresh;
        auto it = pol.map.find(pol.hash(fresh.board));
        CHECK(it != pol.map.end());
        const Node& root = it->second;
        CHECK(root.visit == 1);
        CHECK(root.value != 0.f);
        CHECK(root.q_value == 0.f && root.n_prob == 0.f && root.num_player == 0);
        CHECK(root.board == fresh.board);
        CHECK((int)root.next_node.size() == BOARD_CAPACITY);
        if (faithful) CHECK(root.prob != nullptr);
        for (int i = 0; i < BOARD_SIZE; ++i)
            for (int j = 0; j < BOARD_SIZE; ++j) {
                Board b = fresh.board;
                b[i][j] = B;
                uint64_t h = pol.hash(b);
                CHECK(pol.map.count(h) == 1);
                CHECK(std::find(root.next_node.begin(), root.next_node.end(), h) != root.next_node.end());
                const Node& n = pol.map.at(h);
                CHECK(n.visit == 0 && n.value == 0.f && n.num_player == 1 && n.next_node.empty());
                if (faithful) CHECK(n.n_prob == (*root.prob)[i][j]);
                CHECK(n.prob == nullptr);
            }
    }
    // test_update / test_policy / test_select (:7-40, :91-183)
    {
        Game game;
        AlphaZero pol(&ev, HyperParameter());
        pol.faithful = faithful;
        Simulate sim0 = Simulate::from_game(game);
        pol.init(sim0);
        for (int rep = 0; rep < 2; ++rep) {
            Simulate sim = sim0.deep_clone();
            std::vector<Pos> path;
            while (auto pos = pol.select(sim, (int)path.size())) {
                CHECK(sim.validate(pos->first, pos->second));
                sim.simulate_in(pos->first, pos->second);
                path.push_back(*pos);
            }
            pol.expand(sim);
            pol.update(sim, path);
            pol.ctx.sim_index += 1;
        }
        const Node& node = pol.map.at(pol.hash(game.board));
        CHECK(node.q_value != 0.f);
        CHECK(node.visit == 2);
        CHECK((int)node.next_node.size() == BOARD_CAPACITY);
        std::vector<uint64_t> ch;
        for (uint64_t h : node.next_node)
            if (pol.map.at(h).value != 0.f) ch.push_back(h);
        CHECK(ch.size() == 1);
        if (ch.size() == 1) {
            const Node& child = pol.map.at(ch[0]);
            CHECK(child.n_prob != 0.f);
            CHECK(node.q_value * 2.f == child.value);
            auto d = diff_board(node.board, child.board);
            CHECK(d.has_value());
            if (faithful && d) CHECK((*node.prob)[d->first][d->second] == child.n_prob);
            auto pos = pol.policy(sim0);
            CHECK(pos.has_value());
            if (pos && d) CHECK(*pos == *d);
            auto sel = pol.select(sim0, 0);
            CHECK(sel.has_value() && sim0.validate(sel->first, sel->second));
        }
    }
    // test_self_play (:186-204): light_weight, last.turn == winner
    for (uint32_t gid = 0; gid < 3; ++gid) {
        AlphaZeroPolicy pol(&ev, HyperParameter::light_weight());
        pol.az.faithful = faithful;
        pol.az.ctx.game_id = gid;
        PlayResult r = agent_play(pol);
        CHECK(r.error.empty());
        if (!r.path.empty() && r.winner != N) CHECK(r.path.back().turn == r.winner);
    }
}

// cppbind/libcppconnect6/src/cppbind/cppeval/tests.rs:8-92
static void echo_callback(int player, float* values, float* policies, int len) {
    for (int i = 0; i < len; ++i) values[i] = (float)i;
    for (int i = 0; i < len * BOARD_CAPACITY; ++i) policies[i] *= 2.f;
    values[0] = (float)player;
}
static void test_cppeval() {
    CppEval ev(echo_callback);
    std::vector<Board> boards;
    uint64_t s = 7;
    for (int n = 0; n < 13; ++n) {
        Board b;
        for (auto& r : b)
            for (auto& x : r) s = splitmix64(s), x = player_from((int)(s % 3) - 1);
        boards.push_back(b);
    }
    std::vector<float> v;
    std::vector<FBoard> p;
    CHECK(ev.eval(B, boards, EvalCtx(), v, p));
    CHECK(v.size() == 13 && v[0] == -1.f);
    for (int n = 1; n < 13; ++n) CHECK(v[n] == (float)n);
    for (int n = 0; n < 13; ++n)
        for (int i = 0; i < BOARD_SIZE; ++i)
            for (int j = 0; j < BOARD_SIZE; ++j) CHECK(p[n][i][j] == 2.f * (float)(int)boards[n][i][j]);
}

// SURVEY.md Appendix B.1 — hand-derived trace (constant evaluator, 4 sims, eps = 0, lowest index).
static void const_cb(int, float* values, float* policies, int len) {
    for (int n = 0; n < len; ++n) values[n] = 0.5f;
    for (int i = 0; i < len * BOARD_CAPACITY; ++i) policies[i] = 0.25f;
}
static void test_hand_trace(bool faithful) {
    CppEval ev(const_cb);
    HyperParameter hp;
    hp.num_simulation = 4, hp.epsilon = 0.f;
    AlphaZeroPolicy pol(&ev, hp);
    pol.az.noise = NoiseMode::None;
    pol.az.faithful = faithful;
    Game g;
    auto m1 = pol.next(g);
    CHECK(m1 && *m1 == Pos(0, 2));
    CHECK(pol.az.last_root_visit == 4 && pol.az.last_root_qsum == 1.5f);
    g.set(*m1);
    auto m2 = pol.next(g);
    CHECK(m2 && *m2 == Pos(0, 4));
    CHECK(pol.az.last_root_visit == 5 && pol.az.last_root_qsum == 2.0f);
    g.set(*m2);
    auto m3 = pol.next(g);
    CHECK(m3 && *m3 == Pos(0, 5));
    g.set(*m3);
    Pos want[] = {{0, 6}, {0, 7}, {0, 8}, {0, 9}};
    for (Pos w : want) {
        auto m = pol.next(g);
        CHECK(m && *m == w);
        g.set(*m);
    }
}

// lean and faithful cost models give identical games
static void test_lean_equals_faithful() {
    PhiloxRandomEvaluator ev;
    HyperParameter hp;
    hp.num_simulation = 24;
    for (int noise = 0; noise < 2; ++noise) {
        hp.epsilon = noise ? 0.25f : 0.f;
        std::vector<Pos> seq[2];
        Player win[2];
        for (int f = 0; f < 2; ++f) {
            AlphaZeroPolicy pol(&ev, hp);
            pol.az.faithful = f;
            pol.az.noise = noise ? NoiseMode::Philox : NoiseMode::None;
            pol.az.ctx.seed = 5, pol.az.ctx.game_id = 3;
            Game g;
            for (int m = 0; m < 30; ++m) {
                auto p = pol.next(g);
                if (!p) break;
                seq[f].push_back(*p);
                g.set(*p);
                if (g.is_game_end() != N) break;
            }
            win[f] = g.is_game_end();
        }
        CHECK(seq[0] == seq[1]);
        CHECK(win[0] == win[1]);
    }
}

// Philox4x32-10 known answers (Random123 kat_vectors) + sanity of the fixed-polynomial math
static void test_rng() {
    Philox4 a = philox4x32_10(0, 0, 0, 0, 0, 0);
    CHECK(a.v[0] == 0x6627e8d5u && a.v[1] == 0xe169c58du && a.v[2] == 0xbc57ac4cu && a.v[3] == 0x9b00dbd8u);
    Philox4 b = philox4x32_10(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    CHECK(b.v[0] == 0x408f276du && b.v[1] == 0x41c83b0eu && b.v[2] == 0xa20bc7c6u && b.v[3] == 0x6d5451fdu);
    Philox4 c = philox4x32_10(0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u, 0xa4093822u, 0x299f31d0u);
    CHECK(c.v[0] == 0xd16cfe09u && c.v[1] == 0x94fdccebu && c.v[2] == 0x5001e420u && c.v[3] == 0x24126ea1u);
    double xs[] = {1e-300, 1e-17, 0.001, 0.3, 0.9999, 1.0, 1.5, 2.0, 10.0, 12345.678, 1e300};
    for (double x : xs) CHECK(std::fabs(det_log(x) - std::log(x)) <= 4e-16 * std::fmax(1.0, std::fabs(std::log(x))));
    double es[] = {-744.0, -700.0, -100.5, -1.0, -1e-9, 0.0, 0.3, 1.0, 5.5};
    for (double x : es) CHECK(std::fabs(det_exp(x) - std::exp(x)) <= 1e-14 * std::exp(x) + 1e-320);
    CHECK(det_exp(-800.0) == 0.0);
    // Gamma(0.03): mean alpha; Gamma(2.5): mean 2.5 (loose Monte-Carlo check of the sampler)
    // and of the shape of the distribution: E[X] = Var[X] = shape, E[X^2] = shape (shape + 1),
    // P(X <= x) for Gamma(0.03) at x = 1e-30 is 1e-0.9 / Gamma(1.03) ~ 0.1274 (x^a / Gamma(a+1) for tiny x)
    double m1 = 0, m2 = 0, q1 = 0, q2 = 0, m3 = 0, q3 = 0;
    int tiny = 0;
    const int NS = 400000;
    for (int i = 0; i < NS; ++i) {
        double g1 = gamma_sample(0.03, i, 0, 0, 1, 2), g2 = gamma_sample(2.5, i, 1, 0, 1, 2), g3 = gamma_sample(0.7, i, 2, 0, 1, 2);
        m1 += g1, q1 += g1 * g1, m2 += g2, q2 += g2 * g2, m3 += g3, q3 += g3 * g3;
        tiny += g1 <= 1e-30;
    }
    CHECK(std::fabs(m1 / NS - 0.03) < 0.002);
    CHECK(std::fabs(q1 / NS - 0.03 * 1.03) < 0.004);
    CHECK(std::fabs((double)tiny / NS - 0.1274) < 0.003);
    CHECK(std::fabs(m2 / NS - 2.5) < 0.02);
    CHECK(std::fabs(q2 / NS - 2.5 * 3.5) < 0.15);
    CHECK(std::fabs(m3 / NS - 0.7) < 0.008);
    CHECK(std::fabs(q3 / NS - 0.7 * 1.7) < 0.03);
}


// libconnect6/src/policy/default_policy/tests.rs:6-31 (test_select), :34-63 (test_expand), :66-112 (test_update),
// :115-131 (test_next), :134-150 (test_self_play)
static bool in_possible(const Simulate& sim, Pos p) {
    return std::find(sim.possible.begin(), sim.possible.end(), p) != sim.possible.end();
}
static void test_default_policy() {
    {  // test_select: expand + update until some child has a non-zero win ratio, then select returns a legal move
        Game game;
        Simulate sim = Simulate::from_game(game);
        DefaultPolicy policy(50, 0xC6, 0);
        policy.init(sim);
        for (int guard = 0; guard < 1000; ++guard) {
            if (auto pos = policy.select(sim)) {
                CHECK(in_possible(sim, *pos));
                CHECK(sim.validate(pos->first, pos->second));
                break;
            }
            CHECK(!policy.select(sim));
            policy.cur_i = (uint32_t)guard;
            Pos p = policy.expand(sim);
            Simulate expanded = sim;
            expanded.simulate_in(p.first, p.second);
            policy.update(expanded, {p});
        }
    }
    {  // test_expand
        Game game;
        Simulate sim = Simulate::from_game(game);
        DefaultPolicy policy(50, 0xC6, 1);
        policy.init(sim);
        while (auto pos = policy.select(sim)) sim.simulate_in(pos->first, pos->second);
        Pos p = policy.expand(sim);
        CHECK(in_possible(sim, p));
        CHECK(sim.validate(p.first, p.second));
        CHECK(policy.map.size() == 2);
        auto it = policy.map.find(DefaultPolicy::hash(empty_board()));
        CHECK(it != policy.map.end());
        CHECK(it->second.next_node.size() == 1);
        CHECK(policy.map.count(it->second.next_node[0]) == 1);
    }
    {  // test_update
        Game game;
        Simulate sim = Simulate::from_game(game);
        DefaultPolicy policy(50, 0xC6, 2);
        policy.init(sim);
        std::vector<Pos> path;
        while (auto pos = policy.select(sim)) {
            sim.simulate_in(pos->first, pos->second);
            path.push_back(*pos);
        }
        Pos p = policy.expand(sim);
        Simulate expanded = sim;
        expanded.simulate_in(p.first, p.second);
        path.push_back(p);
        policy.update(expanded, path);
        CHECK(policy.map.size() == 2);
        const DpNode& parent = policy.map.at(DefaultPolicy::hash(empty_board()));
        CHECK(parent.next_node.size() == 1);
        CHECK(parent.visit == 1);  // a random rollout from one stone always ends with a winner in practice
        const DpNode& child = policy.map.at(parent.next_node[0]);
        CHECK(child.next_node.empty());
        CHECK(child.visit == 1);
        CHECK(child.black_win == parent.black_win);
        int num = 0;
        for (auto& row : child.board)
            for (auto c : row) num += c != N;
        CHECK(num == 1);
    }
    {  // test_next
        Game game;
        Simulate sim = Simulate::from_game(game);
        DefaultPolicy policy(50, 0xC6, 3);
        policy.init(sim);
        auto pos = policy.policy(sim);
        CHECK(pos.has_value());
        CHECK(sim.validate(pos->first, pos->second));
        CHECK(in_possible(sim, *pos));
    }
    for (uint32_t game_id = 0; game_id < 4; ++game_id) {  // test_self_play (with_num_iter(10)) + the doc test (2)
        DefaultPolicy policy(game_id == 3 ? 2 : 10, 0xC6, game_id);
        PlayResult result = agent_play(policy);
        CHECK(result.error.empty());
        if (!result.path.empty() && result.winner != N) CHECK(result.path.back().turn == result.winner);
        CHECK(policy.stats.searches >= (uint64_t)policy.num_iter * result.path.size());
    }
}

int main() {
    test_player();
    test_base_search();
    test_boundary_search();
    test_cross_search();
    test_game();
    test_simulate();
    test_augment();
    test_agent();
    test_cppeval();
    test_rng();
    test_alphazero_invariants(true);
    test_alphazero_invariants(false);
    test_hand_trace(true);
    test_hand_trace(false);
    test_lean_equals_faithful();
    test_default_policy();
    std::printf("%d checks, %d failures\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
