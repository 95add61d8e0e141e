// Test program of the host-side mirror of the reference's Rust API (alphazero-connect6_b200/csrc/connect6_host.hpp,
// connect6_alphazero.hpp), built by tests/test_abi.py::build_host_mirror and run on the GPU box.  The checks follow the
// reference's own tests: LC/agent/async_agent/tests.rs:8-47 (run(4): four results, last.turn == winner),
// LC/policy/alphazero_policy/tests.rs:186-204 (self-play with the light-weight budget), LC/agent/agent_impl/tests.rs.
#include <cstdio>
#include <cstdlib>

#include "connect6_alphazero.hpp"

using namespace connect6;

#define CHECK(c)                                                      \
    do {                                                              \
        if (!(c)) {                                                   \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                 \
        }                                                             \
    } while (0)

static bool consistent(const PlayResult& r) {  // path.last().turn == winner whenever somebody won; boards replay
    Game g;
    for (const Path& p : r.path) {
        if (p.turn != g.get_turn() || p.board != g.get_board()) return false;
        if (!g.set(p.pos).empty()) return false;
    }
    if (r.winner != Player::None && (r.path.empty() || r.path.back().turn != r.winner)) return false;
    return g.is_game_end() == r.winner;
}

struct FirstEmpty : Policy {  // a policy with no search in it: first empty cell in row-major order
    std::optional<Pos> next(const Game& game) override {
        for (size_t i = 0; i < BOARD_SIZE; ++i)
            for (size_t j = 0; j < BOARD_SIZE; ++j)
                if (game.get_board()[i][j] == Player::None) return Pos(i, j);
        return std::nullopt;
    }
};

int main() {
    // RandomEvaluator: shapes and range (alphazero_policy/mod.rs:137-169)
    RandomEvaluator rev;
    std::vector<Board> boards(8);
    auto res = rev.eval(Player::Black, boards);
    CHECK(res && res->first.size() == 8 && res->second.size() == 8);
    float lo = 1.f, hi = -1.f;
    for (const FBoard& b : res->second)
        for (const auto& row : b)
            for (float x : row) lo = std::min(lo, x), hi = std::max(hi, x);
    CHECK(lo >= -1.f && hi < 1.f && lo < -0.9f && hi > 0.9f);
    CHECK(rev.device_kind() == c6::EVAL_PHILOX);

    // AsyncAgent, generic form (async_agent/tests.rs:8-47) over a search-free policy
    auto gen = [] { return std::unique_ptr<Policy>(new FirstEmpty()); };
    auto results = make_async_agent(gen).run(4);
    CHECK(results.size() == 4);
    for (const PlayResult& r : results) CHECK(consistent(r) && r.winner == Player::Black && r.path.size() == 81);  // row-major fill: Black's column of six

    // AsyncAgent over AlphaZero + RandomEvaluator with the light-weight budget, one GPU engine per game
    auto az_gen = [] { return std::unique_ptr<Policy>(new AlphaZero(std::unique_ptr<Evaluator>(new RandomEvaluator()), HyperParameter::light_weight(), 0xC6, 0, 12)); };
    auto az = make_async_agent(az_gen).run(3);
    CHECK(az.size() == 3);
    for (const PlayResult& r : az) CHECK(consistent(r) && r.path.size() >= 11);

    // the batched form: all games in one engine; device twin of the RandomEvaluator and a host evaluator
    HyperParameter hp = HyperParameter::light_weight();
    auto batched = AlphaZeroAsyncAgent(std::make_shared<RandomEvaluator>(), hp, 7, 0, 8192, 12).run(64);
    CHECK(batched.size() == 64);
    size_t stones = 0;
    for (const PlayResult& r : batched) {
        CHECK(consistent(r));
        stones += r.path.size();
    }
    struct HostRandom : RandomEvaluator {
        int device_kind() const override { return -1; }  // force the host path: eval() through eval_raw()
    };
    auto hosted = AlphaZeroAsyncAgent(std::make_shared<HostRandom>(), hp, 7, 0, 8192, 12).run(5);
    CHECK(hosted.size() == 5);
    for (const PlayResult& r : hosted) CHECK(consistent(r));
    // the same seed twice gives the same games (counter-based noise and tie-break; the device evaluator is keyed too)
    auto again = AlphaZeroAsyncAgent(std::make_shared<RandomEvaluator>(), hp, 7, 0, 8192, 12).run(64);
    for (size_t g = 0; g < 64; ++g) {
        CHECK(again[g].path.size() == batched[g].path.size());
        for (size_t m = 0; m < again[g].path.size(); ++m) CHECK(again[g].path[m].pos == batched[g].path[m].pos);
    }
    std::printf("host mirror ok: %zu stones in 64 batched games\n", stones);
    return 0;
}
