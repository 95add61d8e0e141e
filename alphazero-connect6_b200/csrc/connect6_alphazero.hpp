// Host-side mirror of the reference's search-policy API, served by the GPU engine:
//   HyperParameter          LC/policy/alphazero_policy/mod.rs:90-125
//   Evaluator (trait)       LC/policy/alphazero_policy/mod.rs:128-134
//   AlphaZero: Policy       LC/policy/alphazero_policy/mod.rs:186-209,464-489   (new / with_param / next)
//   RandomEvaluator         LC/policy/alphazero_policy/mod.rs:137-169
//   CppEval                 CB/libcppconnect6/src/cppbind/cppeval/mod.rs:8-55
//   AsyncAgent              LC/agent/async_agent/mod.rs:61-137 (generic thread-per-game form + the batched GPU form)
// `AlphaZero::next(&game)` runs num_simulation searches for that game on the GPU (an engine with one
// resident game) and returns the chosen position.  Stones that were placed by somebody else since the last
// call are fed to the engine first, so the tree is reused exactly as the reference's `init` reuses it.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <functional>
#include <memory>
#include <mutex>
#include <random>
#include <stdexcept>
#include <thread>
#include <vector>

#include "c6_engine.cuh"
#include "connect6_host.hpp"

namespace connect6 {

struct HyperParameter {
    int num_simulation = 800;
    float epsilon = 0.25f;
    double dirichlet_alpha = 0.03;
    float c_puct = 1.f;
    static HyperParameter light_weight() {
        HyperParameter p;
        p.num_simulation = 2;
        return p;
    }
};

using FBoard = std::array<std::array<float, BOARD_SIZE>, BOARD_SIZE>;

struct Evaluator {
    virtual ~Evaluator() {}
    // None -> the search aborts ("couldn't get value, prob", alphazero_policy/mod.rs:378-380)
    virtual std::optional<std::pair<std::vector<float>, std::vector<FBoard>>> eval(Player turn, const std::vector<Board>& boards) = 0;
    // Batched form used by the engine (one call per side to move and wave): boards arrive as len x 225 floats
    // (-1/0/+1) in `policies`, which is overwritten with the priors.  The default adapts to eval() above.
    // An evaluator the engine has a device twin of says so here (c6::EvalKind), and the search then runs without
    // materialising the leaf batch at all; -1 = evaluate through eval() / eval_raw() on the host.
    virtual int device_kind() const { return -1; }
    virtual void eval_raw(int player, float* values, float* policies, int len) {
        std::vector<Board> boards((size_t)len);
        for (int n = 0; n < len; ++n)
            for (size_t i = 0; i < BOARD_SIZE; ++i)
                for (size_t j = 0; j < BOARD_SIZE; ++j) boards[n][i][j] = player_from((int)policies[(size_t)n * BOARD_CAPACITY + i * BOARD_SIZE + j]);
        auto res = eval(player_from(player), boards);
        if (!res || (int)res->first.size() != len || (int)res->second.size() != len)
            throw std::runtime_error("alpha_zero::expand couldn't get value, prob");
        for (int n = 0; n < len; ++n) {
            values[n] = res->first[n];
            for (size_t i = 0; i < BOARD_SIZE; ++i)
                for (size_t j = 0; j < BOARD_SIZE; ++j) policies[(size_t)n * BOARD_CAPACITY + i * BOARD_SIZE + j] = res->second[n][i][j];
        }
    }
};

// RandomEvaluator (alphazero_policy/mod.rs:137-169): every value and every policy entry ~ U(-1, 1) from an OS-seeded
// generator (the reference's thread_rng).  Handed to AlphaZero it is served by the engine's Philox twin on the device
// (fused into the expansion: the reference's sample program, `cargo run -p sample`, never leaves the GPU).
class RandomEvaluator : public Evaluator {
   public:
    static FBoard rand_board() {  // :139-148
        FBoard b;
        for (auto& row : b)
            for (auto& x : row) x = unit();
        return b;
    }
    std::optional<std::pair<std::vector<float>, std::vector<FBoard>>> eval(Player, const std::vector<Board>& boards) override {
        std::vector<float> values;
        std::vector<FBoard> policies;
        for (size_t i = 0; i < boards.size(); ++i) {
            values.push_back(unit());
            policies.push_back(rand_board());
        }
        return std::make_pair(std::move(values), std::move(policies));
    }
    int device_kind() const override { return c6::EVAL_PHILOX; }

   private:
    static float unit() {
        static thread_local std::mt19937 gen{std::random_device{}()};
        return std::uniform_real_distribution<float>(-1.f, 1.f)(gen);
    }
};

using Callback = void (*)(int player, float* values, float* policies, int len);
class CppEval : public Evaluator {
   public:
    explicit CppEval(Callback cb) : cb_(cb) {}
    std::optional<std::pair<std::vector<float>, std::vector<FBoard>>> eval(Player turn, const std::vector<Board>& boards) override {
        const int len = (int)boards.size();
        std::vector<float> values((size_t)len, 0.f);
        std::vector<FBoard> pol((size_t)len);
        for (int n = 0; n < len; ++n)
            for (size_t i = 0; i < BOARD_SIZE; ++i)
                for (size_t j = 0; j < BOARD_SIZE; ++j) pol[n][i][j] = (float)(int)boards[n][i][j];
        cb_((int)turn, values.data(), len ? &pol[0][0][0] : nullptr, len);
        return std::make_pair(std::move(values), std::move(pol));
    }
    void eval_raw(int player, float* values, float* policies, int len) override { cb_(player, values, policies, len); }

   private:
    Callback cb_;
};

class AlphaZero : public Policy {
   public:
    explicit AlphaZero(std::unique_ptr<Evaluator> evaluator) : AlphaZero(std::move(evaluator), HyperParameter()) {}  // AlphaZero::new
    AlphaZero(std::unique_ptr<Evaluator> evaluator, HyperParameter param, uint32_t seed = 0xC6, int device = 0, int log2_records = 18)
        : evaluator_(std::move(evaluator)), param_(param) {  // AlphaZero::with_param
        c6::Config cfg;
        cfg.num_simulation = param.num_simulation, cfg.epsilon = param.epsilon, cfg.dirichlet_alpha = param.dirichlet_alpha;
        cfg.c_puct = param.c_puct, cfg.seed = seed, cfg.device = device, cfg.slots = 1, cfg.log2_records = log2_records;
        cfg.evaluator = evaluator_->device_kind() >= 0 ? evaluator_->device_kind() : (int)c6::EVAL_CALLBACK;
        engine_.reset(new c6::Engine(cfg));
        for (auto& r : seen_) r.fill(Player::None);
    }

    std::optional<Pos> next(const Game& game) override {
        if (!started_) {
            engine_->start(0, 1);
            started_ = true;
        }
        // stones the engine has not seen yet (placed by another policy): Game::set order within one colour
        // does not matter for the tree, colours follow the stone count
        const Board& b = game.get_board();
        for (int pass = 0; pass < 2; ++pass)
            for (size_t i = 0; i < BOARD_SIZE; ++i)
                for (size_t j = 0; j < BOARD_SIZE; ++j)
                    if (b[i][j] != seen_[i][j]) {
                        if (seen_[i][j] != Player::None) throw std::runtime_error("AlphaZero::next: the game went backwards");
                        if ((int)b[i][j] != c6::turn_of(stones_)) continue;  // the other colour's stone comes first
                        engine_->external_move(0, (int)i, (int)j);
                        seen_[i][j] = b[i][j];
                        stones_ += 1;
                    }
        for (size_t i = 0; i < BOARD_SIZE; ++i)
            for (size_t j = 0; j < BOARD_SIZE; ++j)
                if (b[i][j] != seen_[i][j]) throw std::runtime_error("AlphaZero::next: board does not follow the turn rule");
        c6::GameState before = engine_->slot_state(0);
        if (before.status != c6::ST_ACTIVE) return std::nullopt;
        if (evaluator_->device_kind() >= 0) {
            engine_->step_device();  // num_simulation searches and the stone, one launch
        } else {
            current() = evaluator_.get();
            engine_->waves_callback(&AlphaZero::trampoline, std::max(param_.num_simulation, 1));
            current() = nullptr;
        }
        std::vector<c6::GameRecord> rec;
        c6::GameState after = engine_->slot_state(0);
        int n_moves = after.n_moves;
        const uint8_t* moves = after.moves;
        if (after.status != c6::ST_ACTIVE) {  // the game ended with this stone: the slot was retired, the record has it
            rec = engine_->records();
            n_moves = rec[0].n_moves, moves = rec[0].moves;
        }
        if (n_moves <= stones_) return std::nullopt;  // no children: policy() -> None (:433-452)
        const int cell = moves[n_moves - 1];
        seen_[cell / 15][cell % 15] = player_from(c6::turn_of(stones_));
        stones_ += 1;
        return Pos((size_t)(cell / 15), (size_t)(cell % 15));
    }

    c6::Engine& engine() { return *engine_; }

   private:
    static Evaluator*& current() {
        static thread_local Evaluator* e = nullptr;
        return e;
    }
    static void trampoline(int player, float* values, float* policies, int len) { current()->eval_raw(player, values, policies, len); }

    std::unique_ptr<Evaluator> evaluator_;
    HyperParameter param_;
    std::unique_ptr<c6::Engine> engine_;
    Board seen_;
    int stones_ = 0;
    bool started_ = false;
};

// AsyncAgent (agent/async_agent/mod.rs:61-137): `num` games, a fresh policy from the generator for each, every game an
// `Agent::new(&mut policy).play()` of its own on a pool of host threads, results in completion order.  This is the
// reference's only parallelism; it is kept for arbitrary policies.  For the policy this library accelerates use
// AlphaZeroAsyncAgent below: the same `num` games resident on the GPU at once, one warp each.
template <class PolicyGen>
class AsyncAgent {
   public:
    explicit AsyncAgent(PolicyGen gen, bool debug = false) : gen_(std::move(gen)), debug_(debug) {}  // AsyncAgent::new
    static AsyncAgent debug(PolicyGen gen) { return AsyncAgent(std::move(gen), true); }              // AsyncAgent::debug
    // Panics (throws) if a game returns Err, where the reference's `recv().unwrap()` blocks forever (:129-134)
    std::vector<PlayResult> run(int num) {
        std::vector<PlayResult> results;
        std::mutex m;
        std::string first_error;
        std::atomic<int> next{0};
        const int n_threads = std::max(1, std::min<int>(num, (int)std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&] {
                for (;;) {
                    const int id = next.fetch_add(1);
                    if (id >= num) return;
                    auto policy = gen_();
                    const auto t0 = std::chrono::steady_clock::now();
                    PlayResult res;
                    std::string err;
                    try {
                        err = Agent(*policy).play(res);
                    } catch (const std::exception& e) {
                        err = e.what();
                    }
                    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    std::lock_guard<std::mutex> l(m);
                    if (err.empty()) results.push_back(std::move(res));
                    else if (first_error.empty()) first_error = err;
                    if (debug_) std::printf("run: %d, elapsed %.3fs\n", id, el);
                }
            });
        for (auto& th : pool) th.join();
        if (!first_error.empty()) throw std::runtime_error("AsyncAgent::run: " + first_error);
        return results;
    }

   private:
    PolicyGen gen_;
    bool debug_;
};
template <class PolicyGen>
AsyncAgent<PolicyGen> make_async_agent(PolicyGen gen) { return AsyncAgent<PolicyGen>(std::move(gen)); }

// `AsyncAgent::new(|| AlphaZero::with_param(Box::new(evaluator), param)).run(num)` (what cpp_self_play builds,
// CB/libcppconnect6/src/lib.rs:110-123) as one batched search: all `num` games resident on the GPU (up to max_concurrent;
// more are queued and refill finished slots), the evaluator called once per side to move and simulation wave with the leaves
// of every game, or served by its device twin.  Results in game-id order.
class AlphaZeroAsyncAgent {
   public:
    AlphaZeroAsyncAgent(std::shared_ptr<Evaluator> evaluator, HyperParameter param, uint32_t seed = 0xC6, int device = 0,
                        int max_concurrent = 8192, int log2_records = 18)
        : evaluator_(std::move(evaluator)), param_(param), seed_(seed), device_(device), max_concurrent_(max_concurrent), log2_records_(log2_records) {}
    std::vector<PlayResult> run(int num) {
        std::vector<PlayResult> out;
        if (num <= 0) return out;
        c6::Config cfg;
        cfg.num_simulation = param_.num_simulation, cfg.epsilon = param_.epsilon, cfg.dirichlet_alpha = param_.dirichlet_alpha;
        cfg.c_puct = param_.c_puct, cfg.seed = seed_, cfg.device = device_, cfg.slots = std::min(num, max_concurrent_);
        cfg.log2_records = log2_records_;
        const int kind = evaluator_->device_kind();
        cfg.evaluator = kind >= 0 ? kind : (int)c6::EVAL_CALLBACK;
        c6::Engine engine(cfg);
        engine.start(0, num);
        if (kind >= 0) {
            engine.run_device_to_end();
        } else {
            current() = evaluator_.get();
            try {
                engine.run_callback_to_end(&AlphaZeroAsyncAgent::trampoline, false);
            } catch (...) {
                current() = nullptr;
                throw;
            }
            current() = nullptr;
        }
        for (const c6::GameRecord& r : engine.records()) {
            PlayResult res;
            Game game;
            res.winner = player_from(r.winner);
            for (int m = 0; m < r.n_moves; ++m) {
                const Pos pos((size_t)(r.moves[m] / 15), (size_t)(r.moves[m] % 15));
                res.path.push_back(Path{game.get_turn(), game.get_board(), pos});
                if (!game.set(pos).empty()) throw std::runtime_error("AlphaZeroAsyncAgent::run: illegal stone in a record");
            }
            out.push_back(std::move(res));
        }
        return out;
    }

   private:
    static Evaluator*& current() {
        static thread_local Evaluator* e = nullptr;
        return e;
    }
    static void trampoline(int player, float* values, float* policies, int len) { current()->eval_raw(player, values, policies, len); }
    std::shared_ptr<Evaluator> evaluator_;
    HyperParameter param_;
    uint32_t seed_;
    int device_, max_concurrent_, log2_records_;
};

// DefaultPolicy — pure MCTS with random rollouts (LC/policy/default_policy/mod.rs:94-271: `new()` = 50 searches
// per stone, `with_num_iter(n)`).  The GPU plays whole games (one thread per game, c6_default_policy.cu), so the
// mirror is of `Agent::new(&mut DefaultPolicy::new()).play()` repeated n times — what AsyncAgent::run does with a
// policy generator (LC/agent/async_agent/mod.rs:100-137) — not of a single `next()`.
class DefaultPolicy {
   public:
    DefaultPolicy() {}                                                       // DefaultPolicy::new
    static DefaultPolicy with_num_iter(int num_iter) {                       // DefaultPolicy::with_num_iter
        DefaultPolicy p;
        p.num_iter_ = num_iter;
        return p;
    }
    // n games, results in game-id order (game g uses the Philox streams of (seed, first_game_id + g))
    std::vector<PlayResult> play(int n_games, uint32_t seed = 0xC6, uint32_t first_game_id = 0, int device = 0) const {
        std::vector<c6::GameRecord> rec((size_t)std::max(n_games, 0));
        unsigned long long stats[8];
        c6::default_policy_play(num_iter_, seed, first_game_id, n_games, 0, device, rec.data(), nullptr, nullptr, stats);
        std::vector<PlayResult> out(rec.size());
        for (size_t g = 0; g < rec.size(); ++g) {
            Game game;
            out[g].winner = player_from(rec[g].winner);
            for (int m = 0; m < rec[g].n_moves; ++m) {
                const Pos pos((size_t)(rec[g].moves[m] / 15), (size_t)(rec[g].moves[m] % 15));
                out[g].path.push_back(Path{game.get_turn(), game.get_board(), pos});
                if (!game.set(pos).empty()) throw std::runtime_error("DefaultPolicy::play: illegal stone in a record");
            }
        }
        return out;
    }
    int num_iter() const { return num_iter_; }

   private:
    int num_iter_ = 50;
};

}  // namespace connect6
