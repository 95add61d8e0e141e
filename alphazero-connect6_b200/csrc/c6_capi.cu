// C-ABI of libcppconnect6 (see include/connect6_b200.h).
//   Part 1 mirrors Connect6/cppbind/libcppconnect6/src/lib.rs and cppbind/{rawobj,cppeval,cpp_policy,ffi_test}.
//   Part 2 exposes the engine to C / ctypes.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <vector>
#include <stdexcept>

#include "../../include/connect6_b200.h"
#include "c6_board.cuh"
#include "c6_engine.cuh"
#include "c6_rng.cuh"
#include "connect6_alphazero.hpp"
#include "connect6_host.hpp"

using namespace c6;

static_assert(sizeof(c6_path) == 912, "Path layout (connect6.hpp:26-31)");
static_assert(sizeof(c6_play_result) == 24, "PlayResult layout (connect6.hpp:33-37)");
static_assert(sizeof(c6_vec) == 16, "Vec layout (connect6.hpp:39-42)");
static_assert(sizeof(c6_record) == sizeof(GameRecord), "record layout");

static thread_local std::string g_err;

// Every c6_* entry point: exceptions stop at the C boundary, the message is kept for c6_last_error().
template <class F>
static int guarded(F&& body) {
    try {
        body();
        return 0;
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return 1;
    }
}

[[noreturn]] static void die(const char* where, const char* what) {
    // the reference unwraps and aborts across the FFI (lib.rs:40,107,168)
    std::fprintf(stderr, "libcppconnect6 (B200 engine): %s: %s\n", where, what);
    std::fflush(stderr);
    std::abort();
}

static int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
static uint32_t env_seed() {
    const char* v = std::getenv("C6_SEED");
    if (v && *v) return (uint32_t)std::strtoul(v, nullptr, 0);
    std::random_device rd;  // the reference seeds from the OS (thread_rng)
    return rd();
}

// RawPlayResult::with_result (rawobj/mod.rs:91-108): Path = board BEFORE the stone (agent_impl/mod.rs:139-143)
static c6_play_result expand_record(const GameRecord& r, c6_alloc_path alloc) {
    c6_play_result out;
    out.winner = r.winner;
    out.len = r.n_moves;
    out.paths = alloc(r.n_moves);
    int board[C6_BOARD_SIZE][C6_BOARD_SIZE];
    std::memset(board, 0, sizeof board);
    for (int m = 0; m < r.n_moves; ++m) {
        int turn = turn_of(m);
        c6_path& p = out.paths[m];
        p.turn = turn;
        std::memcpy(p.board, board, sizeof board);
        p.row = r.moves[m] / 15, p.col = r.moves[m] % 15;
        board[p.row][p.col] = turn;
    }
    return out;
}

// C6_DEVICES = "0,1,2,3" | "all": shard the games of one cpp_self_play call over several GPUs of the box by
// game id (one host thread + engine per GPU, no exchange between them).  Default: one GPU
// (C6_DEVICE), callbacks from the calling thread only.
static std::vector<int> env_devices() {
    std::vector<int> devs;
    const char* v = std::getenv("C6_DEVICES");
    if (!v || !*v) return {env_int("C6_DEVICE", 0)};
    if (std::string(v) == "all") {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) n = 1;
        for (int i = 0; i < n; ++i) devs.push_back(i);
        return devs;
    }
    for (const char* q = v; *q;) {
        char* end = nullptr;
        long d = std::strtol(q, &end, 10);
        if (end == q) break;
        devs.push_back((int)d);
        q = (*end == ',') ? end + 1 : end;
    }
    if (devs.empty()) devs.push_back(0);
    return devs;
}

// contiguous block of game ids for shard r of w: sizes differ by at most one (pyconnect6.sharded.shard_range)
static void shard_range(int64_t n, int r, int w, int64_t& lo, int64_t& cnt) {
    const int64_t base = n / w, extra = n % w;
    lo = r * base + std::min<int64_t>(r, extra);
    cnt = base + (r < extra ? 1 : 0);
}

template <class Run>  // Run(device, first_id, count, GameRecord* out) on one thread per shard; rethrows the first failure
static void run_sharded(const std::vector<int>& devices, uint32_t first, int64_t n, GameRecord* out, Run&& run) {
    const int w = (int)std::min<int64_t>((int64_t)devices.size(), std::max<int64_t>(n, 1));
    if (w <= 1) return run(devices[0], first, n, out);
    std::vector<std::thread> th;
    std::vector<std::string> errs((size_t)w);
    for (int r = 0; r < w; ++r)
        th.emplace_back([&, r] {
            int64_t lo, cnt;
            shard_range(n, r, w, lo, cnt);
            try {
                if (cnt > 0) run(devices[r], first + (uint32_t)lo, cnt, out + lo);
            } catch (const std::exception& e) {
                errs[r] = e.what();
            }
        });
    for (auto& t : th) t.join();
    for (auto& e : errs)
        if (!e.empty()) throw std::runtime_error(e);
}

static void print_game_debug(const GameRecord& r) {
    // Agent::debug transcript (agent_impl/mod.rs:127-129,147-161,173-175); per-move time is not
    // observable for a batched search, printed as 0.0.
    connect6::Game g;
    for (int m = 0; m < r.n_moves; ++m) {
        g.print(stdout);
        connect6::SetResult sr;
        g.set({(size_t)(r.moves[m] / 15), (size_t)(r.moves[m] % 15)}, &sr);
        std::printf("%s (%c, %c), remain %d, 0.0 elapsed\n", sr.player == connect6::Player::Black ? "Black" : "White",
                    (char)(sr.position.first + 0x61), (char)(sr.position.second + 0x41), sr.num_remain);
    }
    g.print(stdout);
    std::fflush(stdout);
}

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t _e = (call);                                                  \
        if (_e != cudaSuccess) throw std::runtime_error(cudaGetErrorString(_e));  \
    } while (0)

// Owning handles: CK() throws, and the C boundary must not leak device / pinned memory, streams or events on the way out.
template <class T>
struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    explicit DevBuf(size_t n) { CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T))); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};
struct PinnedBuf {
    void* p = nullptr;
    explicit PinnedBuf(size_t n) { CK(cudaMallocHost(&p, n)); }
    ~PinnedBuf() { cudaFreeHost(p); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
};
struct StreamH {
    cudaStream_t s = nullptr;
    StreamH() { CK(cudaStreamCreate(&s)); }
    ~StreamH() { cudaStreamDestroy(s); }
    StreamH(const StreamH&) = delete;
    StreamH& operator=(const StreamH&) = delete;
};
struct EventH {
    cudaEvent_t e = nullptr;
    EventH() { CK(cudaEventCreate(&e)); }
    ~EventH() { cudaEventDestroy(e); }
    EventH(const EventH&) = delete;
    EventH& operator=(const EventH&) = delete;
};

extern "C" {

// ------------------------------------------------------------------------------------------------
// Part 1
// ------------------------------------------------------------------------------------------------
c6_vec cpp_self_play(c6_callback callback, c6_alloc_path alloc_path, c6_alloc_result alloc_result, int num_simulation,
                     float epsilon, double dirichlet_alpha, float c_puct, bool debug, int num_game_thread) {
    if (!callback) die("cpp_self_play", "null evaluator callback");
    const int n = num_game_thread;  // the reference runs exactly this many games, one task each (lib.rs:110-123)
    c6_vec out;
    out.len = n > 0 ? n : 0;
    out.vec = alloc_result(out.len);
    if (n <= 0) return out;
    try {
        Config base;
        base.num_simulation = num_simulation, base.epsilon = epsilon, base.dirichlet_alpha = dirichlet_alpha, base.c_puct = c_puct;
        base.seed = env_seed();
        base.log2_records = env_int("C6_LOG2_RECORDS", 18);
        base.evaluator = EVAL_CALLBACK;
        base.tie_break = env_int("C6_TIE_BREAK", 1);
        base.virtual_loss = env_int("C6_VIRTUAL_LOSS", 1);  // opt-in multi-leaf mode (not the reference's sequential search)
        const int max_conc = env_int("C6_MAX_CONCURRENT", 8192);
        std::vector<GameRecord> rec((size_t)n);
        auto t0 = std::chrono::steady_clock::now();
        run_sharded(env_devices(), 0, n, rec.data(), [&](int device, uint32_t first, int64_t cnt, GameRecord* dst) {
            Config cfg = base;
            cfg.device = device;
            cfg.slots = (int)std::min<int64_t>(cnt, max_conc);
            Engine eng(cfg);
            eng.start(first, cnt);
            eng.run_callback_to_end(callback, debug);
            std::vector<GameRecord> r = eng.records();
            std::copy(r.begin(), r.end(), dst);
        });
        double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (int i = 0; i < n; ++i) {
            if (debug && n == 1) print_game_debug(rec[i]);
            if (debug && n != 1) std::printf("run: %d, elapsed %.3fs\n", i, el);  // async_agent/mod.rs:117-124
            out.vec[i] = expand_record(rec[i], alloc_path);
        }
    } catch (const std::exception& e) {
        die("cpp_self_play", e.what());
    }
    return out;
}

// CppPolicy + Agent::play (cpp_policy/mod.rs:22-36, agent_impl/mod.rs:121-178): no search, host loop.
c6_vec cpp_play(c6_policy_callback callback, c6_alloc_path alloc_path, c6_alloc_result alloc_result, bool debug,
                int num_game_thread) {
    if (!callback) die("cpp_play", "null policy callback");
    const int n = num_game_thread;
    c6_vec out;
    out.len = n > 0 ? n : 0;
    out.vec = alloc_result(out.len);
    for (int i = 0; i < out.len; ++i) {
        connect6::CppPolicy pol(callback);
        connect6::Agent agent(pol, debug && n == 1);
        connect6::PlayResult res;
        std::string err = agent.play(res);
        if (!err.empty()) die("cpp_play", err.c_str());
        c6_play_result& r = out.vec[i];
        r.winner = (int)res.winner;
        r.len = (int)res.path.size();
        r.paths = alloc_path(r.len);
        for (int m = 0; m < r.len; ++m) {
            c6_path& p = r.paths[m];
            p.turn = (int)res.path[m].turn;
            for (int a = 0; a < 15; ++a)
                for (int b = 0; b < 15; ++b) p.board[a][b] = (int)res.path[m].board[a][b];
            p.row = (int)res.path[m].pos.first, p.col = (int)res.path[m].pos.second;
        }
    }
    return out;
}

// MultiPolicy(AlphaZero as Black on the GPU, IoPolicy on stdin as White) under Agent::debug (lib.rs:140-169)
c6_play_result cpp_play_with(c6_callback callback, c6_alloc_path alloc_path, int num_simulation, float epsilon,
                             double dirichlet_alpha, float c_puct) {
    if (!callback) die("cpp_play_with", "null evaluator callback");
    c6_play_result r;
    r.winner = 0, r.len = 0, r.paths = nullptr;
    try {
        connect6::HyperParameter hp;
        hp.num_simulation = num_simulation, hp.epsilon = epsilon, hp.dirichlet_alpha = dirichlet_alpha, hp.c_puct = c_puct;
        connect6::AlphaZero black(std::unique_ptr<connect6::Evaluator>(new connect6::CppEval(callback)), hp, env_seed(),
                                  env_int("C6_DEVICE", 0), env_int("C6_LOG2_RECORDS", 18));
        connect6::IoPolicy white(stdin, stdout);
        connect6::MultiPolicy multi(black, white);
        connect6::Agent agent(multi, true);
        connect6::PlayResult res;
        std::string err = agent.play(res);
        if (!err.empty()) die("cpp_play_with", err.c_str());
        r.winner = (int)res.winner;
        r.len = (int)res.path.size();
        r.paths = alloc_path(r.len);
        for (int m = 0; m < r.len; ++m) {
            c6_path& p = r.paths[m];
            p.turn = (int)res.path[m].turn;
            for (int a = 0; a < 15; ++a)
                for (int b = 0; b < 15; ++b) p.board[a][b] = (int)res.path[m].board[a][b];
            p.row = (int)res.path[m].pos.first, p.col = (int)res.path[m].pos.second;
        }
    } catch (const std::exception& e) {
        die("cpp_play_with", e.what());
    }
    return r;
}

c6_path test_new_raw_path(void) {  // ffi_test/mod.rs:13-15
    c6_path p;
    std::memset(&p, 0, sizeof p);
    return p;
}
c6_path test_with_raw_path(void) {  // ffi_test/mod.rs:19-34
    c6_path p;
    for (int i = 0; i < 15; ++i)
        for (int j = 0; j < 15; ++j) p.board[i][j] = (i * 15 + j) % 3 - 1;
    p.turn = 1, p.row = 0, p.col = 15 % 5 + 1;
    return p;
}
c6_path test_echo_raw_path(int turn, int* board, int row, int col) {  // :38-57
    c6_path p;
    std::memcpy(p.board, board, sizeof p.board);
    p.turn = turn, p.row = row, p.col = col;
    return p;
}
c6_play_result test_with_raw_play_result(c6_alloc_path allocator) {  // :61-86
    c6_play_result r;
    r.winner = -1, r.len = 10, r.paths = allocator(10);
    int player = -1;
    for (int i = 0; i < 10; ++i) {
        c6_path& p = r.paths[i];
        std::memset(&p, 0, sizeof p);
        for (int j = 0; j < i + 1; ++j) p.board[j][j] = (i + j) % 3 - 1;
        p.turn = player, p.row = i, p.col = i + 1;
        player = -player;
    }
    return r;
}
c6_play_result test_echo_raw_play_result(int winner, c6_path* path, int len, c6_alloc_path allocator) {  // :90-121
    auto norm = [](int v) { return v == -1 ? -1 : (v == 1 ? 1 : 0); };  // Player::from
    c6_play_result r;
    r.winner = norm(winner), r.len = len, r.paths = allocator(len);
    for (int i = 0; i < len; ++i) {
        c6_path& p = r.paths[i];
        for (int a = 0; a < 15; ++a)
            for (int b = 0; b < 15; ++b) p.board[a][b] = norm(path[i].board[a][b]);
        p.turn = norm(path[i].turn), p.row = path[i].row, p.col = path[i].col;
    }
    return r;
}
c6_vec_int test_with_raw_vec(c6_alloc_int allocator) {  // :125-129
    c6_vec_int v;
    v.len = 6, v.vec = allocator(6);
    for (int i = 0; i < 6; ++i) v.vec[i] = i;
    return v;
}
c6_vec_int test_echo_raw_vec(int* ptr, int len, c6_alloc_int allocator) {  // :133-145
    c6_vec_int v;
    v.len = len, v.vec = allocator(len);
    for (int i = 0; i < len; ++i) v.vec[i] = ptr[i];
    return v;
}
// CppEval::eval on host boards (cppeval/mod.rs:27-43): values pre-zeroed, policies = board_to_float
c6_vec_float test_echo_cppeval(int turn, int* boards, int len, c6_callback callback, c6_alloc_float allocator) {  // :149-197
    auto norm = [](int v) { return v == -1 ? -1 : (v == 1 ? 1 : 0); };
    std::vector<float> values((size_t)len, 0.f), pol((size_t)len * 225);
    for (size_t i = 0; i < pol.size(); ++i) pol[i] = (float)norm(boards[i]);
    callback(norm(turn), values.data(), pol.data(), len);
    c6_vec_float v;
    v.len = len + len * 225, v.vec = allocator(v.len);
    std::memcpy(v.vec, values.data(), (size_t)len * 4);
    std::memcpy(v.vec + len, pol.data(), pol.size() * 4);
    return v;
}
c6_vec_int test_cpp_policy(float* board, c6_policy_callback callback, c6_alloc_int allocator) {  // :201-224
    connect6::Board b;
    for (int i = 0; i < 15; ++i)
        for (int j = 0; j < 15; ++j) b[i][j] = connect6::player_from((int)board[i * 15 + j]);
    connect6::CppPolicy pol(callback);
    auto res = pol.callback(b);
    c6_vec_int v;
    v.len = res ? 2 : 0, v.vec = allocator(v.len);
    if (res) v.vec[0] = (int)res->first, v.vec[1] = (int)res->second;
    return v;
}

// ------------------------------------------------------------------------------------------------
// Part 2
// ------------------------------------------------------------------------------------------------
struct c6_engine {
    std::unique_ptr<Engine> eng;
    int64_t total = 0;
};

void c6_default_config(c6_config* cfg) {
    Config d;
    cfg->num_simulation = d.num_simulation, cfg->epsilon = d.epsilon, cfg->dirichlet_alpha = d.dirichlet_alpha;
    cfg->c_puct = d.c_puct, cfg->seed = d.seed, cfg->device = 0, cfg->slots = 1, cfg->log2_records = d.log2_records;
    cfg->evaluator = C6_EVAL_CALLBACK, cfg->trace = 0, cfg->max_moves = 0, cfg->tie_break = d.tie_break, cfg->virtual_loss = 1;
}
const char* c6_last_error(void) { return g_err.c_str(); }
int c6_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}
static Config to_config(const c6_config* c) {
    Config cfg;
    cfg.num_simulation = c->num_simulation, cfg.epsilon = c->epsilon, cfg.dirichlet_alpha = c->dirichlet_alpha;
    cfg.c_puct = c->c_puct, cfg.seed = c->seed, cfg.device = c->device, cfg.slots = c->slots;
    cfg.log2_records = c->log2_records, cfg.evaluator = c->evaluator, cfg.trace = c->trace, cfg.max_moves = c->max_moves;
    cfg.tie_break = c->tie_break, cfg.virtual_loss = c->virtual_loss;
    return cfg;
}
c6_engine* c6_engine_create(const c6_config* cfg) {
    try {
        auto* e = new c6_engine();
        e->eng.reset(new Engine(to_config(cfg)));
        return e;
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return nullptr;
    }
}
void c6_engine_destroy(c6_engine* e) { delete e; }
int c6_engine_start(c6_engine* e, uint32_t first, int64_t total) {
    return guarded([&] {
        e->total = total;
        e->eng->start(first, total);
    });
}
int c6_engine_step_device(c6_engine* e) { return guarded([&] { e->eng->step_device(); }); }
int c6_engine_wave_callback(c6_engine* e, c6_callback cb) { return guarded([&] { e->eng->wave_callback(cb); }); }
int c6_engine_waves_device_callback(c6_engine* e, c6_device_callback cb, void* user, int waves) {
    return guarded([&] {
        e->eng->waves_device_callback(cb, user, waves);
    });
}
int c6_engine_run_device_callback(c6_engine* e, c6_device_callback cb, void* user) {
    return guarded([&] {
        e->eng->run_device_callback_to_end(cb, user);
    });
}
int c6_engine_run_device(c6_engine* e) { return guarded([&] { e->eng->run_device_to_end(); }); }
int c6_engine_run_callback(c6_engine* e, c6_callback cb) { return guarded([&] { e->eng->run_callback_to_end(cb, false); }); }
int c6_engine_sync(c6_engine* e) { return guarded([&] { e->eng->sync(); }); }
void c6_engine_cancel(c6_engine* e) { e->eng->cancel(); }
int c6_engine_external_move(c6_engine* e, int slot, int row, int col) { return guarded([&] { e->eng->external_move(slot, row, col); }); }
int c6_engine_waves_callback(c6_engine* e, c6_callback cb, int waves) { return guarded([&] { e->eng->waves_callback(cb, waves); }); }
int c6_engine_timed_steps(c6_engine* e, int steps, float* ms) { return guarded([&] { *ms = e->eng->timed_steps_device(steps); }); }
int c6_engine_timed_waves(c6_engine* e, c6_callback cb, int waves, float* ms, double* h2d, double* d2h) {
    return guarded([&] {
        *ms = e->eng->timed_waves_callback(cb, waves, h2d, d2h);
    });
}
int c6_engine_host_times(c6_engine* e, double* wait_seconds, double* callback_seconds) {
    return guarded([&] {
        *wait_seconds = e->eng->host_wait_seconds();
        *callback_seconds = e->eng->host_callback_seconds();
    });
}
long long c6_engine_host_blocked_waits(c6_engine* e) { return e->eng->host_blocked_waits(); }
int c6_engine_kernel_times(c6_engine* e, double* sims_kernel_ms, long long* sims_kernel_launches) {
    return guarded([&] {
        *sims_kernel_ms = e->eng->sims_kernel_ms();
        *sims_kernel_launches = e->eng->sims_kernel_launches();
    });
}

void c6_release_cached_memory(void) { release_cached_memory(); }
int c6_engine_stats(c6_engine* e, c6_stats* out) {
    return guarded([&] {
        Stats s = e->eng->stats();
        out->sims = s.sims, out->sum_ck = s.sum_ck, out->sum_cleaf = s.sum_cleaf,
        out->sum_depth = s.sum_depth, out->evals = s.evals, out->gc_copied = s.gc_copied, out->games_done = s.games_done,
        out->peak_rec = s.peak_rec, out->peak_lnk = s.peak_lnk, out->kernel_launches = (uint64_t)e->eng->kernel_launches(),
        out->bytes_per_game = e->eng->bytes_per_game();
        out->degraded = s.degraded;
    });
}
int64_t c6_engine_games_done(c6_engine* e) {
    try {
        return e->eng->games_done();
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
}
int c6_engine_records(c6_engine* e, c6_record* out, int64_t n) {
    return guarded([&] {
        auto r = e->eng->records();
        if ((int64_t)r.size() < n) n = (int64_t)r.size();
        std::memcpy(out, r.data(), (size_t)n * sizeof(GameRecord));
    });
}
int c6_engine_trace(c6_engine* e, int32_t* cv, int32_t* rn, float* rw) {
    return guarded([&] {
        std::vector<int32_t> a, b;
        std::vector<float> c;
        e->eng->trace(a, b, c);
        std::memcpy(cv, a.data(), a.size() * 4);
        std::memcpy(rn, b.data(), b.size() * 4);
        std::memcpy(rw, c.data(), c.size() * 4);
    });
}
int c6_engine_trace_game(c6_engine* e, int64_t game, int32_t* cv, int32_t* rn, float* rw) {
    return guarded([&] { e->eng->trace_game(game, cv, rn, rw); });
}
long long c6_engine_validate(c6_engine* e, unsigned long long* detail) {
    try {
        return e->eng->validate(detail);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
}
long long c6_engine_debug_read(c6_engine* e, int what, int slot, void* out, long long max_bytes) {
    try {
        return (long long)e->eng->debug_read(what, slot, out, (size_t)max_bytes);
    } catch (const std::exception& ex) {
        g_err = ex.what();
        return -1;
    }
}
int c6_self_play_device(const c6_config* cfg, uint32_t first, int64_t n_games, c6_record* out, c6_stats* stats) {
    c6_engine* e = c6_engine_create(cfg);
    if (!e) return 1;
    int rc = c6_engine_start(e, first, n_games);
    if (!rc) rc = c6_engine_run_device(e);
    if (!rc && out) rc = c6_engine_records(e, out, n_games);
    if (!rc && stats) rc = c6_engine_stats(e, stats);
    c6_engine_destroy(e);
    return rc;
}

// n_games of self-play with a device evaluator, sharded by game id over `devices` (one host thread + engine per GPU)
int c6_self_play_device_sharded(const c6_config* cfg, const int* devices, int n_devices, uint32_t first, int64_t n_games,
                                c6_record* out, c6_stats* stats) {
    return guarded([&] {
        std::vector<int> devs(devices, devices + std::max(n_devices, 0));
        if (devs.empty()) devs.push_back(cfg->device);
        std::vector<GameRecord> rec((size_t)std::max<int64_t>(n_games, 0));
        std::mutex m;
        c6_stats tot;
        std::memset(&tot, 0, sizeof tot);
        run_sharded(devs, first, n_games, rec.data(), [&](int device, uint32_t f, int64_t cnt, GameRecord* dst) {
        Config c = to_config(cfg);
        c.device = device;
        c.slots = (int)std::min<int64_t>(std::max<int64_t>(cnt, 1), cfg->slots > 0 ? cfg->slots : cnt);
        Engine eng(c);
        eng.start(f, cnt);
        eng.run_device_to_end();
        std::vector<GameRecord> r = eng.records();
        std::copy(r.begin(), r.end(), dst);
        Stats s = eng.stats();
        std::lock_guard<std::mutex> l(m);
        tot.sims += s.sims, tot.sum_ck += s.sum_ck, tot.sum_cleaf += s.sum_cleaf, tot.sum_depth += s.sum_depth, tot.evals += s.evals;
        tot.gc_copied += s.gc_copied, tot.games_done += s.games_done, tot.kernel_launches += (uint64_t)eng.kernel_launches();
        tot.peak_rec = std::max(tot.peak_rec, s.peak_rec), tot.peak_lnk = std::max(tot.peak_lnk, s.peak_lnk);
        tot.bytes_per_game = eng.bytes_per_game();
        tot.degraded += s.degraded;
        });
        if (out && n_games > 0) std::memcpy(out, rec.data(), (size_t)n_games * sizeof(GameRecord));
        if (stats) *stats = tot;
    });
}

int c6_win_check_host(const uint32_t* boards, int8_t* winners, int64_t n, int device) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        if (n <= 0) return;
        DevBuf<uint32_t> db((size_t)n * 16);
        DevBuf<int8_t> dw((size_t)n);
        CK(cudaMemcpy(db.p, boards, (size_t)n * 64, cudaMemcpyHostToDevice));
        launch_win_check(db.p, dw.p, n, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(winners, dw.p, (size_t)n, cudaMemcpyDeviceToHost));
    });
}

int c6_bench_win_check(int64_t n, uint32_t seed, int iters, int device, float* ms_per_iter, uint32_t* boards_out,
                       int8_t* winners_out, int64_t n_out) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<uint32_t> db((size_t)n * 16);
        DevBuf<int8_t> dw((size_t)n);
        StreamH st;
        EventH a, b;
        launch_random_boards(db.p, n, seed, st.s);
        launch_win_check(db.p, dw.p, n, st.s);  // warm-up
        CK(cudaStreamSynchronize(st.s));
        CK(cudaEventRecord(a.e, st.s));
        for (int i = 0; i < iters; ++i) launch_win_check(db.p, dw.p, n, st.s);
        CK(cudaEventRecord(b.e, st.s));
        CK(cudaEventSynchronize(b.e));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a.e, b.e));
        *ms_per_iter = iters > 0 ? ms / iters : 0.f;
        if (n_out > n) n_out = n;
        if (boards_out && n_out > 0) CK(cudaMemcpy(boards_out, db.p, (size_t)n_out * 64, cudaMemcpyDeviceToHost));
        if (winners_out && n_out > 0) CK(cudaMemcpy(winners_out, dw.p, (size_t)n_out, cudaMemcpyDeviceToHost));
    });
}

int c6_bench_playouts(int64_t first, int64_t n, uint32_t seed, int iters, int device, float* ms_per_iter, int8_t* winners_out,
                      uint8_t* stones_out, uint32_t* boards_out, int64_t n_out, uint64_t* total_stones) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<int8_t> dw((size_t)n);
        DevBuf<uint8_t> ds((size_t)n);
        DevBuf<uint32_t> db((size_t)n * 16);
        StreamH st;
        EventH a, b;
        launch_playouts(first, n, seed, dw.p, ds.p, db.p, st.s);
        CK(cudaStreamSynchronize(st.s));
        CK(cudaEventRecord(a.e, st.s));
        for (int i = 0; i < iters; ++i) launch_playouts(first, n, seed, dw.p, ds.p, db.p, st.s);
        CK(cudaEventRecord(b.e, st.s));
        CK(cudaEventSynchronize(b.e));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, a.e, b.e));
        *ms_per_iter = iters > 0 ? ms / iters : 0.f;
        if (n_out > n) n_out = n;
        if (winners_out && n_out > 0) CK(cudaMemcpy(winners_out, dw.p, (size_t)n_out, cudaMemcpyDeviceToHost));
        if (stones_out && n_out > 0) CK(cudaMemcpy(stones_out, ds.p, (size_t)n_out, cudaMemcpyDeviceToHost));
        if (boards_out && n_out > 0) CK(cudaMemcpy(boards_out, db.p, (size_t)n_out * 64, cudaMemcpyDeviceToHost));
        if (total_stones) {
            std::vector<uint8_t> hs((size_t)n);
            CK(cudaMemcpy(hs.data(), ds.p, (size_t)n, cudaMemcpyDeviceToHost));
            uint64_t t = 0;
            for (uint8_t x : hs) t += x;
            *total_stones = t;
        }
    });
}

// Replay feed with DEVICE pointers: records (n_games x 240 B) and the exclusive prefix sum of their stone counts in HBM, the
// four float32 tensors of the reference trainer's ReplayBuffer::sample layout out (cpp_weighted/ReplayBuffer.hpp:38-65).
int c6_expand_records_device(const c6_record* d_records, const long long* d_row_offset, int n_games, float* d_winners, float* d_players,
                             float* d_boards, float* d_positions, void* stream) {
    return guarded([&] {
        launch_expand_records(reinterpret_cast<const GameRecord*>(d_records), d_row_offset, n_games, d_winners, d_players, d_boards,
                              d_positions, static_cast<cudaStream_t>(stream));
        CK(cudaGetLastError());
    });
}
// The same from host records into host arrays sized for sum(n_moves) samples (tests, small feeds); returns the sample count in *n_rows.
int c6_expand_records_host(const c6_record* records, int n_games, int device, float* winners, float* players, float* boards,
                           float* positions, long long* n_rows) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        std::vector<long long> off((size_t)std::max(n_games, 0) + 1, 0);
        for (int g = 0; g < n_games; ++g) off[g + 1] = off[g] + records[g].n_moves;
        const long long rows = off[std::max(n_games, 0)];
        if (n_rows) *n_rows = rows;
        if (n_games <= 0 || rows == 0) return;
        DevBuf<c6_record> dr((size_t)n_games);
        DevBuf<long long> doff((size_t)n_games);
        DevBuf<float> dw((size_t)rows), dp((size_t)rows), dpos((size_t)rows), db((size_t)rows * 225);
        CK(cudaMemcpy(dr.p, records, (size_t)n_games * sizeof(c6_record), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(doff.p, off.data(), (size_t)n_games * sizeof(long long), cudaMemcpyHostToDevice));
        launch_expand_records(reinterpret_cast<const GameRecord*>(dr.p), doff.p, n_games, dw.p, dp.p, db.p, dpos.p, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(winners, dw.p, (size_t)rows * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(players, dp.p, (size_t)rows * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(positions, dpos.p, (size_t)rows * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(boards, db.p, (size_t)rows * 225 * 4, cudaMemcpyDeviceToHost));
    });
}

int c6_default_policy_play(int num_iter, uint32_t seed, uint32_t first_game_id, int64_t n_games, int max_moves, int device,
                           c6_record* out, int32_t* root_visit, int32_t* root_black_win, c6_dp_stats* stats) {
    static_assert(sizeof(c6_record) == sizeof(GameRecord), "record layout");
    return guarded([&] {
        unsigned long long st[8];
        if (n_games > 0 && !out) throw std::runtime_error("c6_default_policy_play: out is NULL");
        const float ms = default_policy_play(num_iter, seed, first_game_id, (long long)n_games, max_moves, device,
        reinterpret_cast<GameRecord*>(out), root_visit, root_black_win, st);
        if (stats) {
        stats->searches = st[0], stats->rollouts = st[1], stats->rollout_stones = st[2], stats->select_steps = st[3];
        stats->children_scanned = st[4], stats->replaced = st[5], stats->nodes = st[6], stats->errors = st[7];
        stats->kernel_ms = ms;
        }
    });
}

// Host<->device copy bandwidth of this box with pinned buffers (GB/s): H2D alone, D2H alone, and both at
// once on two streams (per direction).  This is the ceiling of the host-callback evaluator path.
int c6_probe_pcie(int64_t bytes, int iters, int device, double* h2d_gbs, double* d2h_gbs, double* both_gbs) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        PinnedBuf h1((size_t)bytes), h2((size_t)bytes);
        DevBuf<char> d1((size_t)bytes), d2((size_t)bytes);
        std::memset(h1.p, 1, (size_t)bytes);
        std::memset(h2.p, 2, (size_t)bytes);
        StreamH s1, s2;
        EventH a, b, c;
        auto timed = [&](int mode) {
            float ms = 0;
            for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
                CK(cudaDeviceSynchronize());
                CK(cudaEventRecord(a.e, s1.s));
                CK(cudaStreamWaitEvent(s2.s, a.e, 0));
                for (int i = 0; i < iters; ++i) {
                    if (mode != 1) CK(cudaMemcpyAsync(d1.p, h1.p, (size_t)bytes, cudaMemcpyHostToDevice, s1.s));
                    if (mode != 0) CK(cudaMemcpyAsync(h2.p, d2.p, (size_t)bytes, cudaMemcpyDeviceToHost, s2.s));
                }
                CK(cudaEventRecord(c.e, s2.s));
                CK(cudaStreamWaitEvent(s1.s, c.e, 0));
                CK(cudaEventRecord(b.e, s1.s));
                CK(cudaEventSynchronize(b.e));
                CK(cudaEventElapsedTime(&ms, a.e, b.e));
            }
            return (double)bytes * iters / (ms * 1e-3) / 1e9;
        };
        *h2d_gbs = timed(0);
        *d2h_gbs = timed(1);
        *both_gbs = timed(2);
    });
}
int c6_probe_gamma(double shape, uint32_t k0, uint32_t k1, float* out, int n, int device) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<float> d((size_t)n);
        launch_probe_gamma(shape, k0, k1, d.p, n, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(out, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    });
}
int c6_probe_math32(const float* xs, float* lg2s, float* ex2s, int n, int device) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<float> dx((size_t)n), dl((size_t)n), de((size_t)n);
        CK(cudaMemcpy(dx.p, xs, (size_t)n * 4, cudaMemcpyHostToDevice));
        launch_probe_math32(dx.p, dl.p, de.p, n, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(lg2s, dl.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ex2s, de.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    });
}
int c6_probe_tie(uint32_t seed, uint32_t game, uint32_t count, uint32_t* out, int n, int device) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<uint32_t> d((size_t)n);
        launch_probe_tie(seed, game, count, d.p, n, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(out, d.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    });
}
int c6_probe_math(const double* xs, double* logs, double* exps, int n, int device) {
    return guarded([&] {
        CK(cudaSetDevice(device));
        DevBuf<double> dx((size_t)n), dl((size_t)n), de((size_t)n);
        CK(cudaMemcpy(dx.p, xs, (size_t)n * 8, cudaMemcpyHostToDevice));
        launch_probe_math(dx.p, dl.p, de.p, n, 0);
        CK(cudaGetLastError());
        CK(cudaMemcpy(logs, dl.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(exps, de.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// Sample host evaluator (the twin of Connect6/sample/cpp/main.cpp:6-28's random `callback`), used by
// bench.py's end-to-end leg and the samples: U(-1,1) values and policies.  Host code on purpose —
// it stands for the user's evaluator on the far side of the C-ABI.
// ------------------------------------------------------------------------------------------------
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace {
// Minimal persistent fork-join pool for the sample evaluators (a user evaluator would bring its own).  Workers poll the
// generation counter for the host's polling budget (c6::host_spin_budget_us(): 200 us with cores to spare, 5 us when
// several ranks share the box) and then sleep on a condition variable.
class ForkJoin {
   public:
    explicit ForkJoin(int n) : n_(n), spin_us_(env_int("C6_SAMPLE_SPIN_US", c6::host_spin_budget_us())) {
        for (int t = 1; t < n_; ++t) th_.emplace_back([this, t] { loop(t); });
    }
    ~ForkJoin() {
        {
            std::lock_guard<std::mutex> l(m_);
            stop_ = true;
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& x : th_) x.join();
    }
    int size() const { return n_; }
    void run(const std::function<void(int)>& f) {
        f_ = &f;
        left_.store(n_ - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> l(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
        f(0);
        const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us_);
        while (left_.load(std::memory_order_acquire) != 0) {
            if (std::chrono::steady_clock::now() < until) cpu_relax(); else std::this_thread::yield();
        }
    }

   private:
    static void cpu_relax() {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    void loop(int t) {
        unsigned long seen = 0;
        for (;;) {
            const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us_);
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (std::chrono::steady_clock::now() < until) {
                    cpu_relax();
                } else {
                    std::unique_lock<std::mutex> l(m_);
                    sleepers_.fetch_add(1, std::memory_order_release);
                    cv_.wait(l, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                    sleepers_.fetch_sub(1, std::memory_order_release);
                }
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            (*f_)(t);
            left_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_, spin_us_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(int)>* f_ = nullptr;
    std::atomic<unsigned long> gen_{0};
    std::atomic<int> left_{0}, sleepers_{0};
    bool stop_ = false;
};

// the pool shared by the sample evaluators: at most 4 threads (more compete with the engine's host thread and the CUDA
// driver's own on a 16-vCPU box, profiles/e2e4, e2e5), and never more than this rank's share of the host cores
ForkJoin& sample_pool() {
    static ForkJoin pool(env_int("C6_SAMPLE_THREADS", std::max(1, std::min(c6::host_cores_per_rank(), 4))));
    return pool;
}
std::mutex& sample_pool_mutex() {  // the pool is shared: callers from several engines take turns
    static std::mutex m;
    return m;
}
}  // namespace

// U[-1,1) values and policies from a counter hash of the flat element index (two 32-bit multiplies per float,
// no carried dependence).  The policy buffer is write-only for a random evaluator: the 8-wide body uses
// non-temporal stores, so the fill costs one pass of host memory traffic instead of two (no read for ownership).
static inline uint32_t sample_mix(uint32_t x) {
    x *= 0xC2B2AE3Du, x ^= x >> 15, x *= 0x85EBCA77u, x ^= x >> 13;
    return x;
}
static inline float sample_unit(uint32_t x) { return (float)(int)(x >> 8) * (1.0f / 8388608.0f) - 1.0f; }
static void sample_fill_scalar(uint32_t base, float* out, size_t lo, size_t hi) {
    for (size_t e = lo; e < hi; ++e) out[e] = sample_unit(sample_mix(base + (uint32_t)e));
}
#if defined(__GNUC__) && defined(__x86_64__)
#include <immintrin.h>
template <bool NT>
__attribute__((target("avx2")))
static void sample_fill_avx2(uint32_t base, float* out, size_t lo, size_t hi) {
    size_t e = lo;
    for (; e < hi && ((uintptr_t)(out + e) & 31u); ++e) out[e] = sample_unit(sample_mix(base + (uint32_t)e));
    const __m256i step = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7);
    const __m256i m1 = _mm256_set1_epi32((int)0xC2B2AE3Du), m2 = _mm256_set1_epi32((int)0x85EBCA77u);
    const __m256 scale = _mm256_set1_ps(1.0f / 8388608.0f), one = _mm256_set1_ps(1.0f);
    for (; e + 8 <= hi; e += 8) {
        __m256i x = _mm256_add_epi32(_mm256_set1_epi32((int)(base + (uint32_t)e)), step);
        x = _mm256_mullo_epi32(x, m1);
        x = _mm256_xor_si256(x, _mm256_srli_epi32(x, 15));
        x = _mm256_mullo_epi32(x, m2);
        x = _mm256_xor_si256(x, _mm256_srli_epi32(x, 13));
        const __m256 f = _mm256_sub_ps(_mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_srli_epi32(x, 8)), scale), one);
        if (NT) _mm256_stream_ps(out + e, f); else _mm256_store_ps(out + e, f);
    }
    for (; e < hi; ++e) out[e] = sample_unit(sample_mix(base + (uint32_t)e));
    if (NT) _mm_sfence();
}
// the same with 512-bit registers: one non-temporal store writes a whole cache line
template <bool NT>
__attribute__((target("avx512f")))
static void sample_fill_avx512(uint32_t base, float* out, size_t lo, size_t hi) {
    size_t e = lo;
    for (; e < hi && ((uintptr_t)(out + e) & 63u); ++e) out[e] = sample_unit(sample_mix(base + (uint32_t)e));
    const __m512i step = _mm512_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15);
    const __m512i m1 = _mm512_set1_epi32((int)0xC2B2AE3Du), m2 = _mm512_set1_epi32((int)0x85EBCA77u);
    const __m512 scale = _mm512_set1_ps(1.0f / 8388608.0f), one = _mm512_set1_ps(1.0f);
    for (; e + 16 <= hi; e += 16) {
        __m512i x = _mm512_add_epi32(_mm512_set1_epi32((int)(base + (uint32_t)e)), step);
        x = _mm512_mullo_epi32(x, m1);
        x = _mm512_xor_si512(x, _mm512_srli_epi32(x, 15));
        x = _mm512_mullo_epi32(x, m2);
        x = _mm512_xor_si512(x, _mm512_srli_epi32(x, 13));
        const __m512 f = _mm512_sub_ps(_mm512_mul_ps(_mm512_cvtepi32_ps(_mm512_srli_epi32(x, 8)), scale), one);
        if (NT) _mm512_stream_ps(out + e, f); else _mm512_store_ps(out + e, f);
    }
    for (; e < hi; ++e) out[e] = sample_unit(sample_mix(base + (uint32_t)e));
    if (NT) _mm_sfence();
}
// policy = f(board): reads every board cell and overwrites it in place (cached read-modify-write: the evaluator of a
// real network reads its input).  p = 1/2 - board/4 + cell pattern/1024 — cheap on purpose: the leg measures the path.
__attribute__((target("avx2")))
static void reading_fill_avx2(float* pol, size_t lo, size_t hi, float* acc_out) {
    size_t e = lo;
    float acc = 0.f;
    for (; e < hi && ((uintptr_t)(pol + e) & 31u); ++e) acc += pol[e], pol[e] = 0.5f - 0.25f * pol[e] + (float)(e & 1023u) * (1.0f / 1048576.0f);
    __m256 vacc = _mm256_setzero_ps();
    const __m256 half = _mm256_set1_ps(0.5f), q = _mm256_set1_ps(-0.25f), sc = _mm256_set1_ps(1.0f / 1048576.0f);
    const __m256i step = _mm256_setr_epi32(0, 1, 2, 3, 4, 5, 6, 7), mask = _mm256_set1_epi32(1023);
    for (; e + 8 <= hi; e += 8) {
        const __m256 b = _mm256_load_ps(pol + e);
        vacc = _mm256_add_ps(vacc, b);
        const __m256 pat = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_and_si256(_mm256_add_epi32(_mm256_set1_epi32((int)(uint32_t)e), step), mask)), sc);
        _mm256_store_ps(pol + e, _mm256_add_ps(_mm256_add_ps(half, _mm256_mul_ps(q, b)), pat));
    }
    float tmp[8];
    _mm256_storeu_ps(tmp, vacc);
    for (float t : tmp) acc += t;
    for (; e < hi; ++e) acc += pol[e], pol[e] = 0.5f - 0.25f * pol[e] + (float)(e & 1023u) * (1.0f / 1048576.0f);
    *acc_out = acc;
}
static const bool g_have_avx2 = __builtin_cpu_supports("avx2");
static const bool g_have_avx512 = __builtin_cpu_supports("avx512f");
#else
static const bool g_have_avx2 = false, g_have_avx512 = false;
template <bool NT> static void sample_fill_avx2(uint32_t, float*, size_t, size_t) {}
template <bool NT> static void sample_fill_avx512(uint32_t, float*, size_t, size_t) {}
static void reading_fill_avx2(float*, size_t, size_t, float*) {}
#endif
static void reading_fill_scalar(float* pol, size_t lo, size_t hi, float* acc_out) {
    float acc = 0.f;
    for (size_t e = lo; e < hi; ++e) acc += pol[e], pol[e] = 0.5f - 0.25f * pol[e] + (float)(e & 1023u) * (1.0f / 1048576.0f);
    *acc_out = acc;
}
static void sample_fill(uint32_t base, float* values, float* policies, int lo, int hi) {
    for (int n = lo; n < hi; ++n) values[n] = sample_unit(sample_mix((base ^ 0x5bd1e995u) + (uint32_t)n));
    // runtime dispatch: a vector body must not run on an x86-64 host without its instruction set.  C6_SAMPLE_ISA = avx512 |
    // avx2 | scalar narrows the choice, C6_SAMPLE_NT=0 uses cached stores (tuning knobs of the sample evaluator only).
    using Fill = void (*)(uint32_t, float*, size_t, size_t);
    static const Fill fill = [] {
        const char* isa = std::getenv("C6_SAMPLE_ISA");
        const std::string want = isa ? isa : "";
        const bool nt = env_int("C6_SAMPLE_NT", 1) != 0;
        if (g_have_avx512 && (want.empty() || want == "avx512")) return nt ? (Fill)sample_fill_avx512<true> : (Fill)sample_fill_avx512<false>;
        if (g_have_avx2 && want != "scalar") return nt ? (Fill)sample_fill_avx2<true> : (Fill)sample_fill_avx2<false>;
        return (Fill)sample_fill_scalar;
    }();
    fill(base * 0x9E3779B1u, policies, (size_t)lo * 225, (size_t)hi * 225);
}

extern "C" void c6_sample_random_callback(int player, float* values, float* policies, int len) {
    static std::atomic<uint32_t> call_no{0};
    const uint32_t base = (call_no.fetch_add(1) * 0x9E3779B9u) ^ ((uint32_t)(player + 2) << 28);
    static const int pool_min = env_int("C6_SAMPLE_POOL_MIN", 512);
    if (len < pool_min) return sample_fill(base, values, policies, 0, len);
    ForkJoin& pool = sample_pool();
    std::lock_guard<std::mutex> l(sample_pool_mutex());
    const int nt = pool.size();
    pool.run([&](int t) { sample_fill(base, values, policies, (int)((long long)len * t / nt), (int)((long long)len * (t + 1) / nt)); });
}

// The second sample evaluator: every board is read (value = mean stone colour of the board as seen by `player`, priors a
// function of the cell) and overwritten in place.
extern "C" void c6_sample_reading_callback(int player, float* values, float* policies, int len) {
    auto fill = [&](int lo, int hi) {
        for (int n = lo; n < hi; ++n) {
            float acc = 0.f;
            (g_have_avx2 ? reading_fill_avx2 : reading_fill_scalar)(policies, (size_t)n * 225, (size_t)(n + 1) * 225, &acc);
            values[n] = (float)player * acc * (1.0f / 225.0f);
        }
    };
    static const int pool_min = env_int("C6_SAMPLE_POOL_MIN", 512);
    if (len < pool_min) return fill(0, len);
    ForkJoin& pool = sample_pool();
    std::lock_guard<std::mutex> l(sample_pool_mutex());
    const int nt = pool.size();
    pool.run([&](int t) { fill((int)((long long)len * t / nt), (int)((long long)len * (t + 1) / nt)); });
}

// Pin the calling thread (and the threads it creates afterwards: the sample pool, the engine's launcher) to the CPUs of the
// NUMA node the GPU hangs off, so that pinned evaluator buffers are first-touched next to it.  Returns the node, or -1 when
// the topology is not exposed (single node, containers without /sys PCI NUMA data): then nothing is changed.
#include <sched.h>
#include <fstream>
extern "C" int c6_bind_host_to_device_numa(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) return -1;
    for (char* c = bus; *c; ++c) *c = (char)std::tolower((unsigned char)*c);
    int node = -1;
    {
        std::ifstream f(std::string("/sys/bus/pci/devices/") + bus + "/numa_node");
        if (!(f >> node) || node < 0) return -1;
    }
    std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist");
    std::string list;
    if (!(f >> list) || list.empty()) return -1;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n_cpus = 0;
    for (size_t i = 0; i < list.size();) {
        char* end = nullptr;
        long a = std::strtol(list.c_str() + i, &end, 10), b = a;
        if (end == list.c_str() + i) break;
        if (*end == '-') b = std::strtol(end + 1, &end, 10);
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) CPU_SET((int)c, &set), ++n_cpus;
        i = (size_t)(end - list.c_str()) + (*end == ',' ? 1 : 0);
        if (*end != ',') break;
    }
    if (n_cpus == 0 || sched_setaffinity(0, sizeof set, &set) != 0) return -1;
    return node;
}
extern "C" int c6_host_cores_per_rank(void) { return c6::host_cores_per_rank(); }
extern "C" int c6_host_spin_budget_us(void) { return c6::host_spin_budget_us(); }
