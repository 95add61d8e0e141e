// Kernels + host driver of the batched self-play engine.  See c6_engine.cuh / DESIGN.md.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "c6_search.cuh"

namespace c6 {

static thread_local std::string g_last_error;
const char* last_error() { return g_last_error.c_str(); }

#define C6_CUDA(call)                                                                                 \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            char _buf[512];                                                                           \
            std::snprintf(_buf, sizeof _buf, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                          __LINE__, cudaGetErrorString(_e));                                          \
            g_last_error = _buf;                                                                      \
            throw std::runtime_error(_buf);                                                           \
        }                                                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_zobrist(const Dev& D, uint64_t* Zs) {
    for (int i = threadIdx.x; i < 512; i += blockDim.x) Zs[i] = D.Z[i];
    __syncthreads();
}

__global__ void k_start(Dev D, int n_start) {
    __shared__ uint64_t Zs[512];
    load_zobrist(D, Zs);
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= D.G) return;
    Warp w = make_warp(D, g, Zs, lane);
    if (g < n_start)
        init_game(D, w, D.first_id + (uint32_t)g);
    else if (lane == 0)
        w.gs->status = ST_DONE;
}

// Device-evaluator path, kernel 1: num_sim searches for every resident game.  Persistent warps pull
// game slots from a counter, so unequal game lengths balance themselves.  Register-capped for
// occupancy: the path is latency-bound, resident warps are what hides HBM latency.
#ifndef C6_SIM_BLOCKS_PER_SM
#define C6_SIM_BLOCKS_PER_SM 7
#endif
#ifndef C6_FUSE_MOVE_END
#define C6_FUSE_MOVE_END 1  // tune_j: 41.9 -> 41.0 ms per step (the separate kernel took 2.1 ms)
#endif
__device__ __noinline__ void finish_move_outofline(const Dev& D, Warp& w) { finish_move(D, w); }
template <int EVAL, int NOISE>
__global__ void __launch_bounds__(128, C6_SIM_BLOCKS_PER_SM) k_sims_device(Dev D) {
    // Zobrist keys are read straight from global memory (__ldg: 4 KB, L1-resident) so that eight blocks
    // of per-warp stages fit the shared memory of an SM.
    const uint64_t* Zs = D.Z;
    __shared__ WarpStage<NOISE> stage[4];
    __shared__ MathTab mt;
    if (NOISE || C6_SQRT_TAB) load_math_tab(mt);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    WarpStage<NOISE>& st = stage[threadIdx.x >> 5];
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(D.cursor, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= D.G) break;
        Warp w = make_warp(D, g, Zs, lane);
        if (w.gs->status != ST_ACTIVE) continue;
        Cursor cs;
        load_cursor(*w.gs, cs, st, lane);
        if (lane == 0) st.rootrec = bits_of(load_rec(&w.rec[w.gs->root]));
        __syncwarp();
        for (int s = 0; s < D.num_sim; ++s) {
            if (cs.err) {  // the pool is full: once per stone the search starts again from a fresh root (fresh_root)
                if (w.gs->retried) break;
                cs.err = 0;
                store_cursor(D, *w.gs, cs, st, lane);
                fresh_root(D, w);
                if (lane == 0) w.gs->retried = 1;
                load_cursor(*w.gs, cs, st, lane);
                if (lane == 0) st.rootrec = bits_of(load_rec(&w.rec[w.gs->root]));
                __syncwarp();
                s = 0;
            }
            Leaf L;
            descend<NOISE, true>(D, w, cs, L, st, mt);
            if (L.stones == CELLS) {
                terminal_leaf<true>(D, w, st, L);
            } else if (EVAL == EVAL_PHILOX) {
                EvPhilox ev;
                ev.init(D, cs, lane);
                expand_leaf<true>(D, w, cs, L, ev, ev.value(), st);
            } else if (EVAL == EVAL_DET) {
                EvHash<false> ev;
                ev.init(L, st.B, st.W, Zs, lane);
                expand_leaf<true>(D, w, cs, L, ev, ev.value(), st);
            } else {
                EvHash<true> ev;
                ev.init(L, st.B, st.W, Zs, lane);
                expand_leaf<true>(D, w, cs, L, ev, ev.value(), st);
            }
            cs.sim_index += 1;
            if (cs.err && s + 1 == D.num_sim && !w.gs->retried) s -= 1;  // the last simulation overflowed: take the restart branch
        }
        store_cursor(D, *w.gs, cs, st, lane);
#if C6_FUSE_MOVE_END
        // the stone itself, in the same launch: this warp's collection overlaps the other warps' searches instead of
        // running as a latency-bound kernel of its own
        if (w.gs->status == ST_ACTIVE) finish_move_outofline(D, w);
#endif
    }
}

// Device-evaluator path, kernel 2 (also the tail of the callback path): move choice, stone, win
// check, record, node collection / slot refill for every resident game.
__global__ void __launch_bounds__(128) k_move_end(Dev D, int g0, int g1) {
    __shared__ uint64_t Zs[512];
    load_zobrist(D, Zs);
    const int lane = threadIdx.x & 31;
    int g = g0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= g1) return;
    Warp w = make_warp(D, g, Zs, lane);
    if (w.gs->status == ST_ACTIVE) finish_move(D, w);
}

// A stone decided outside the engine for the game in `slot` (interactive play, external game drivers).
__global__ void __launch_bounds__(32) k_external_move(Dev D, int slot, int cell) {
    __shared__ uint64_t Zs[512];
    load_zobrist(D, Zs);
    Warp w = make_warp(D, slot, Zs, threadIdx.x & 31);
    if (w.gs->status == ST_ACTIVE) external_move(D, w, bit_of_cell(cell));
}

// Device-evaluator path in multi-leaf mode (Config::virtual_loss = K > 1, opt-in): the same per-game work as k_sims_device on the
// wave schedule of c6_search.cuh (wave_descents / wave_expansions; oracle twin AlphaZero::search_wave).
template <int EVAL, int NOISE>
__global__ void __launch_bounds__(128, C6_SIM_BLOCKS_PER_SM) k_sims_device_vl(Dev D) {
    const uint64_t* Zs = D.Z;
    __shared__ WarpStage<NOISE> stage[4];
    __shared__ MathTab mt;
    if (NOISE || C6_SQRT_TAB) load_math_tab(mt);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    WarpStage<NOISE>& st = stage[threadIdx.x >> 5];
    for (;;) {
        int g = 0;
        if (lane == 0) g = atomicAdd(D.cursor, 1);
        g = __shfl_sync(FULL, g, 0);
        if (g >= D.G) break;
        Warp w = make_warp(D, g, Zs, lane);
        if (w.gs->status != ST_ACTIVE) continue;
        Cursor cs;
        load_cursor(*w.gs, cs, st, lane);
        int done = 0;
        while (done < D.num_sim) {
            if (cs.err) {  // pool full: restart the stone's search from a fresh root, once
                if (w.gs->retried) break;
                cs.err = 0;
                store_cursor(D, *w.gs, cs, st, lane);
                fresh_root(D, w);
                if (lane == 0) w.gs->retried = 1;
                load_cursor(*w.gs, cs, st, lane);
                done = 0;
            }
            wave_descents<NOISE>(D, w, cs, st, mt, done);
            if (EVAL == EVAL_PHILOX) {
                wave_expansions<NOISE>(D, w, cs, st, [&](const PendingLeaf& p, const Leaf&) {
                    EvPhilox ev;
                    ev.init(D, cs.game_id, p.sim_index, lane);
                    return ev;
                });
            } else if (EVAL == EVAL_DET) {
                wave_expansions<NOISE>(D, w, cs, st, [&](const PendingLeaf&, const Leaf& L) {
                    EvHash<false> ev;
                    ev.init(L, st.B, st.W, Zs, lane);
                    return ev;
                });
            } else {
                wave_expansions<NOISE>(D, w, cs, st, [&](const PendingLeaf&, const Leaf& L) {
                    EvHash<true> ev;
                    ev.init(L, st.B, st.W, Zs, lane);
                    return ev;
                });
            }
        }
        store_cursor(D, *w.gs, cs, st, lane);
        if (w.gs->status == ST_ACTIVE) finish_move_outofline(D, w);
    }
}

// Callback path, kernel 1 — one wave of one pipeline group (slots g0..g1), one warp per game:
//   back half  (do_back):  recover_way8 + get_from on the evaluator's answers, expansion of the waiting leaves, their values;
//                          once the game's num_sim simulations of this stone are done also policy(), Game::set, win check,
//                          node collection / slot refill (what k_move_end does on the device-evaluator path);
//   front half (do_front): the descents of the next wave (one on the reference's schedule, up to K in multi-leaf mode); the
//                          leaves are parked in the game state; leaf_flag[g * K + i] = 0 (nothing to evaluate), 1 (Black to
//                          move at leaf i) or 2 (White).
// Both halves of consecutive waves touch only this game's own state, so they share a launch: a wave of a group
// is k_step -> k_gather -> (evaluator) -> one H2D copy.
template <int NOISE>
__global__ void __launch_bounds__(128) k_step(Dev D, int g0, int g1, int do_back, int do_front) {
    const uint64_t* Zs = D.Z;  // global (__ldg in descend / expand_leaf)
    __shared__ WarpStage<NOISE> stage[4];
    __shared__ MathTab mt;
    if (NOISE || C6_SQRT_TAB) load_math_tab(mt);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    WarpStage<NOISE>& st = stage[threadIdx.x >> 5];
    const int g = g0 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (g >= g1) return;
    Warp w = make_warp(D, g, Zs, lane);
    GameState& gs = *w.gs;
    if (do_back && gs.status == ST_ACTIVE) {
        if (D.num_sim > 0 && gs.n_pend > 0) {
            Cursor cs;
            load_cursor(gs, cs, st, lane);
            const size_t nk = (size_t)(g1 - g0) * D.K;
            const float* region = D.batch + (size_t)g0 * D.K * 8 * (CELLS + 1);
            wave_expansions<NOISE>(D, w, cs, st, [&](const PendingLeaf& p, const Leaf&) {
                EvBatch ev;
                ev.values = region + (size_t)p.batch_pos * 8;
                ev.pol = region + nk * 8 + (size_t)p.batch_pos * 8 * CELLS;
                ev.lane = lane;
                return ev;
            });
            if (cs.err && !gs.retried) {  // pool full: restart the stone's search from a fresh root, once
                cs.err = 0;
                store_cursor(D, gs, cs, st, lane);
                fresh_root(D, w);
                if (lane == 0) gs.retried = 1, gs.sims_this_move = 0;
                __syncwarp();
            } else {
                store_cursor(D, gs, cs, st, lane);
            }
        }
        if (gs.status == ST_ACTIVE && gs.sims_this_move >= D.num_sim) finish_move(D, w);
        __syncwarp();
    }
    if (!do_front) return;
    int n_pend = 0;
    if (gs.status == ST_ACTIVE && D.num_sim > 0) {
        Cursor cs;
        load_cursor(gs, cs, st, lane);
        int stm = gs.sims_this_move;
        wave_descents<NOISE>(D, w, cs, st, mt, stm);
        if (lane == 0) gs.sims_this_move = stm;
        store_cursor(D, gs, cs, st, lane);
        n_pend = gs.n_pend;
    } else if (lane == 0) {
        gs.n_pend = 0;
    }
    if (lane < D.K) D.leaf_flag[(size_t)g * D.K + lane] = (uint8_t)(lane < n_pend ? (turn_of(gs.pend[lane].stones) < 0 ? 1 : 2) : 0);
}

// Callback path, kernel 2: the waiting leaves of the group (slot q = (game - g0) * K + i for leaf i of a game) ordered into a
// Black-to-move run followed by a White-to-move run (the Callback ABI takes one `player` per call, CB/libcppconnect6/src/
// cppbind/cppeval/mod.rs:8-13; every warp ranks its own leaf from the group's flag bytes), then augment_way8 + board_to_float
// for every leaf in the API layout float[len][15][15] (LC/policy/alphazero_policy/augment.rs:60-73, CB/.../cpp_support.rs:10-18),
// one warp per leaf slot.  The batch region of a group of n games is [values n*K*8 | boards n*K*8*225] floats.  A lane builds
// four consecutive floats of the batch and stores them as one float4 (7,200 B per leaf = 450 float4: coalesced 512 B per warp).
// to_host: the boards go straight to the mapped pinned host buffer the user Callback reads (no copy call);
// the last block to finish publishes {Black leaves, White leaves, seq} to the host, which polls seq.
__global__ void __launch_bounds__(128) k_gather(Dev D, int g0, int g1, int group, int seq, int to_host, int publish) {
    __shared__ uint32_t boards[4][16];
    const int lane = threadIdx.x & 31;
    const int K = D.K;
    const int nk = (g1 - g0) * K;
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint8_t* flags = D.leaf_flag + (size_t)g0 * K;
    const int myflag = q < nk ? (int)flags[q] : 0;
    int nB = 0, nW = 0, rank = 0;
    if (myflag || threadIdx.x < 32) {  // warp-uniform
        for (int base = 0; base < nk; base += 32) {
            const int i = base + lane;
            const int f = i < nk ? (int)flags[i] : 0;
            const unsigned mb = __ballot_sync(FULL, f == 1), mw = __ballot_sync(FULL, f == 2);
            nB += __popc(mb), nW += __popc(mw);
            const unsigned mine = myflag == 1 ? mb : mw;
            if (base + 32 <= q) rank += __popc(mine);
            else if (base <= q) rank += __popc(mine & ((1u << (q - base)) - 1u));
        }
    }
    if (myflag) {
        PendingLeaf& p = D.gs[g0 + q / K].pend[q % K];
        const int slot = myflag == 1 ? rank : nB + rank;
        if (lane == 0) p.batch_pos = slot;
        uint32_t* sb = boards[threadIdx.x >> 5];  // words 0-7 black, 8-15 white
        if (lane < 16) sb[lane] = lane < 8 ? p.B[lane] : p.W[lane - 8];
        __syncwarp();
        float* region = (to_host ? D.hbatch : D.batch) + (size_t)g0 * K * 8 * (CELLS + 1);
        float4* out = reinterpret_cast<float4*>(region + (size_t)nk * 8 + (size_t)slot * 8 * CELLS);  // 7,200 B per leaf: 16-byte aligned
        for (int v = lane; v < 8 * CELLS / 4; v += 32) {
            float f[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int i = 4 * v + t;
                int s = i / CELLS, o = i - s * CELLS, rr = o / 15, cc = o - rr * 15, r, c;
                sym_inv(s, rr, cc, r, c);
                const uint32_t bit = (uint32_t)(r * 16 + c);
                const uint32_t bw = sb[bit >> 5], ww = sb[8 + (bit >> 5)];
                f[t] = ((bw >> (bit & 31)) & 1u) ? -1.f : (((ww >> (bit & 31)) & 1u) ? 1.f : 0.f);
            }
            out[v] = make_float4(f[0], f[1], f[2], f[3]);
        }
        if (!to_host && lane < 8) region[(size_t)slot * 8 + lane] = 0.f;  // values pre-zeroed (cppeval/mod.rs:34); the host path zeroes its own
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int done = atomicAdd(D.gdone + group, 1);
        if (done == (int)gridDim.x - 1) {
            D.gdone[group] = 0;
            volatile int* hs = D.hsync + 4 * group;
            hs[0] = nB, hs[1] = nW;
            __threadfence_system();
            if (publish) hs[2] = seq;
        }
    }
}

// D2H-by-copy variant (C6_ZEROCOPY=0): the sequence number follows the copy in stream order
__global__ void k_publish(Dev D, int group, int seq) { *(volatile int*)(D.hsync + 4 * group + 2) = seq; }

// Invariant checker (tests / debugging): for every resident game, every entry of the expanded list must
// (1) carry the Zobrist key of its board, (2) own an expanded record (N >= 1, real link run) that (3) the key
// table maps its key to, whose (4) children are unforwarded records inside the live part of the ring, and
// (5) a child board that is itself expanded is linked to THAT record (one record per board), (6) no record is
// listed twice.  Violations are counted in viol[0]; the first few are described in viol[1..].
__global__ void __launch_bounds__(128) k_validate(Dev D, unsigned long long* viol) {
    __shared__ uint64_t Zs[512];
    load_zobrist(D, Zs);
    const int lane = threadIdx.x & 31;
    int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= D.G) return;
    Warp w = make_warp(D, g, Zs, lane);
    GameState& gs = *w.gs;
    if (gs.status != ST_ACTIVE) return;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1, Im = D.I - 1;
    auto report = [&](int code, uint32_t a, uint32_t b, uint32_t c) {
        unsigned long long n = atomicAdd(viol, 1ull);
        if (n < 16) {
            viol[1 + n * 4 + 0] = ((unsigned long long)g << 32) | (unsigned)code;
            viol[1 + n * 4 + 1] = a, viol[1 + n * 4 + 2] = b, viol[1 + n * 4 + 3] = c;
        }
    };
    if (lane == 0) {
        uint32_t used = gs.rec_head - gs.rec_tail;
        if (((gs.root - gs.rec_tail) & Rm) >= used) report(9, gs.root, gs.rec_tail, gs.rec_head);
        const Rec rr = load_rec(&w.rec[gs.root & Rm]);
        if (rr.n >= 1 && xh_find(w.idx, Im, gs.epoch, gs.key) != gs.root) report(1, gs.root, 0, 0);
    }
    const uint32_t n_exp = gs.exp_count;
    for (uint32_t ei = 0; ei < n_exp; ++ei) {
        uint32_t eb[8], ew[8];
        int est = 0;
        uint64_t key = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            eb[j] = w.expB(D)[(size_t)ei * 16 + j], ew[j] = w.expB(D)[(size_t)ei * 16 + 8 + j];
            est += __popc(eb[j]) + __popc(ew[j]);
            if ((eb[j] >> lane) & 1u) key ^= Zs[32 * j + lane];
            if ((ew[j] >> lane) & 1u) key ^= Zs[256 + 32 * j + lane];
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) key ^= __shfl_xor_sync(FULL, key, off);
        const uint32_t erec = w.exp_rec(D)[ei];
        const uint64_t ekey = w.exp_key(D)[ei];
        Rec r = load_rec(&w.rec[erec & Rm]);
        if (lane == 0) {
            if (key != ekey) {
                report(2, ei, erec, (uint32_t)est);
                report(20, (uint32_t)(key ^ ekey), (uint32_t)((key ^ ekey) >> 32), 0);
            }
            if (r.n < 1 || r.link == NIL || r.link == FWD) report(3, ei, erec, (uint32_t)r.n);
            const uint32_t val = xh_find(w.idx, Im, gs.epoch, ekey);
            if (val != erec) report(4, ei, erec, val);
            if (((erec - gs.rec_tail) & Rm) >= gs.rec_head - gs.rec_tail) report(8, ei, erec, gs.rec_tail);
        }
        for (uint32_t e2 = ei + 1 + lane; e2 < n_exp; e2 += 32)
            if (w.exp_rec(D)[e2] == erec) report(5, ei, e2, erec);
        if (r.n < 1 || r.link == NIL || r.link == FWD) continue;
        const int turn = turn_of(est);
        int pre = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t e = ~(eb[j] | ew[j]) & valid_mask(j);
            if ((e >> lane) & 1u) {
                int k = pre + __popc(e & ((1u << lane) - 1u));
                uint32_t ci = w.lnk[(r.link + (uint32_t)k) & Lm];
                Rec ch = load_rec(&w.rec[ci & Rm]);
                if (ch.link == FWD) report(6, ei, ci, (uint32_t)k);
                const uint32_t val = xh_find(w.idx, Im, gs.epoch, ekey ^ Zs[(turn > 0 ? 256 : 0) + 32 * j + lane]);
                if (val != NIL ? val != ci : ch.n >= 1) report(7, ei, ci, val);
                if (((ci - gs.rec_tail) & Rm) >= gs.rec_head - gs.rec_tail) report(10, ei, ci, gs.rec_tail);
            }
            pre += __popc(e);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// One pipeline group of the callback path: a contiguous range of game slots with its own stream, its own
// region of the batch buffers (device and pinned host) and its own position in the simulation budget.
struct Group {
    int g0 = 0, g1 = 0;
    cudaStream_t s = nullptr;
    cudaEvent_t ev = nullptr;
    int seq = 0;  // sequence number of the last leaf batch asked for (k_gather publishes it to hsync when the batch is ready)
};

// Blocks an engine frees are kept for the next engine of the process (same device, same kind, same size): a run of short
// jobs — the reference's 2-simulation sample game, a test suite — otherwise spends its time in cudaMalloc / cudaHostAlloc /
// cudaFree (6 ms each way for a one-game engine against 5 ms of search, profiles/r02o_config1.jsonl).  Bounded: at most
// C6_ALLOC_CACHE_MB (default 256) in total and 16 blocks, so a 30 GB node pool is never held back.  Contents are whatever the
// last owner left: the constructor clears what it reads before writing, as it must for fresh memory too.
void release_cached_memory();
namespace {
struct CachedBlock {
    int device;
    bool pinned;
    size_t bytes;
    void* p;
};
std::mutex g_cache_mu;
std::vector<CachedBlock> g_cache;
size_t g_cache_bytes = 0;
size_t cache_limit_bytes() {
    static const size_t lim = [] {
        const char* v = std::getenv("C6_ALLOC_CACHE_MB");
        return (size_t)(v && *v ? std::max(0, std::atoi(v)) : 256) << 20;
    }();
    return lim;
}
void* cached_alloc(int device, size_t bytes, bool pinned) {
    {
        std::lock_guard<std::mutex> l(g_cache_mu);
        for (size_t i = 0; i < g_cache.size(); ++i)
            if (g_cache[i].device == device && g_cache[i].pinned == pinned && g_cache[i].bytes == bytes) {
                void* p = g_cache[i].p;
                g_cache_bytes -= bytes;
                g_cache.erase(g_cache.begin() + (long)i);
                return p;
            }
    }
    void* p = nullptr;
    auto try_alloc = [&] { return pinned ? cudaHostAlloc(&p, bytes, cudaHostAllocMapped) : cudaMalloc(&p, bytes); };
    if (try_alloc() != cudaSuccess) {  // out of memory with blocks held back: give them up and try once more
        cudaGetLastError();
        release_cached_memory();
        C6_CUDA(try_alloc());
    }
    return p;
}
// the caller has synchronised every stream that touched the block
void cached_free(int device, void* p, size_t bytes, bool pinned) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> l(g_cache_mu);
        if (bytes <= cache_limit_bytes() / 2 && g_cache_bytes + bytes <= cache_limit_bytes() && g_cache.size() < 16) {
            g_cache.push_back(CachedBlock{device, pinned, bytes, p});
            g_cache_bytes += bytes;
            return;
        }
    }
    if (pinned) cudaFreeHost(p); else cudaFree(p);
}
}  // namespace
void release_cached_memory() {
    std::vector<CachedBlock> blocks;
    {
        std::lock_guard<std::mutex> l(g_cache_mu);
        blocks.swap(g_cache);
        g_cache_bytes = 0;
    }
    for (const CachedBlock& b : blocks) {
        cudaSetDevice(b.device);
        if (b.pinned) cudaFreeHost(b.p); else cudaFree(b.p);
    }
}

struct DeviceArrays {
    Dev dev{};
    size_t arena_bytes = 0, pinned_bytes = 0, batch_bytes = 0, records_bytes = 0, trace_rows = 0;
    char* arena = nullptr;     // the one device allocation every array of `dev` is carved from (records / traces apart)
    char* h_pinned = nullptr;  // the one small pinned allocation: h_sync, h_gcount
    float* h_batch = nullptr;  // mapped pinned host memory, same layout as Dev::batch
    int* h_sync = nullptr;     // mapped pinned host memory, per group {Black leaves, White leaves, seq, -}
    std::vector<Group> groups;
    unsigned long long* h_gcount = nullptr;
    size_t rec_bytes = 0, lnk_bytes = 0, idx_bytes = 0;
    int sm_count = 0;
    int persistent_blocks = 0;
};

static void make_zobrist(uint64_t* Zb) {  // same sequence as the test oracle's table, by bit position
    std::memset(Zb, 0, 512 * sizeof(uint64_t));
    uint64_t s = 0xC0221EC76B200ull;
    for (int p = 0; p < 2; ++p)
        for (int i = 0; i < CELLS; ++i) {
            s = splitmix64(s);
            Zb[p * 256 + (i / 15) * 16 + (i % 15)] = s;
        }
}

void make_zobrist_table(uint64_t* Zb) { make_zobrist(Zb); }  // c6_default_policy.cu

size_t Engine::bytes_per_game() const {
    size_t R = (size_t)1 << cfg_.log2_records;
    return R * sizeof(Rec) + 2 * R * sizeof(uint32_t) + d_->dev.I * sizeof(Slot) + d_->dev.E * (64 + 4 + 8) + 232 * 4 +
           sizeof(GameState);
}

Engine::Engine(const Config& cfg) : cfg_(cfg) {
    if (cfg_.slots < 1) throw std::runtime_error("c6: slots must be >= 1");
    if (cfg_.log2_records < 10 || cfg_.log2_records > 24) throw std::runtime_error("c6: log2_records out of range [10,24]");
    if (cfg_.virtual_loss < 1) cfg_.virtual_loss = 1;
    if (cfg_.virtual_loss > VL_MAX) throw std::runtime_error("c6: virtual_loss out of range [1," + std::to_string(VL_MAX) + "]");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_last_error = "c6: no CUDA device available (this library has no CPU fallback)";
        throw std::runtime_error(g_last_error);
    }
    C6_CUDA(cudaSetDevice(cfg_.device));
    if (const char* fg = std::getenv("C6_L2_FETCH")) {  // tuning knob: 32 / 64 / 128 B L2 fetch granularity
        cudaError_t fe = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)std::atoi(fg));
        if (fe != cudaSuccess) cudaGetLastError();
    }
    C6_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    d_ = new DeviceArrays();
    Dev& D = d_->dev;
    D.G = cfg_.slots;
    D.R = 1u << cfg_.log2_records;
    D.L = D.R * 2;
    // Expanded nodes alive in one game: at most num_simulation new ones per stone on top of the subtree kept from the stone
    // before (itself at most the visits of the chosen child).  E = 4,096 holds the reference default (800); larger budgets scale.
    D.E = std::max<uint32_t>(4096u, (uint32_t)(((long long)std::max(cfg_.num_simulation, 0) * 5 / 2 + 1023) / 1024 * 1024));
    D.I = 8192;  // key table of expanded nodes: power of two >= 2 * E
    while (D.I < 2 * D.E) D.I *= 2;
    // A stone's search creates at most 225 records per simulation: a FRESH tree must fit the ring, or the restart of an
    // overflowing search (fresh_root) could not finish either.
    if ((long long)std::max(cfg_.num_simulation, 0) * CELLS + 2 * CELLS + 2 > (long long)D.R) {
        char buf[256];
        std::snprintf(buf, sizeof buf, "c6: %d simulations per stone need more than %u node records per game; raise log2_records (C6_LOG2_RECORDS) to at least %d",
                      cfg_.num_simulation, D.R, (int)std::ceil(std::log2((double)cfg_.num_simulation * CELLS + 2 * CELLS + 2)));
        g_last_error = buf;
        throw std::runtime_error(buf);
    }
    D.num_sim = cfg_.num_simulation, D.eps = cfg_.epsilon, D.gp = gamma_params(cfg_.dirichlet_alpha), D.cpuct = cfg_.c_puct;
    D.seed = cfg_.seed, D.trace = cfg_.trace, D.max_moves = cfg_.max_moves, D.tie_break = cfg_.tie_break;
    D.K = cfg_.virtual_loss;
    const size_t G = (size_t)D.G;
    d_->rec_bytes = G * D.R * sizeof(Rec);
    d_->lnk_bytes = G * D.L * sizeof(uint32_t);
    d_->idx_bytes = G * D.I * sizeof(Slot);
    size_t free_b = 0, total_b = 0;
    C6_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t need = d_->rec_bytes + d_->lnk_bytes + d_->idx_bytes + G * D.E * 76 + G * D.K * 8 * (CELLS + 1) * 4 + (64u << 20);
    if (need > free_b) {
        char buf[256];
        std::snprintf(buf, sizeof buf,
                      "c6: node pool for %d games at 2^%d records needs %.1f GB, device has %.1f GB free; lower "
                      "C6_LOG2_RECORDS or the number of concurrent games",
                      D.G, cfg_.log2_records, need / 1e9, free_b / 1e9);
        g_last_error = buf;
        throw std::runtime_error(buf);
    }
    int n_groups = 1;
    const bool callback = cfg_.evaluator == EVAL_CALLBACK || cfg_.evaluator == EVAL_DEVICE_CALLBACK;
    if (callback) {
        // pipeline depth of the callback path: while the host evaluates one group, the others run their
        // kernels and copies.  Small batches are not worth splitting.
        const char* pe = std::getenv("C6_PIPELINE");
        n_groups = pe && *pe ? std::atoi(pe) : std::max(1, std::min(16, D.G / 1024));  // 4 groups of 1024 games at G = 4096: profiles/e2e/e2e7.md
        n_groups = std::max(1, std::min(n_groups, std::min(D.G, 16)));
    }
    // ONE device allocation, carved into the arrays (256-byte aligned), and one small pinned block: a short job (the
    // reference's 2-simulation sample game) is otherwise dominated by ~20 cudaMalloc / cudaFree round trips to the driver.
    size_t top = 0;
    auto carve = [&top](size_t bytes) { const size_t o = top; top = (top + bytes + 255) & ~(size_t)255; return o; };
    const size_t batch_bytes = callback ? G * D.K * 8 * (CELLS + 1) * sizeof(float) : 0;
    const size_t o_rec = carve(d_->rec_bytes), o_lnk = carve(d_->lnk_bytes), o_idx = carve(d_->idx_bytes);
    const size_t o_expB = carve(G * D.E * 64), o_exp_rec = carve(G * D.E * 4), o_exp_key = carve(G * D.E * 8);
    const size_t o_path = carve(G * D.K * 232 * 4);
    // zero-initialised tail: game states, group counters, leaf flags, the cursors
    const size_t o_zero = top;
    const size_t o_gs = carve(G * sizeof(GameState)), o_gdone = carve(n_groups * sizeof(int)), o_flag = carve(G * D.K);
    const size_t o_cursor = carve(sizeof(int)), o_gcount = carve(4 * sizeof(unsigned long long));
    const size_t zero_bytes = top - o_zero;
    const size_t o_Z = carve(512 * sizeof(uint64_t)), o_batch = carve(batch_bytes);
    d_->arena_bytes = top;
    d_->arena = (char*)cached_alloc(cfg_.device, top, false);
    char* base = d_->arena;
    D.rec = (Rec*)(base + o_rec), D.lnk = (uint32_t*)(base + o_lnk), D.idx = (Slot*)(base + o_idx);
    D.expB = (uint32_t*)(base + o_expB), D.exp_rec = (uint32_t*)(base + o_exp_rec), D.exp_key = (uint64_t*)(base + o_exp_key);
    D.path = (uint32_t*)(base + o_path), D.gs = (GameState*)(base + o_gs), D.gdone = (int*)(base + o_gdone);
    D.leaf_flag = (uint8_t*)(base + o_flag), D.cursor = (int*)(base + o_cursor), D.gcount = (unsigned long long*)(base + o_gcount);
    if (callback) D.batch = (float*)(base + o_batch);
    C6_CUDA(cudaMemsetAsync(D.idx, 0, d_->idx_bytes, stream_));  // epoch 0 = never valid
    C6_CUDA(cudaMemsetAsync(base + o_zero, 0, zero_bytes, stream_));
    uint64_t hz[512];
    make_zobrist(hz);
    C6_CUDA(cudaMemcpyAsync(base + o_Z, hz, sizeof hz, cudaMemcpyHostToDevice, stream_));
    D.Z = (const uint64_t*)(base + o_Z);
    // pinned: per group {Black leaves, White leaves, seq, -} (mapped: k_gather publishes into it), then the 4 game counters
    const size_t sync_bytes = ((size_t)4 * n_groups * sizeof(int) + 63) & ~(size_t)63;
    d_->pinned_bytes = sync_bytes + 4 * sizeof(unsigned long long);
    d_->h_pinned = (char*)cached_alloc(cfg_.device, d_->pinned_bytes, true);
    std::memset(d_->h_pinned, 0, d_->pinned_bytes);
    d_->h_sync = (int*)d_->h_pinned, d_->h_gcount = (unsigned long long*)(d_->h_pinned + sync_bytes);
    C6_CUDA(cudaHostGetDevicePointer((void**)&D.hsync, d_->h_sync, 0));
    if (const char* ul = std::getenv("C6_LAUNCHER")) use_launcher_ = std::atoi(ul) != 0;
    if (const char* uz = std::getenv("C6_ZEROCOPY")) zero_copy_ = std::atoi(uz) != 0;
    d_->groups.resize(n_groups);
    for (int p = 0; p < n_groups; ++p) {
        Group& gr = d_->groups[p];
        gr.g0 = (int)((long long)D.G * p / n_groups), gr.g1 = (int)((long long)D.G * (p + 1) / n_groups);
        if (n_groups == 1) {
            gr.s = stream_;
        } else {
            C6_CUDA(cudaStreamCreateWithFlags(&gr.s, cudaStreamNonBlocking));
        }
        // blocking-sync event: a thread that waits on it sleeps in the driver instead of spinning (callback path:
        // the calling thread waits for a group's leaf batch; eight ranks on one host must not burn eight cores on it)
        C6_CUDA(cudaEventCreateWithFlags(&gr.ev, cudaEventDisableTiming | cudaEventBlockingSync));
    }
    spin_us_ = host_spin_budget_us();
    if (cfg_.evaluator == EVAL_CALLBACK) {
        d_->batch_bytes = batch_bytes;
        d_->h_batch = (float*)cached_alloc(cfg_.device, batch_bytes, true);
        C6_CUDA(cudaHostGetDevicePointer((void**)&D.hbatch, d_->h_batch, 0));
    }
    C6_CUDA(cudaDeviceGetAttribute(&d_->sm_count, cudaDevAttrMultiProcessorCount, cfg_.device));
    C6_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::free_records() {
    Dev& D = d_->dev;
    cached_free(cfg_.device, D.records, d_->records_bytes, false);
    cached_free(cfg_.device, D.tr_visits, d_->trace_rows * CELLS * 4, false);
    cached_free(cfg_.device, D.tr_n, d_->trace_rows * 4, false);
    cached_free(cfg_.device, D.tr_w, d_->trace_rows * 4, false);
    D.records = nullptr, D.tr_visits = nullptr, D.tr_n = nullptr, D.tr_w = nullptr;
    d_->records_bytes = 0, d_->trace_rows = 0;
}

Engine::~Engine() {
    if (!d_) return;
    cudaSetDevice(cfg_.device);
    cudaStreamSynchronize(stream_);
    for (Group& gr : d_->groups) {
        if (gr.s && gr.s != stream_) cudaStreamSynchronize(gr.s), cudaStreamDestroy(gr.s);
        if (gr.ev) cudaEventDestroy(gr.ev);
    }
    // nothing of this engine is in flight any more: the blocks may go to the next one
    free_records();
    cached_free(cfg_.device, d_->arena, d_->arena_bytes, false);
    cached_free(cfg_.device, d_->h_batch, d_->batch_bytes, true);
    cached_free(cfg_.device, d_->h_pinned, d_->pinned_bytes, true);
    cudaStreamDestroy(stream_);
    delete d_;
}

static inline int warps_grid(int G) { return (G * 32 + 127) / 128; }

void Engine::start(uint32_t first_id, int64_t total_games) {
    Dev& D = d_->dev;
    C6_CUDA(cudaSetDevice(cfg_.device));
    sync();
    total_games_ = total_games, first_id_ = first_id;
    D.total_games = total_games, D.first_id = first_id;
    free_records();
    d_->records_bytes = (size_t)std::max<int64_t>(total_games, 1) * sizeof(GameRecord);
    D.records = (GameRecord*)cached_alloc(cfg_.device, d_->records_bytes, false);
    C6_CUDA(cudaMemsetAsync(D.records, 0, d_->records_bytes, stream_));
    if (cfg_.trace) {
        size_t n = (size_t)total_games * CELLS;
        d_->trace_rows = n;
        D.tr_visits = (int32_t*)cached_alloc(cfg_.device, n * CELLS * 4, false);
        D.tr_n = (int32_t*)cached_alloc(cfg_.device, n * 4, false);
        D.tr_w = (float*)cached_alloc(cfg_.device, n * 4, false);
        C6_CUDA(cudaMemsetAsync(D.tr_visits, 0xFF, n * CELLS * 4, stream_));
        C6_CUDA(cudaMemsetAsync(D.tr_n, 0, n * 4, stream_));
        C6_CUDA(cudaMemsetAsync(D.tr_w, 0, n * 4, stream_));
    }
    int n_start = (int)std::min<int64_t>(D.G, total_games);
    unsigned long long gc[4] = {0, (unsigned long long)n_start, 0, 0};
    C6_CUDA(cudaMemcpyAsync(D.gcount, gc, sizeof gc, cudaMemcpyHostToDevice, stream_));
    k_start<<<warps_grid(D.G), 128, 0, stream_>>>(D, n_start);
    launches_ += 1;
    C6_CUDA(cudaGetLastError());
    C6_CUDA(cudaStreamSynchronize(stream_));
}

static int noise_mode(const Config& c) { return c.epsilon == 0.f ? 0 : (c.dirichlet_alpha < 1.0 ? 1 : 2); }

// the one (evaluator, noise) variant of the fused kernel this engine runs
using SimKernel = void (*)(Dev);
static SimKernel sim_kernel(int evaluator, int noise, bool multi_leaf) {
    static const SimKernel table[3][3] = {
        {k_sims_device<EVAL_PHILOX, 0>, k_sims_device<EVAL_PHILOX, 1>, k_sims_device<EVAL_PHILOX, 2>},
        {k_sims_device<EVAL_DET, 0>, k_sims_device<EVAL_DET, 1>, k_sims_device<EVAL_DET, 2>},
        {k_sims_device<EVAL_NEG, 0>, k_sims_device<EVAL_NEG, 1>, k_sims_device<EVAL_NEG, 2>}};
    static const SimKernel table_vl[3][3] = {
        {k_sims_device_vl<EVAL_PHILOX, 0>, k_sims_device_vl<EVAL_PHILOX, 1>, k_sims_device_vl<EVAL_PHILOX, 2>},
        {k_sims_device_vl<EVAL_DET, 0>, k_sims_device_vl<EVAL_DET, 1>, k_sims_device_vl<EVAL_DET, 2>},
        {k_sims_device_vl<EVAL_NEG, 0>, k_sims_device_vl<EVAL_NEG, 1>, k_sims_device_vl<EVAL_NEG, 2>}};
    return (multi_leaf ? table_vl : table)[evaluator - EVAL_PHILOX][noise];
}

void Engine::step_device() {
    Dev& D = d_->dev;
    if (cfg_.evaluator == EVAL_CALLBACK || cfg_.evaluator == EVAL_DEVICE_CALLBACK) throw std::runtime_error("c6: step_device needs a built-in device evaluator");
    const SimKernel kern = sim_kernel(cfg_.evaluator, noise_mode(cfg_), cfg_.virtual_loss > 1);
    if (d_->persistent_blocks == 0) {
        int per_sm = 0;
        C6_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0));
        if (per_sm < 1) per_sm = 1;
        d_->persistent_blocks = std::min(d_->sm_count * per_sm, warps_grid(D.G));
    }
    C6_CUDA(cudaMemsetAsync(D.cursor, 0, sizeof(int), stream_));
    if (cfg_.num_simulation > 0) {
        cudaEvent_t k0 = nullptr, k1 = nullptr;
        if (time_kernels_) {
            C6_CUDA(cudaEventCreate(&k0));
            C6_CUDA(cudaEventCreate(&k1));
            C6_CUDA(cudaEventRecord(k0, stream_));
        }
        kern<<<d_->persistent_blocks, 128, 0, stream_>>>(D);
        if (time_kernels_) {
            C6_CUDA(cudaEventRecord(k1, stream_));
            kev_.push_back(k0), kev_.push_back(k1);
        }
        launches_ += 1;
    }
#if C6_FUSE_MOVE_END
    if (cfg_.num_simulation <= 0)
#endif
    {
        k_move_end<<<warps_grid(D.G), 128, 0, stream_>>>(D, 0, D.G);
        launches_ += 1;
    }
    C6_CUDA(cudaGetLastError());
}

// ---- callback path ------------------------------------------------------------------------------
// One wave of one group: k_step (expansion + backup for the evaluator's answer to the previous leaf, move end
// when the stone's budget is spent, descent to the next leaf) -> k_gather (leaf boards in the API layout,
// written straight into the mapped pinned host buffer; the last block publishes the counts and a sequence
// number the host polls) -> user Callback per side to move -> one H2D copy of values and policies.  Three
// driver calls per group and wave, no events.  Groups run on their own streams: while the calling thread sits
// in the Callback of one group, the kernels and copies of the others proceed, and PCIe carries both directions.
static inline void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}
static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// Host cores this process may keep busy: hardware threads divided by the number of ranks sharing the box
// (LOCAL_WORLD_SIZE under torchrun; C6_LOCAL_RANKS overrides).
int host_cores_per_rank() {
    int hw = (int)std::thread::hardware_concurrency();
    if (hw < 1) hw = 1;
    int ranks = 1;
    for (const char* name : {"C6_LOCAL_RANKS", "LOCAL_WORLD_SIZE"})
        if (const char* v = std::getenv(name)) {
            if (*v && std::atoi(v) > 0) {
                ranks = std::atoi(v);
                break;
            }
        }
    return std::max(1, hw / ranks);
}
// How long a host thread may poll before it blocks (C6_SPIN_US overrides): with cores to spare polling is the
// fastest hand-over; with 4 cores per rank (8 ranks on a 32-CPU box) every polling thread takes a core from the
// evaluator, so it blocks almost at once.
int host_spin_budget_us() {
    if (const char* v = std::getenv("C6_SPIN_US"))
        if (*v) return std::max(0, std::atoi(v));
    return host_cores_per_rank() >= 8 ? 200 : 5;
}

// Enqueue for group p: [H2D of `len` evaluated boards] -> k_step(do_back, do_front) [-> k_gather].
void Engine::group_issue(int p, size_t len, bool do_back, bool do_front) {
    Dev& D = d_->dev;
    Group& gr = d_->groups[p];
    const int n = gr.g1 - gr.g0, grid = warps_grid(n);
    const size_t nk = (size_t)n * D.K;  // leaf slots of the group
    const size_t roff = (size_t)gr.g0 * D.K * 8 * (CELLS + 1);
    const bool host_cb = cfg_.evaluator == EVAL_CALLBACK;
    if (do_back && len && host_cb) {
        const size_t bytes = (nk * 8 + len * CELLS) * sizeof(float);  // all values of the region + the evaluated boards
        C6_CUDA(cudaMemcpyAsync(D.batch + roff, d_->h_batch + roff, bytes, cudaMemcpyHostToDevice, gr.s));
        h2d_bytes_ += (double)bytes;
    }
    const int noise = noise_mode(cfg_);
    if (noise == 1)
        k_step<1><<<grid, 128, 0, gr.s>>>(D, gr.g0, gr.g1, do_back, do_front);
    else if (noise == 2)
        k_step<2><<<grid, 128, 0, gr.s>>>(D, gr.g0, gr.g1, do_back, do_front);
    else
        k_step<0><<<grid, 128, 0, gr.s>>>(D, gr.g0, gr.g1, do_back, do_front);
    launches_ += 1;
    if (do_front && cfg_.num_simulation > 0) {
        gr.seq += 1;
        const bool zc = host_cb && zero_copy_;
        k_gather<<<warps_grid((int)nk), 128, 0, gr.s>>>(D, gr.g0, gr.g1, p, gr.seq, zc ? 1 : 0, (zc || !host_cb) ? 1 : 0);
        launches_ += 1;
        if (host_cb && !zc) {
            const size_t boff = roff + nk * 8;
            C6_CUDA(cudaMemcpyAsync(d_->h_batch + boff, D.batch + boff, nk * 8 * CELLS * sizeof(float), cudaMemcpyDeviceToHost, gr.s));
            k_publish<<<1, 1, 0, gr.s>>>(D, p, gr.seq);
            launches_ += 1;
        }
        C6_CUDA(cudaEventRecord(gr.ev, gr.s));
    }
    C6_CUDA(cudaGetLastError());
}

// The calling thread's half of a wave of one group: wait for the leaf batch, run the user Callback once per
// side to move (cppeval/mod.rs:8-13).  Returns the number of boards to copy back (8 per evaluated leaf).
size_t Engine::group_eval(int p, Callback cb) {
    Dev& D = d_->dev;
    Group& gr = d_->groups[p];
    if (cfg_.num_simulation <= 0) return 0;
    const int n = gr.g1 - gr.g0;
    const double t0 = now_s();
    volatile int* hs = d_->h_sync + 4 * p;
    const int want = gr.seq;
    // bounded poll of the sequence number the gather kernel publishes, then a blocking wait on the event recorded
    // behind it (the thread sleeps in the driver; a failed kernel surfaces as the event's error)
    const double spin_until = t0 + 1e-6 * spin_us_;
    while (hs[2] != want) {
        cpu_relax();
        if (now_s() >= spin_until) {
            if (abort_.load(std::memory_order_relaxed)) throw std::runtime_error("c6: wave aborted");
            C6_CUDA(cudaEventSynchronize(gr.ev));
            if (hs[2] != want) throw std::runtime_error("c6: leaf batch was never published");
            host_blocked_waits_ += 1;
        }
    }
    const double t1 = now_s();
    host_wait_s_ += t1 - t0;
    const int nB = hs[0], nW = hs[1];
    const size_t len = (size_t)(nB + nW) * 8, roff = (size_t)gr.g0 * D.K * 8 * (CELLS + 1);
    const size_t nk = (size_t)n * D.K;
    if (!len) return 0;
    if (cfg_.evaluator == EVAL_DEVICE_CALLBACK) {
        float* dv = D.batch + roff;  // zeroed by k_gather
        float* db = dv + nk * 8;
        if (nB) dcb_(-1, dv, db, nB * 8, (void*)gr.s, dcb_user_);
        if (nW) dcb_(1, dv + (size_t)nB * 8, db + (size_t)nB * 8 * CELLS, nW * 8, (void*)gr.s, dcb_user_);
        host_callback_s_ += now_s() - t1;
        return 0;  // nothing to copy back
    }
    float* hv = d_->h_batch + roff;
    float* hb = hv + nk * 8;
    std::memset(hv, 0, len * sizeof(float));  // values pre-zeroed (cppeval/mod.rs:34)
    const double t2 = now_s();
    if (nB) cb(-1, hv, hb, nB * 8);
    if (nW) cb(1, hv + (size_t)nB * 8, hb + (size_t)nB * 8 * CELLS, nW * 8);
    host_callback_s_ += now_s() - t2;
    d2h_bytes_ += (double)(zero_copy_ ? len : nk * 8) * CELLS * sizeof(float) + 12;
    return len;
}

// Per-group hand-over between the calling thread (evaluator callbacks) and the launching thread (every CUDA
// enqueue of the wave).  A group's stream is only ever touched by the launching thread; issued = w+1 hands
// wave w to the caller, cb_seq = w+1 hands it back together with the number of boards to copy.
struct alignas(64) GroupSync {
    std::atomic<int> issued{0}, cb_seq{0};
    size_t len = 0;
};

// `waves` simulations for every resident game with the CUDA enqueues on a second thread: side by side with the
// evaluator instead of in series with it (profiles/e2e5-7).  The user Callback still runs on the calling thread
// only, one call at a time — the reference's single-thread contract (lib.rs:98-108) holds.
void Engine::waves_callback_threaded(Callback cb, int waves) {
    const int P = (int)d_->groups.size();
    std::unique_ptr<GroupSync[]> gs(new GroupSync[P]);
    abort_.store(false);
    std::exception_ptr launcher_error;
    auto wait_for = [&](std::atomic<int>& a, int want) {
        const double until = now_s() + 1e-6 * spin_us_;
        while (a.load(std::memory_order_acquire) < want) {
            if (abort_.load(std::memory_order_relaxed)) return false;
            if (now_s() < until) cpu_relax(); else std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
        return true;
    };
    std::thread launcher([&] {
        try {
            C6_CUDA(cudaSetDevice(cfg_.device));
            for (int p = 0; p < P; ++p) {
                group_issue(p, 0, false, true);
                gs[p].issued.store(1, std::memory_order_release);
            }
            for (int wv = 0; wv < waves; ++wv)
                for (int p = 0; p < P; ++p) {
                    if (!wait_for(gs[p].cb_seq, wv + 1)) return;
                    group_issue(p, gs[p].len, true, wv + 1 < waves);
                    gs[p].issued.store(wv + 2, std::memory_order_release);
                }
        } catch (...) {
            launcher_error = std::current_exception();
            abort_.store(true);
        }
    });
    bool cancelled = false;
    std::exception_ptr caller_error;
    try {
        for (int wv = 0; wv < waves && !cancelled; ++wv)
            for (int p = 0; p < P; ++p) {
                const double t0 = now_s();
                if (!wait_for(gs[p].issued, wv + 1)) { cancelled = true; break; }
                host_wait_s_ += now_s() - t0;
                gs[p].len = group_eval(p, cb);
                gs[p].cb_seq.store(wv + 1, std::memory_order_release);
                if (cancel_.load()) { cancelled = true; abort_.store(true); break; }
            }
    } catch (...) {
        caller_error = std::current_exception();
        abort_.store(true);
    }
    launcher.join();
    if (launcher_error) std::rethrow_exception(launcher_error);
    if (caller_error) std::rethrow_exception(caller_error);
    if (cancelled) {
        sync();
        throw std::runtime_error("c6: cancelled");
    }
}

// `waves` simulations for every resident game, software-pipelined over the groups.
void Engine::waves_callback(Callback cb, int waves) {
    if (cfg_.evaluator != EVAL_CALLBACK) throw std::runtime_error("c6: engine was not created for the callback path");
    const int P = (int)d_->groups.size();
    if (waves <= 0) return;
    if (use_launcher_ && P > 1 && cfg_.num_simulation > 0) return waves_callback_threaded(cb, waves);
    abort_.store(false);
    for (int p = 0; p < P; ++p) group_issue(p, 0, false, true);
    for (int wv = 0; wv < waves; ++wv)
        for (int p = 0; p < P; ++p) {
            const size_t len = group_eval(p, cb);
            group_issue(p, len, true, wv + 1 < waves);
            if (cancel_.load()) {
                sync();
                throw std::runtime_error("c6: cancelled");
            }
        }
}

void Engine::wave_callback(Callback cb) { waves_callback(cb, 1); }

void Engine::waves_device_callback(DeviceCallback cb, void* user, int waves) {
    if (cfg_.evaluator != EVAL_DEVICE_CALLBACK) throw std::runtime_error("c6: engine was not created for the device-callback path");
    dcb_ = cb, dcb_user_ = user;
    const int P = (int)d_->groups.size();
    if (waves <= 0) return;
    abort_.store(false);
    for (int p = 0; p < P; ++p) group_issue(p, 0, false, true);
    for (int wv = 0; wv < waves; ++wv)
        for (int p = 0; p < P; ++p) {
            group_eval(p, nullptr);
            group_issue(p, 0, true, wv + 1 < waves);
            if (cancel_.load()) {
                sync();
                throw std::runtime_error("c6: cancelled");
            }
        }
}

void Engine::run_device_callback_to_end(DeviceCallback cb, void* user) {
    const int per_move = std::max(cfg_.num_simulation, 1);
    while (!all_done()) waves_device_callback(cb, user, per_move);
}

void Engine::external_move(int slot, int row, int col) {
    Dev& D = d_->dev;
    if (slot < 0 || slot >= D.G || row < 0 || row >= BOARD_SIZE || col < 0 || col >= BOARD_SIZE) throw std::runtime_error("c6: external_move out of range");
    sync();
    GameState gs = slot_state(slot);
    const uint32_t bit = bit_of_cell(row * BOARD_SIZE + col);
    if (gs.status != ST_ACTIVE) throw std::runtime_error("c6: external_move on a slot without an active game");
    if (((gs.B[bit >> 5] | gs.W[bit >> 5]) >> (bit & 31)) & 1u) throw std::runtime_error("already set position");
    k_external_move<<<1, 32, 0, stream_>>>(D, slot, row * BOARD_SIZE + col);
    launches_ += 1;
    C6_CUDA(cudaGetLastError());
    C6_CUDA(cudaStreamSynchronize(stream_));
}

void Engine::sync() {
    for (Group& gr : d_->groups)
        if (gr.s && gr.s != stream_) C6_CUDA(cudaStreamSynchronize(gr.s));
    C6_CUDA(cudaStreamSynchronize(stream_));
}

static void check_errors(DeviceArrays* d, cudaStream_t s) {
    C6_CUDA(cudaMemcpyAsync(d->h_gcount, d->dev.gcount, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    C6_CUDA(cudaStreamSynchronize(s));
    if (d->h_gcount[2]) {
        std::vector<GameState> gs(d->dev.G);
        cudaMemcpy(gs.data(), d->dev.gs, gs.size() * sizeof(GameState), cudaMemcpyDeviceToHost);
        int code = 0, slot = -1;
        for (int i = 0; i < d->dev.G; ++i)
            if (gs[i].status == ST_ERROR) {
                code = gs[i].err, slot = i;
                break;
            }
        char buf[320];
        if (code == ERR_EXP_FULL)
            std::snprintf(buf, sizeof buf, "c6: node pool exhausted in game slot %d (expanded list: %u entries, sized from num_simulation) even after a "
                          "restart from a fresh root", slot, d->dev.E);
        else
            std::snprintf(buf, sizeof buf, "c6: node pool exhausted in game slot %d (%s) even after a restart from a fresh root; raise log2_records / "
                          "C6_LOG2_RECORDS (now %u records per game)", slot, code == ERR_REC_FULL ? "record ring" : "link ring", d->dev.R);
        g_last_error = buf;
        throw std::runtime_error(buf);
    }
}

int64_t Engine::games_done() {
    sync();
    check_errors(d_, stream_);
    return (int64_t)d_->h_gcount[0];
}
bool Engine::all_done() { return games_done() >= total_games_; }

void Engine::run_device_to_end() {
    // a game lasts at most 225 stones; poll every few moves
    while (!all_done()) {
        if (cancel_.load()) throw std::runtime_error("c6: cancelled");
        for (int i = 0; i < 4; ++i) step_device();
    }
}

GameState Engine::slot_state(int slot) {
    GameState g;
    sync();
    C6_CUDA(cudaMemcpy(&g, d_->dev.gs + slot, sizeof g, cudaMemcpyDeviceToHost));
    return g;
}

void Engine::run_callback_to_end(Callback cb, bool debug) {
    (void)debug;
    const int per_move = std::max(cfg_.num_simulation, 1);
    while (!all_done()) waves_callback(cb, per_move);
}

float Engine::timed_steps_device(int k) {
    cudaEvent_t a, b;
    C6_CUDA(cudaEventCreate(&a));
    C6_CUDA(cudaEventCreate(&b));
    C6_CUDA(cudaStreamSynchronize(stream_));
    C6_CUDA(cudaEventRecord(a, stream_));
    time_kernels_ = true;
    for (int i = 0; i < k; ++i) step_device();
    time_kernels_ = false;
    C6_CUDA(cudaEventRecord(b, stream_));
    C6_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    C6_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a), cudaEventDestroy(b);
    for (size_t i = 0; i + 1 < kev_.size(); i += 2) {
        float kms = 0;
        C6_CUDA(cudaEventElapsedTime(&kms, kev_[i], kev_[i + 1]));
        sims_kernel_ms_ += kms, sims_kernel_launches_ += 1;
        cudaEventDestroy(kev_[i]), cudaEventDestroy(kev_[i + 1]);
    }
    kev_.clear();
    check_errors(d_, stream_);
    return ms;
}

float Engine::timed_waves_callback(Callback cb, int k, double* h2d, double* d2h) {
    cudaEvent_t a, b;
    C6_CUDA(cudaEventCreate(&a));
    C6_CUDA(cudaEventCreate(&b));
    sync();
    double h0 = h2d_bytes_, d0 = d2h_bytes_;
    // the timed region spans every group stream: the start event gates them all, the end event waits for them all
    C6_CUDA(cudaEventRecord(a, stream_));
    for (Group& gr : d_->groups)
        if (gr.s != stream_) C6_CUDA(cudaStreamWaitEvent(gr.s, a, 0));
    waves_callback(cb, k);
    for (Group& gr : d_->groups)
        if (gr.s != stream_) {
            C6_CUDA(cudaEventRecord(gr.ev, gr.s));
            C6_CUDA(cudaStreamWaitEvent(stream_, gr.ev, 0));
        }
    C6_CUDA(cudaEventRecord(b, stream_));
    C6_CUDA(cudaEventSynchronize(b));
    float ms = 0;
    C6_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a), cudaEventDestroy(b);
    if (h2d) *h2d = h2d_bytes_ - h0;
    if (d2h) *d2h = d2h_bytes_ - d0;
    check_errors(d_, stream_);
    return ms;
}


long long Engine::validate(unsigned long long* detail) {
    Dev& D = d_->dev;
    unsigned long long* dv = nullptr;
    C6_CUDA(cudaMalloc(&dv, 65 * sizeof(unsigned long long)));
    C6_CUDA(cudaMemsetAsync(dv, 0, 65 * sizeof(unsigned long long), stream_));
    k_validate<<<warps_grid(D.G), 128, 0, stream_>>>(D, dv);
    C6_CUDA(cudaGetLastError());
    unsigned long long h[65];
    C6_CUDA(cudaMemcpyAsync(h, dv, sizeof h, cudaMemcpyDeviceToHost, stream_));
    C6_CUDA(cudaStreamSynchronize(stream_));
    cudaFree(dv);
    if (detail) std::memcpy(detail, h + 1, 64 * sizeof(unsigned long long));
    return (long long)h[0];
}

// debug accessor (tests): copy raw per-slot arrays to the host.  what: 0 GameState, 1 expanded boards,
// 2 expanded keys, 3 expanded record indices.  Returns bytes copied.
size_t Engine::debug_read(int what, int slot, void* out, size_t max_bytes) {
    sync();
    Dev& D = d_->dev;
    const void* src = nullptr;
    size_t n = 0;
    if (what == 0) src = D.gs + slot, n = sizeof(GameState);
    else if (what == 1) src = D.expB + (size_t)slot * D.E * 16, n = (size_t)D.E * 64;
    else if (what == 2) src = D.exp_key + (size_t)slot * D.E, n = (size_t)D.E * 8;
    else if (what == 3) src = D.exp_rec + (size_t)slot * D.E, n = (size_t)D.E * 4;
    if (!src) return 0;
    n = std::min(n, max_bytes);
    C6_CUDA(cudaMemcpy(out, src, n, cudaMemcpyDeviceToHost));
    return n;
}

Stats Engine::stats() {
    sync();
    check_errors(d_, stream_);
    std::vector<GameState> gs(d_->dev.G);
    C6_CUDA(cudaMemcpy(gs.data(), d_->dev.gs, gs.size() * sizeof(GameState), cudaMemcpyDeviceToHost));
    Stats s{};
    for (auto& g : gs) {
        s.sims += g.sims, s.sum_ck += g.sum_ck, s.sum_cleaf += g.sum_cleaf, s.sum_depth += g.sum_depth;
        s.evals += g.evals, s.gc_copied += g.gc_copied;
        s.peak_rec = std::max(s.peak_rec, g.peak_rec), s.peak_lnk = std::max(s.peak_lnk, g.peak_lnk);
        s.degraded += g.degraded;
    }
    s.games_done = d_->h_gcount[0];
    return s;
}

std::vector<GameRecord> Engine::records() {
    sync();
    std::vector<GameRecord> r((size_t)total_games_);
    if (total_games_)
        C6_CUDA(cudaMemcpy(r.data(), d_->dev.records, r.size() * sizeof(GameRecord), cudaMemcpyDeviceToHost));
    return r;
}

void Engine::trace(std::vector<int32_t>& cv, std::vector<int32_t>& rn, std::vector<float>& rw) {
    sync();
    if (!cfg_.trace) throw std::runtime_error("c6: trace not enabled");
    size_t n = (size_t)total_games_ * CELLS;
    cv.resize(n * CELLS), rn.resize(n), rw.resize(n);
    C6_CUDA(cudaMemcpy(cv.data(), d_->dev.tr_visits, cv.size() * 4, cudaMemcpyDeviceToHost));
    C6_CUDA(cudaMemcpy(rn.data(), d_->dev.tr_n, rn.size() * 4, cudaMemcpyDeviceToHost));
    C6_CUDA(cudaMemcpy(rw.data(), d_->dev.tr_w, rw.size() * 4, cudaMemcpyDeviceToHost));
}

void Engine::trace_game(int64_t game, int32_t* cv, int32_t* rn, float* rw) {
    sync();
    if (!cfg_.trace) throw std::runtime_error("c6: trace not enabled");
    if (game < 0 || game >= total_games_) throw std::runtime_error("c6: trace_game: game index out of range");
    const size_t row = (size_t)game * CELLS;
    C6_CUDA(cudaMemcpy(cv, d_->dev.tr_visits + row * CELLS, (size_t)CELLS * CELLS * 4, cudaMemcpyDeviceToHost));
    C6_CUDA(cudaMemcpy(rn, d_->dev.tr_n + row, CELLS * 4, cudaMemcpyDeviceToHost));
    C6_CUDA(cudaMemcpy(rw, d_->dev.tr_w + row, CELLS * 4, cudaMemcpyDeviceToHost));
}

}  // namespace c6
