// Device-side building blocks of the self-play engine (one warp per game).
// Included only by c6_engine.cu.  Reference citations: LC = /root/reference/Connect6/libconnect6/src.
#pragma once
#include "c6_board.cuh"
#include "c6_engine.cuh"
#include "c6_rng.cuh"

namespace c6 {

constexpr unsigned FULL = 0xFFFFFFFFu;

struct Dev {
    Rec* rec;
    uint32_t* lnk;
    Slot* idx;       // [G][I] key table of the expanded nodes (key -> record)
    uint32_t* expB;  // [G][Ecap][16] bitboards of expanded nodes (8 words black, 8 white)
    uint32_t* exp_rec;
    uint64_t* exp_key;
    uint32_t* path;  // [G][K][232]: one row per leaf a game may have waiting
    GameState* gs;
    GameRecord* records;
    const uint64_t* Z;  // [2][256] Zobrist keys by bit position
    int* cursor;
    unsigned long long* gcount;  // [0] games finished, [1] next game offset to start, [2] error count
    // evaluator batch (callback paths): per pipeline group of n slots a region of n*8*226 floats at offset
    // g0*8*226, laid out [values n*8 | boards n*8*225]; boards in the API layout float[len][15][15] (in: boards,
    // out: policies)
    float* batch;                // device copy
    float* hbatch;               // the same layout in mapped pinned host memory (host-callback path)
    uint8_t* leaf_flag;          // [G] 0 nothing to evaluate, 1 Black to move at the leaf, 2 White
    int* gdone;                  // [groups] blocks of the running k_gather that have finished
    int* hsync;                  // mapped pinned host memory, per group {Black leaves, White leaves, seq, -}
    int32_t* tr_visits;
    int32_t* tr_n;
    float* tr_w;
    uint32_t R, L, I, E;  // capacities per game (powers of two except E)
    int G;
    long long total_games;
    uint32_t first_id;
    int num_sim;
    float eps;
    GammaParams gp;  // Dirichlet alpha and the sampler constants derived from it
    float cpuct;
    uint32_t seed;
    int trace;
    int max_moves;
    int K;          // Config::virtual_loss: leaves a game may have waiting at once (1 = the reference's schedule)
    int tie_break;  // 0: exact PUCT ties go to the lowest rank, 1: uniformly among the tied (Philox keyed (sim, depth))
};

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// memory helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Slot load_slot(const Slot* p) {  // L2-coherent (slots are CAS targets)
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(p));
    Slot s;
    s.key = ((uint64_t)v.y << 32) | v.x;
    s.idx = v.z;
    s.epoch = v.w;
    return s;
}
__device__ __forceinline__ Rec load_rec(const Rec* p) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    Rec r;
    r.n = (int32_t)v.x, r.w = __uint_as_float(v.y), r.p = __uint_as_float(v.z), r.link = v.w;
    return r;
}
__device__ __forceinline__ void store_rec(Rec* p, const Rec& r) {
    *reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)r.n, __float_as_uint(r.w), __float_as_uint(r.p), r.link);
}
__device__ __forceinline__ uint32_t hash_slot(uint64_t key) { return (uint32_t)(key ^ (key >> 32)); }

// Key table of the expanded nodes of one game.  Only the warp that owns the game touches it, one lane at a
// time, so plain L2 loads and a 16 B store are enough; the table holds at most E <= I/2 live entries.
__device__ __forceinline__ uint32_t xh_find(const Slot* idx, uint32_t Imask, uint32_t epoch, uint64_t key) {
    uint32_t pos = hash_slot(key) & Imask;
    for (;;) {
        const Slot s = load_slot(&idx[pos]);
        if (s.epoch != epoch) return NIL;
        if (s.key == key) return s.idx;
        pos = (pos + 1) & Imask;
    }
}
__device__ __forceinline__ void xh_insert(Slot* idx, uint32_t Imask, uint32_t epoch, uint64_t key, uint32_t val) {
    uint32_t pos = hash_slot(key) & Imask;
    while (__ldcg(&idx[pos].epoch) == epoch) pos = (pos + 1) & Imask;
    __stcg(reinterpret_cast<uint4*>(&idx[pos]), make_uint4((uint32_t)key, (uint32_t)(key >> 32), val, epoch));
}

struct Leaf {  // the board itself is in the warp's stage (WarpStage::B / W)
    uint64_t key;
    int stones;
    uint32_t rec;
    int depth;
    uint32_t last_bit;  // bit position of the last stone on the path (valid when depth >= 1)
    float leaf_w, leaf_p;  // W and P of the leaf's record as the descent read them (its N is 0)
};

struct Warp {  // per-warp view of one game slot: hot bases as pointers, the rest derived from the slot number on use
    Rec* rec;
    uint32_t* lnk;
    Slot* idx;
    GameState* gs;
    const uint64_t* Z;  // shared-memory copy
    uint32_t* pth;      // the path row in use (the waiting leaf being descended to / expanded)
    int lane;
    int g;
    __device__ __forceinline__ uint32_t* expB(const Dev& D) const { return D.expB + (size_t)g * D.E * 16; }
    __device__ __forceinline__ uint32_t* exp_rec(const Dev& D) const { return D.exp_rec + (size_t)g * D.E; }
    __device__ __forceinline__ uint64_t* exp_key(const Dev& D) const { return D.exp_key + (size_t)g * D.E; }
    __device__ __forceinline__ uint32_t* path(const Dev&) const { return pth; }
};

__device__ __forceinline__ Warp make_warp(const Dev& D, int g, const uint64_t* Zs, int lane) {
    Warp w;
    w.rec = D.rec + (size_t)g * D.R;
    w.lnk = D.lnk + (size_t)g * D.L;
    w.idx = D.idx + (size_t)g * D.I;
    w.gs = D.gs + g;
    w.Z = Zs;
    w.pth = D.path + (size_t)g * D.K * 232;
    w.lane = lane;
    w.g = g;
    return w;
}

__device__ __forceinline__ void set_error(Warp& w, const Dev& D, int code) {
    if (w.lane == 0) {
        w.gs->status = ST_ERROR;
        w.gs->err = code;
        atomicAdd(D.gcount + 2, 1ull);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// game start / collection / move
// ------------------------------------------------------------------------------------------------
__device__ inline void init_game(const Dev& D, Warp& w, uint32_t game_id) {
    GameState& g = *w.gs;
    if (w.lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g.B[j] = 0, g.W[j] = 0;
        g.key = 0, g.root = 0, g.stones = 0, g.status = ST_ACTIVE, g.winner = 0, g.err = 0;
        g.game_id = game_id;
        g.epoch += 1;
        g.rec_head = 1, g.rec_tail = 0, g.lnk_head = 0, g.lnk_tail = 0, g.exp_count = 0;
        g.sim_index = 0, g.sims_this_move = 0, g.n_moves = 0, g.n_pend = 0, g.retried = 0;
        Rec r;
        r.n = 0, r.w = 0.f, r.p = 0.f, r.link = NIL;
        store_rec(&w.rec[0], r);
#pragma unroll
        for (int j = 0; j < 8; ++j) g.orph[j] = 0;
    }
    __syncwarp();
}

// A game's node pool is full (a search that concentrates on one line keeps a large subtree from stone to stone; the
// reference's HashMap has no such limit): instead of failing every game of the batch, this game drops its tree and searches
// the current board again from an unexpanded root — what a fresh AlphaZero instance would do.  Counted in Stats::degraded.
__device__ inline void fresh_root(const Dev& D, Warp& w) {
    GameState& g = *w.gs;
    if (w.lane == 0) {
        g.epoch += 2;  // empties the key table (the collector may have inserted under epoch + 1 already)
        g.rec_tail = g.rec_head, g.lnk_tail = g.lnk_head, g.exp_count = 0;
        const uint32_t root = g.rec_head & (D.R - 1);
        g.rec_head += 1;
        Rec r;
        r.n = 0, r.w = 0.f, r.p = 0.f, r.link = NIL;
        store_rec(&w.rec[root], r);
        g.root = root;
#pragma unroll
        for (int j = 0; j < 8; ++j) g.orph[j] = 0;
        g.degraded += 1;
        g.n_pend = 0;
    }
    __syncwarp();
}

// After a real move: keep exactly the records that can ever be reached again — expanded nodes whose
// board contains the new root board, and their children — compacted into the free part of the rings
// (the reference keeps every deeper node forever, LC/policy/alphazero_policy/mod.rs:474-487; nodes
// that do not contain the root can never be looked up again, so dropping them changes nothing).
__device__ inline void collect(const Dev& D, Warp& w) {
    GameState& g = *w.gs;
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1, Im = D.I - 1;
    const uint32_t old_rec_head = g.rec_head, old_lnk_head = g.lnk_head;
    const uint32_t old_rec_tail = g.rec_tail, old_lnk_tail = g.lnk_tail;
    uint32_t rec_head = old_rec_head, lnk_head = old_lnk_head;
    const uint32_t epoch = g.epoch + 1;
    uint32_t rootB[8], rootW[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) rootB[j] = g.B[j], rootW[j] = g.W[j];
    const uint32_t n_exp = g.exp_count;
    uint32_t wcur = 0;
    unsigned long long copied = 0;
    bool overflow = false;
    // Most expanded nodes lack the stone just played and are dropped: the containment test runs for 32 entries at a
    // time, one per lane; only the survivors are then handled by the whole warp, in list order.
    for (uint32_t base = 0; base < n_exp && !overflow; base += 32) {
      bool keep = false;
      if (base + (uint32_t)lane < n_exp) {
          const uint4* q = reinterpret_cast<const uint4*>(w.expB(D) + (size_t)(base + (uint32_t)lane) * 16);
          const uint4 b0 = q[0], b1 = q[1], w0 = q[2], w1 = q[3];
          const uint32_t xb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w}, xw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          keep = true;
#pragma unroll
          for (int j = 0; j < 8; ++j) keep = keep && ((xb[j] & rootB[j]) == rootB[j]) && ((xw[j] & rootW[j]) == rootW[j]);
      }
      unsigned todo = __ballot_sync(FULL, keep);
      while (todo) {
        const uint32_t ei = base + (uint32_t)(__ffs((int)todo) - 1);
        todo &= todo - 1;
        uint32_t eb[8], ew[8];
        int est = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            eb[j] = w.expB(D)[(size_t)ei * 16 + j];
            ew[j] = w.expB(D)[(size_t)ei * 16 + 8 + j];
            est += __popc(eb[j]) + __popc(ew[j]);
        }
        const uint32_t erec = w.exp_rec(D)[ei];
        const uint64_t ekey = w.exp_key(D)[ei];
        const int C = CELLS - est;
        if (rec_head - old_rec_tail + (uint32_t)C + 1 > D.R || lnk_head - old_lnk_tail + (uint32_t)C + 3 > D.L) {
            overflow = true;
            break;
        }
        // the node's own record
        Rec old = load_rec(&w.rec[erec]);
        uint32_t new_e, old_link;
        if (old.link == FWD) {
            new_e = (uint32_t)old.n;
            old_link = w.rec[new_e].link;
        } else {
            new_e = rec_head & Rm;
            rec_head += 1;
            old_link = old.link;
            if (lane == 0) {
                store_rec(&w.rec[new_e], old);
                Rec f;
                f.n = (int32_t)new_e, f.w = 0.f, f.p = 0.f, f.link = FWD;
                store_rec(&w.rec[erec], f);
            }
            copied += 1;
        }
        if (lane == 0) xh_insert(w.idx, Im, epoch, ekey, new_e);  // expanded nodes stay findable by key
        __syncwarp();
        const uint32_t new_link = (lnk_head + 3u) & ~3u;  // link runs start on 16-byte boundaries (four links per cp.async)
        lnk_head = new_link + (uint32_t)C;
        if (lane == 0) w.rec[new_e].link = new_link & Lm;
        // children
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ~(eb[j] | ew[j]) & valid_mask(j);
        uint32_t ci[8];
        Rec ch[8];
        int pre = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ci[j] = NIL;
            if ((e[j] >> lane) & 1u) {
                int k = pre + __popc(e[j] & ((1u << lane) - 1u));
                ci[j] = w.lnk[(old_link + (uint32_t)k) & Lm] & Rm;
            }
            pre += __popc(e[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (ci[j] != NIL) ch[j] = load_rec(&w.rec[ci[j]]);
        pre = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            bool mine = (e[j] >> lane) & 1u;
            bool need = mine && ch[j].link != FWD;
            unsigned bal = __ballot_sync(FULL, need);
            if (mine) {
                uint32_t ni;
                if (need) {
                    ni = (rec_head + __popc(bal & ((1u << lane) - 1u))) & Rm;
                    store_rec(&w.rec[ni], ch[j]);
                    Rec f;
                    f.n = (int32_t)ni, f.w = 0.f, f.p = 0.f, f.link = FWD;
                    store_rec(&w.rec[ci[j]], f);
                } else {
                    ni = (uint32_t)ch[j].n;
                }
                int k = pre + __popc(e[j] & ((1u << lane) - 1u));
                w.lnk[(new_link + (uint32_t)k) & Lm] = ni;
            }
            rec_head += __popc(bal);
            copied += __popc(bal);
            pre += __popc(e[j]);
        }
        // keep the entry (in-place compaction of the expanded list: wcur <= ei)
        if (lane < 16) w.expB(D)[(size_t)wcur * 16 + lane] = w.expB(D)[(size_t)ei * 16 + lane];
        if (lane == 0) w.exp_rec(D)[wcur] = new_e, w.exp_key(D)[wcur] = ekey;
        wcur += 1;
        __syncwarp();
      }
    }
    if (overflow) return fresh_root(D, w);  // the kept subtree does not fit beside the old one: the next search starts afresh
    // the root itself (unexpanded roots are not on the expanded list)
    Rec rr = load_rec(&w.rec[g.root]);
    uint32_t new_root;
    if (rr.link == FWD) {
        new_root = (uint32_t)rr.n;
    } else {
        new_root = rec_head & Rm;
        rec_head += 1;
        copied += 1;
        if (lane == 0) store_rec(&w.rec[new_root], rr);
    }
    __syncwarp();
    // Orphans: a kept expanded node X (other than the root) is reached by later expansions through one of its
    // expanded parents X - s (s a stone of the colour placed last, not on the root board).  If none is left —
    // its former parent lacked the stone just played — flag X's stone count, so that expansions creating
    // children of that size look them up by key (c6_search.cuh: resolve_rare).
    uint32_t orph = 0;  // lane j < 8 holds word j
    for (uint32_t ei = 0; ei < wcur; ++ei) {
        int est = 0;
        uint32_t eb[8], ew[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            eb[j] = w.expB(D)[(size_t)ei * 16 + j];
            ew[j] = w.expB(D)[(size_t)ei * 16 + 8 + j];
            est += __popc(eb[j]) + __popc(ew[j]);
        }
        if (est == g.stones) continue;  // the root itself
        const uint64_t ekey = w.exp_key(D)[ei];
        const int m = turn_of(est - 1);  // colour of the stone placed last
        bool safe = false;
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            const uint32_t cand = m < 0 ? (eb[j] & ~rootB[j]) : (ew[j] & ~rootW[j]);
            if ((cand >> lane) & 1u) safe = safe || xh_find(w.idx, Im, epoch, ekey ^ w.Z[(m > 0 ? 256 : 0) + 32 * j + lane]) != NIL;
        }
        if (!__any_sync(FULL, safe) && lane == (est >> 5)) orph |= 1u << (est & 31);
    }
    if (lane < 8) g.orph[lane] = orph;
    if (lane == 0) {
        g.root = new_root;
        g.epoch = epoch;
        g.rec_tail = old_rec_head, g.rec_head = rec_head;
        g.lnk_tail = old_lnk_head, g.lnk_head = lnk_head;
        g.exp_count = wcur;
        g.gc_copied += copied;
    }
    __syncwarp();
}

__device__ inline void finish_game(const Dev& D, Warp& w, int winner, int finished) {
    GameState& g = *w.gs;
    const int lane = w.lane;
    GameRecord* out = D.records + (g.game_id - D.first_id);
    for (int i = lane; i < 232; i += 32) out->moves[i] = g.moves[i];
    unsigned long long ng = 0;
    if (lane == 0) {
        out->winner = (int8_t)winner;
        out->n_moves = (uint8_t)g.n_moves;
        out->finished = (uint8_t)finished;
        out->game_id = g.game_id;
        g.winner = winner;
        atomicAdd(D.gcount + 0, 1ull);
        ng = atomicAdd(D.gcount + 1, 1ull);
    }
    ng = __shfl_sync(FULL, ng, 0);
    __syncwarp();
    if ((long long)ng < D.total_games)
        init_game(D, w, D.first_id + (uint32_t)ng);
    else if (lane == 0)
        g.status = ST_DONE;
    __syncwarp();
}

// policy() + the tail of next() + one iteration of Agent::play:
// LC/policy/alphazero_policy/mod.rs:433-452,471-488 and LC/agent/agent_impl/mod.rs:131-171.
__device__ inline void finish_move(const Dev& D, Warp& w) {
    GameState& g = *w.gs;
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1;
    Rec r = load_rec(&w.rec[g.root]);
    if (lane == 0) g.sims_this_move = 0, g.retried = 0;
    uint32_t B[8], W[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) B[j] = g.B[j], W[j] = g.W[j];
    const int stones = g.stones;
    const uint32_t gi = g.game_id - D.first_id;
    int32_t* trv = D.trace ? D.tr_visits + ((size_t)gi * CELLS + g.n_moves) * CELLS : nullptr;
    if (D.trace && lane == 0) {
        D.tr_n[(size_t)gi * CELLS + g.n_moves] = r.n;
        D.tr_w[(size_t)gi * CELLS + g.n_moves] = r.w;
    }
    if (r.n == 0) {  // no children: policy() -> None -> Agent::play ends with winner None (:135-137)
        if (trv)
            for (int i = lane; i < CELLS; i += 32) trv[i] = -1;
        return finish_game(D, w, 0, 1);
    }
    uint32_t e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = ~(B[j] | W[j]) & valid_mask(j);
    uint32_t ci[8];
    int cn[8];
    int pre = 0, s_local = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        ci[j] = NIL, cn[j] = 0;
        if ((e[j] >> lane) & 1u) {
            int k = pre + __popc(e[j] & ((1u << lane) - 1u));
            ci[j] = w.lnk[(r.link + (uint32_t)k) & Lm] & Rm;
        }
        pre += __popc(e[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (ci[j] != NIL) cn[j] = w.rec[ci[j]].n, s_local += cn[j];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s_local += __shfl_xor_sync(FULL, s_local, off);
    const float visit_sum = (float)s_local;
    if (trv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t bit = 32u * j + lane;
            if ((valid_mask(j) >> lane) & 1u) trv[cell_of_bit(bit)] = (ci[j] != NIL) ? cn[j] : -1;
        }
    }
    float bp = 0.f;
    uint32_t bb = NIL, bc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (ci[j] != NIL) {
            float visit = (float)cn[j];
            float p = visit / (visit_sum - visit + 1.f);  // :444-447
            if (bb == NIL || p >= bp) bp = p, bb = 32u * j + lane, bc = ci[j];  // max_by keeps the LAST maximum
        }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        float op = __shfl_xor_sync(FULL, bp, off);
        uint32_t ob = __shfl_xor_sync(FULL, bb, off), oc = __shfl_xor_sync(FULL, bc, off);
        bool take = ob != NIL && (bb == NIL || op > bp || (op == bp && ob > bb));
        if (take) bp = op, bb = ob, bc = oc;
    }
    // Game::set (LC/game/game_impl/mod.rs:148-169) + is_game_end
    const int turn = turn_of(stones);
    const uint32_t wj = bb >> 5, m = 1u << (bb & 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t mm = (j == (int)wj) ? m : 0u;
        if (turn < 0) B[j] |= mm; else W[j] |= mm;
    }
    const int winner = win_check(B, W);
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g.B[j] = B[j], g.W[j] = W[j];
        g.key ^= w.Z[(turn > 0 ? 256 : 0) + bb];
        g.stones = stones + 1;
        g.moves[g.n_moves] = (uint8_t)cell_of_bit(bb);
        g.n_moves += 1;
        g.root = bc;
    }
    __syncwarp();
    if (winner != 0) return finish_game(D, w, winner, 1);
    if (stones + 1 == CELLS) return finish_game(D, w, 0, 1);  // next() on a full board returns None
    if (D.max_moves > 0 && g.n_moves >= D.max_moves) return finish_game(D, w, 0, 0);
    collect(D, w);
}

// A stone placed by somebody else (the White side of cpp_play_with, or any caller of AlphaZero::next with
// a game that moved on): Game::set, then what `init` does on the next search (alphazero_policy/mod.rs:
// 302-306: the node of the new board is reused with its statistics if the table has it, else created),
// then the same collection as after an own move.
__device__ inline void external_move(const Dev& D, Warp& w, uint32_t bit) {
    GameState& g = *w.gs;
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Im = D.I - 1;
    uint32_t B[8], W[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) B[j] = g.B[j], W[j] = g.W[j];
    const int stones = g.stones;
    const int turn = turn_of(stones);
    const uint32_t wj = bit >> 5, m = 1u << (bit & 31);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t mm = (j == (int)wj) ? m : 0u;
        if (turn < 0) B[j] |= mm; else W[j] |= mm;
    }
    const int winner = win_check(B, W);
    const uint64_t key = g.key ^ w.Z[(turn > 0 ? 256 : 0) + bit];
    if (g.rec_head - g.rec_tail + 1 > D.R) return set_error(w, D, ERR_REC_FULL);
    uint32_t root = 0;
    if (lane == 0) {
        // an expanded root links the record of every child board; otherwise the new board may still be an
        // expanded node of an earlier search (key table); otherwise it is new
        const Rec rr = load_rec(&w.rec[g.root]);
        uint32_t val = NIL;
        if (rr.n >= 1) {
            int k = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t e = ~(g.B[j] | g.W[j]) & valid_mask(j);
                if (j < (int)wj) k += __popc(e);
                if (j == (int)wj) k += __popc(e & (m - 1u));
            }
            val = w.lnk[(rr.link + (uint32_t)k) & (D.L - 1)] & Rm;
        } else {
            val = xh_find(w.idx, Im, g.epoch, key);
        }
        if (val != NIL) {
            root = val;
        } else {
            root = g.rec_head & Rm;
            g.rec_head += 1;
            Rec r;
            r.n = 0, r.w = 0.f, r.p = 0.f, r.link = NIL;
            store_rec(&w.rec[root], r);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) g.B[j] = B[j], g.W[j] = W[j];
        g.key = key;
        g.stones = stones + 1;
        g.moves[g.n_moves] = (uint8_t)cell_of_bit(bit);
        g.n_moves += 1;
        g.root = root;
        g.sims_this_move = 0;
    }
    __syncwarp();
    if (winner != 0) return finish_game(D, w, winner, 1);
    if (stones + 1 == CELLS) return finish_game(D, w, 0, 1);
    if (D.max_moves > 0 && g.n_moves >= D.max_moves) return finish_game(D, w, 0, 0);
    collect(D, w);
}

#endif  // __CUDACC__

}  // namespace c6
