// The per-simulation hot path (one warp per game): select -> leaf evaluation -> expand -> backup.
//   select / maximum_from   LC/policy/alphazero_policy/mod.rs:243-323  (looped as in search() :417-420)
//   expand / get_from       LC/policy/alphazero_policy/mod.rs:218-234,331-381, augment.rs:60-92
//   update                  LC/policy/alphazero_policy/mod.rs:386-408
//
// Shaped by the ncu evidence under profiles/: the first fused kernel (r01a) was 105 KB of SASS with
// `stall_no_instruction` as its top stall (instruction-cache thrash: 20 warps at unrelated places of an
// unrolled body), spent two thirds of its instructions in the f64 Gamma sampler and pushed 100+ GB per
// launch through local memory; up to r01e every per-child loop ran one round per board WORD (8 rounds, 216 of
// 256 lanes busy in the opening, far fewer later).  Hence:
//   (1) every per-child phase is a ROLLED loop over child RANKS (k = lane, lane + 32, ... < C: ceil(C/32) rounds)
//       that reads its operands from a per-warp shared-memory stage filled by cp.async, all requests of a
//       phase in flight before the first is waited for — one copy of the PUCT, Gamma and Philox code each;
//   (2) everything a move needs from the game state lives in registers (`Cursor`);
//   (3) an expansion does not hash its children: a child shared with another parent ("transposition",
//       `entry(hash).or_insert`, :367-370) is found by following the link run of the expanded sibling
//       the select step has just read (r01d: the per-child board-key index cost 60 KB of DRAM traffic per
//       simulation against 12.5 KB of algorithmic bytes); only expanded nodes are kept in a key table;
//   (4) no stack frame: every helper on the default path is inlined into a single call site.
#pragma once
#include "c6_device.cuh"

namespace c6 {

// Per-move copy of the mutable game bookkeeping, identical in every lane of the warp: what the inner loops need in
// registers; the rest — bounds that do not change during a move and the statistics counters — sits in the warp's stage
// (WarpStage::meta, lane 0 updates the counters).  The r02a capture had these spilled: 219 M local loads / stores per
// launch whose 128-byte lines competed with the node pool for L1 and L2 (22 % of the L2 sectors), and long-scoreboard
// stalls on lines as trivial as `sum_ck += C`.
struct Cursor {
    uint32_t rec_head, lnk_head, exp_count, sim_index, game_id;
    int err;
};
struct WarpMeta {
    uint32_t rec_tail, lnk_tail, epoch;
    uint32_t orph_win, stones0;  // GameState::orph bits of stone counts stones0 .. stones0 + 31 (stones0 = the root's)
    uint32_t sum_ck, sum_cleaf, sum_depth, evals, sims, peak_rec, peak_lnk;
};

template <class Stage>
__device__ __forceinline__ void load_cursor(const GameState& g, Cursor& c, Stage& st, int lane) {
    c.rec_head = g.rec_head, c.lnk_head = g.lnk_head, c.exp_count = g.exp_count, c.sim_index = g.sim_index;
    c.game_id = g.game_id;
    c.err = 0;
    __syncwarp();
    if (lane == 0) {
        WarpMeta& m = st.meta;
        m.rec_tail = g.rec_tail, m.lnk_tail = g.lnk_tail, m.epoch = g.epoch;
        m.sum_ck = 0, m.sum_cleaf = 0, m.sum_depth = 0, m.evals = 0, m.sims = 0;
        m.peak_rec = g.peak_rec, m.peak_lnk = g.peak_lnk;
        m.stones0 = (uint32_t)g.stones;
        const uint32_t wi = m.stones0 >> 5;
        m.orph_win = __funnelshift_r(g.orph[wi], wi + 1 < 8 ? g.orph[wi + 1] : 0u, m.stones0 & 31);
    }
    __syncwarp();
}
// is some expanded node with `s1` stones flagged as having lost its expanded parents (GameState::orph)?
__device__ __forceinline__ bool orphan_flag(const GameState& g, const WarpMeta& m, uint32_t s1) {
    const uint32_t j = s1 - m.stones0;
    return j < 32 ? ((m.orph_win >> j) & 1u) != 0 : ((g.orph[s1 >> 5] >> (s1 & 31)) & 1u) != 0;
}
template <class Stage>
__device__ __forceinline__ void store_cursor(const Dev& D, GameState& g, const Cursor& c, Stage& st, int lane) {
    __syncwarp();
    if (lane == 0) {
        const WarpMeta& m = st.meta;
        g.rec_head = c.rec_head, g.lnk_head = c.lnk_head, g.exp_count = c.exp_count, g.sim_index = c.sim_index;
        g.sum_ck += m.sum_ck, g.sum_cleaf += m.sum_cleaf, g.sum_depth += m.sum_depth, g.evals += m.evals, g.sims += m.sims;
        g.peak_rec = m.peak_rec, g.peak_lnk = m.peak_lnk;
        if (c.err) {
            g.status = ST_ERROR, g.err = c.err;
            atomicAdd(D.gcount + 2, 1ull);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// per-warp shared-memory stage
// ------------------------------------------------------------------------------------------------
constexpr int PATH_CACHE = 12;  // path records kept in shared memory (deeper ones are re-read by backup)
// NOISE: 0 = epsilon == 0 (the draw is multiplied by 0: skipped), 1 = Dirichlet alpha < 1, 2 = alpha >= 1
// One stage per warp, indexed by child RANK (position in the node's link run = row-major order of its empty
// cells): slot [k >> 5][k & 31] belongs to child k and is handled by lane k & 31 in round k >> 5, so a node with
// C children costs ceil(C/32) rounds of every per-child loop — not one round per board word — and the link run
// is read with fully coalesced requests.  Both memory phases fill the stage with cp.async from ROLLED loops
// (every request of a lane is in flight before the first is waited for), so neither the addresses nor the
// results occupy registers and the code of a phase exists once.
template <int NOISE>
struct WarpStage {
    uint4 child[8][32];  // select: {N, W, P, link} of child k; expand reads the leaf's siblings here
    union {
        uint32_t ci[8][32];  // select: record index of child k
        float prior[8][32];  // expand: prior of new child k
    };
    union {
        float lw[8][32];      // select: log2 Gamma draw of child k, then its softmax weight, then its PUCT score
        uint32_t pos[8][32];  // expand: record index of child k
    };
    uint32_t B[8], W[8];  // bit boards of the node the descent stands on / of the leaf being expanded (warp-uniform data)
    uint32_t empty[8];  // empty-cell mask of the current node, one board word each
    uint8_t cell[256];  // expand: bit position of the empty cell of rank k; select: ranks whose draw needs the slow path
    // fused kernel only: what this simulation's descent has already read of the records on its path, so that backup
    // stores without loading (the game's warp is the only writer of its pool), and the root's record across simulations
    uint4 prec[PATH_CACHE];
    uint32_t pidx[PATH_CACHE];
    uint4 rootrec;
    WarpMeta meta;  // per-move bounds and statistics counters of the game (see Cursor)
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// L2 only: index slots are claimed with atomics, an L1 copy could be stale
__device__ __forceinline__ void cp_async16_cg(void* smem_dst, const void* gmem_src) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
// Zobrist keys: 4 KB read by every expansion.  The fused kernel leaves ~40 KB of L1 per SM beside its
// shared-memory stages and streams node records through it, so the table is kept with an evict-last hint.
#ifndef C6_Z_EVICT_LAST
#define C6_Z_EVICT_LAST 1
#endif
__device__ __forceinline__ uint64_t ldz(const uint64_t* p) {
#if C6_Z_EVICT_LAST
    uint64_t v;
    asm("ld.global.nc.L1::evict_last.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}
#ifndef C6_CHILD_CG
#define C6_CHILD_CG 1  // child records bypass L1 (each is read once per step; tune_h: +1%)
#endif
#ifndef C6_LINK16
#define C6_LINK16 1  // link runs fetched four links per 16-byte request (runs are 16-byte aligned in the ring)
#endif
#ifndef C6_SQRT_TAB
#define C6_SQRT_TAB 0  // sqrt(S - n) of the PUCT pass from a 1,024-entry shared-memory table (same bits; A/B: scripts/gpu_ab.sh)
#endif
#ifndef C6_W_UNROLL
#define C6_W_UNROLL 1
#endif
#ifndef C6_PUCT_UNROLL
#define C6_PUCT_UNROLL 1
#endif
#ifndef C6_EV_UNROLL
#define C6_EV_UNROLL 1  // the evaluator's two Philox blocks per cell interleaved: +2.2 % (r02f A/B, scripts/gpu_tune3.sh)
#endif
#ifndef C6_EV_UNROLL
#define C6_EV_UNROLL 1  // the evaluator's two Philox blocks per cell interleaved: +2.2 % (r02f A/B, scripts/gpu_tune3.sh)
#endif
#ifndef C6_NOISE_UNROLL
#define C6_NOISE_UNROLL 1
#endif
#ifndef C6_PRIOR_UNROLL
#define C6_PRIOR_UNROLL 1
#endif
#define C6_PRAGMA_(x) _Pragma(#x)
#define C6_UNROLL(n) C6_PRAGMA_(unroll n)
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// The same copies with an L2 eviction policy.  A game's ROOT is descended through by every simulation: its link run and its
// children's records (~4 KB per game, 17 MB for 4,096 games) are the one part of the node pool that is re-read 800 times per
// stone, while everything deeper streams through; the root step asks L2 to keep them (evict_last), so they are served from
// L2 instead of DRAM (r02c capture: 36.9 GB of DRAM reads per launch at an L2 hit rate of 33 %).
#ifndef C6_ROOT_L2
#define C6_ROOT_L2 2  // 0: no hints; 1: root step evict_last; 2: also evict_first below the root.  r02j A/B (gpurun_out/r02j,
                      // profiles/r02j_l2_hints.md): DRAM traffic per launch 54.9 GB -> 51.1 GB -> 48.9 GB (1.34 x -> 1.20 x the
                      // algorithmic bytes), time unchanged (+0.2 % / +0.8 %: the kernel is issue-bound)
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void cp_async4_hint(void* smem_dst, const void* gmem_src, uint64_t pol) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_ca_hint(void* smem_dst, const void* gmem_src, uint64_t pol) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_cg_hint(void* smem_dst, const void* gmem_src, uint64_t pol) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(pol) : "memory");
}

// Memory phase of a select step: for the node with C children whose link run starts at `link`, record indices
// into st.ci, then records into st.child; ceil(C/32) requests of each kind in flight per lane.
template <int NOISE>
__device__ __forceinline__ void stage_children(const Warp& w, WarpStage<NOISE>& st, uint32_t link, int C, int lane, uint32_t Rm,
                                               uint32_t Lm) {
#pragma unroll 1
    for (int k = 4 * lane; k < C; k += 128) cp_async16(&st.ci[0][k], &w.lnk[(link + (uint32_t)k) & Lm]);  // 16-byte aligned runs
    cp_async_wait_all();
    __syncwarp();  // a lane's links were fetched by another lane
#pragma unroll 1
    for (int k = lane; k < C; k += 32) {
#if C6_CHILD_CG
        cp_async16_cg(&st.child[0][k], &w.rec[st.ci[0][k] & Rm]);
#else
        cp_async16(&st.child[0][k], &w.rec[st.ci[0][k] & Rm]);
#endif
    }
    cp_async_wait_all();
}

// Bit position of the empty cell of rank k (warp-uniform k) from the empty masks in st.empty.
template <int NOISE>
__device__ __forceinline__ uint32_t kth_empty(const WarpStage<NOISE>& st, uint32_t k, uint32_t lt, int lane) {
    uint32_t bit = 0, pre = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = st.empty[j];
        const uint32_t c = (uint32_t)__popc(e);
        if (k - pre < c) {  // warp-uniform; k - pre wraps to a huge value once pre > k
            const bool hit = ((e >> lane) & 1u) && (uint32_t)__popc(e & lt) == k - pre;
            bit = 32u * j + (uint32_t)(__ffs((int)__ballot_sync(FULL, hit)) - 1);
        }
        pre += c;
    }
    return bit;
}

// ------------------------------------------------------------------------------------------------
// select
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ Rec rec_of(const uint4& v) {
    Rec r;
    r.n = (int32_t)v.x, r.w = __uint_as_float(v.y), r.p = __uint_as_float(v.z), r.link = v.w;
    return r;
}
__device__ __forceinline__ uint4 bits_of(const Rec& r) { return make_uint4((uint32_t)r.n, __float_as_uint(r.w), __float_as_uint(r.p), r.link); }

// Warp-wide maximum of floats through the integer reduction unit (redux.sync): floats map to unsigned keys that order the
// same way (-0 is canonicalised to +0 first: the two compare equal as floats and must get one key).  No NaNs here.
__device__ __forceinline__ uint32_t float_key(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k); }

// Four rounds of the noise draw of a select step: children k, k + 32, k + 64, k + 96 of this lane share a Philox block.
template <int NOISE>
__device__ __forceinline__ void noise_rounds(const Dev& D, const Cursor& cs, WarpStage<NOISE>& st, const MathTab& mt, int depth, int C,
                                             int lane, int& k, int kend, float& lmax, unsigned& pend) {
    const uint32_t nk0 = D.seed ^ DOMAIN_NOISE;
    if (NOISE == 1) {
#pragma unroll 1
        for (; k < kend; k += 128) {
            const Philox4 r = philox4x32_10(cs.sim_index, (uint32_t)depth, (uint32_t)(((k >> 7) << 5) | lane), 0u, nk0, cs.game_id);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                float l;
                const bool ok = gamma_log2_fast(mt, D.gp, w[h], l);
                if (k + 32 * h < C) {
                    st.lw[0][k + 32 * h] = l;
                    if (ok) lmax = fmaxf(lmax, l); else pend |= (1u << h) << (k >> 5);
                }
            }
        }
    } else {
#pragma unroll 1
        for (; k < kend; k += 128)
            for (int kk = k; kk < C && kk <= k + 96; kk += 32) {
                const float l = gamma_large_log2(D.gp.shape, cs.sim_index, (uint32_t)depth, (uint32_t)kk, nk0, cs.game_id);
                st.lw[0][kk] = l;
                lmax = fmaxf(lmax, l);
            }
    }
}

// FUSED: called from k_sims_device, where the same warp runs descent, expansion and backup of a simulation back to back.
// The board of the node the descent stands on lives in the warp's stage (st.B / st.W, warp-uniform data: one copy per warp
// instead of sixteen registers per lane).
template <int NOISE, bool FUSED = false>
__device__ __forceinline__ void descend(const Dev& D, Warp& w, Cursor& cs, Leaf& L, WarpStage<NOISE>& st, const MathTab& mt) {
    const bool tie_philox = D.tie_break != 0;
    const GameState& g = *w.gs;
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1;
    const uint32_t lt = (1u << lane) - 1u;
    if (lane < 8) st.B[lane] = g.B[lane], st.W[lane] = g.W[lane];
    L.key = g.key;
    L.stones = g.stones;
    L.last_bit = 0;
    uint32_t cur = g.root;
    int depth = 0;
    const float eps = D.eps, one_m_eps = 1.f - D.eps, cpuct = D.cpuct;
    Rec r = FUSED ? rec_of(st.rootrec) : load_rec(&w.rec[cur]);
    for (;;) {
        if (r.n == 0) break;  // no children: leaf (:251-253)
        if (lane == 0) {
            if (FUSED && depth < PATH_CACHE) st.pidx[depth] = cur, st.prec[depth] = bits_of(r);
            if (!FUSED || depth >= PATH_CACHE) w.path(D)[depth] = cur;
        }
        const int C = CELLS - L.stones;
        const int turn = turn_of(L.stones);
        if (lane == 0) st.meta.sum_ck += (uint32_t)C;
        uint32_t chosen_bit = 0, child = 0;
        __syncwarp();  // st.B / st.W of this node are in place
        if (lane < 8) st.empty[lane] = ~(st.B[lane] | st.W[lane]) & (lane < 7 ? 0x7FFF7FFFu : 0x00007FFFu);
        if (C == 1) {  // single child: taken without scoring (:254-257)
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t e = st.empty[j];
                if (e) chosen_bit = 32 * j + (__ffs((int)e) - 1);
            }
            child = w.lnk[r.link & Lm];
            r = load_rec(&w.rec[child & Rm]);
        } else {
            // Memory phases and noise phase interleaved.  The link run, then the child records, are two dependent trips
            // (ceil(C/32) requests of each kind in flight per lane, cp.async straight into the stage).  The Dirichlet(alpha; C)
            // draw made at every node with C >= 2 (:262,283) depends on nothing but C: log2 Gamma draws keyed (sim, depth,
            // child k), two per Philox block.  A node that has been descended through before (N >= 2) has a visited child,
            // so its noise is certainly needed: the draws run under the two trips.  A node with N == 1 has usually no
            // visited child (see below) and waits for the records first.
            const bool early = NOISE != 0 && r.n >= 2;
            float lmax = -3.0e38f;
            unsigned pend = 0;
            int s_local = 0;
            int kn = lane;  // next rank this lane draws
#if C6_ROOT_L2
            const bool hinted = C6_ROOT_L2 == 2 || depth == 0;
            const uint64_t pol = depth == 0 ? l2_policy_evict_last() : l2_policy_evict_first();
#else
            const bool hinted = false;
            const uint64_t pol = 0;
#endif
            // the link run: 16-byte aligned in the ring, four links per request (a run's last request may read up to three
            // entries past it: unused)
#if C6_LINK16
            if (hinted) {
#pragma unroll 1
                for (int k = 4 * lane; k < C; k += 128) cp_async16_ca_hint(&st.ci[0][k], &w.lnk[(r.link + (uint32_t)k) & Lm], pol);
            } else {
#pragma unroll 1
                for (int k = 4 * lane; k < C; k += 128) cp_async16(&st.ci[0][k], &w.lnk[(r.link + (uint32_t)k) & Lm]);
            }
#else
            if (hinted) {
#pragma unroll 1
                for (int k = lane; k < C; k += 32) cp_async4_hint(&st.ci[0][k], &w.lnk[(r.link + (uint32_t)k) & Lm], pol);
            } else {
#pragma unroll 1
                for (int k = lane; k < C; k += 32) cp_async4(&st.ci[0][k], &w.lnk[(r.link + (uint32_t)k) & Lm]);
            }
#endif
#pragma unroll 1
            for (int part = 0; part < 3; ++part) {
                if (NOISE && (part < 2 ? early : (!early && s_local != 0))) noise_rounds(D, cs, st, mt, depth, C, lane, kn, part == 0 ? min(C, 128) : C, lmax, pend);
                if (part == 2) break;
                cp_async_wait_all();
                if (part == 0) {
                    __syncwarp();  // a lane's links were fetched by another lane (four per request)
                    if (hinted) {
#pragma unroll 1
                        for (int k = lane; k < C; k += 32) cp_async16_cg_hint(&st.child[0][k], &w.rec[st.ci[0][k] & Rm], pol);
                    } else {
#pragma unroll 1
                        for (int k = lane; k < C; k += 32) {
#if C6_CHILD_CG
                            cp_async16_cg(&st.child[0][k], &w.rec[st.ci[0][k] & Rm]);
#else
                            cp_async16(&st.child[0][k], &w.rec[st.ci[0][k] & Rm]);
#endif
                        }
                    }
                } else {
                    // visit sum (:265)
#pragma unroll 1
                    for (int k = lane; k < C; k += 32) s_local += (int)st.child[0][k].x;
                    s_local = __reduce_add_sync(FULL, s_local);  // one redux.sync instead of a five-step butterfly
                }
            }
            const float visit_sum = (float)s_local;
            uint32_t bb = NIL;
            if (s_local == 0) {
                // No child has been visited: q = 0 and u = prob * sqrt(0 - 0) / 1 = +-0 for every child, a
                // C-way exact tie whatever the priors and the noise are — the reference draws its Dirichlet sample here
                // too and then picks uniformly (:291-295).  One tie draw instead of C Gamma draws and the PUCT pass.
                bb = tie_philox ? tie_pick(cs.sim_index, (uint32_t)depth, D.seed, cs.game_id, (uint32_t)C) : 0u;
            } else {
                // Pass B: the ~5% of draws that need the sampler's other branch or another trial are listed and finished
                // together, one per lane (one diverged excursion per step); pass C turns the draws into softmax weights
                // 2^(L - Lmax) and sums them in a fixed order: per lane the children k = lane, lane + 32, ..., then a butterfly.
                float inv = 0.f;
                if (NOISE) {
                    if (NOISE == 1 && __any_sync(FULL, pend != 0)) {
                        // list the pending ranks (any order): exclusive prefix sum of the per-lane counts
                        const int mycnt = __popc(pend);
                        int incl = mycnt;
#pragma unroll
                        for (int off = 1; off < 32; off <<= 1) {
                            const int o = __shfl_up_sync(FULL, incl, off);
                            if (lane >= off) incl += o;
                        }
                        const int np = __shfl_sync(FULL, incl, 31);
                        int at = incl - mycnt;
                        while (pend) {
                            const int i = __ffs((int)pend) - 1;
                            pend &= pend - 1;
                            st.cell[at++] = (uint8_t)(32 * i + lane);
                        }
                        __syncwarp();
#pragma unroll 1
                        for (int i = lane; i < np; i += 32) {
                            const uint32_t k = st.cell[i];
                            const float l = gamma_log2_full(mt, D.gp.b, D.gp.inv_a, D.gp.c1, D.gp.am1, cs.sim_index, (uint32_t)depth, k,
                                                            D.seed ^ DOMAIN_NOISE, cs.game_id);
                            st.lw[0][k] = l;
                            lmax = fmaxf(lmax, l);
                        }
                    }
                    lmax = key_float(__reduce_max_sync(FULL, float_key(lmax)));
                    __syncwarp();  // pass B wrote other lanes' slots
                    float part = 0.f;
C6_UNROLL(C6_W_UNROLL)
                    for (int k = lane; k < C; k += 32) {
                        const float wk = ex2_det(st.lw[0][k] - lmax);
                        st.lw[0][k] = wk;
                        part = part + wk;
                    }
#pragma unroll
                    for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(FULL, part, off);
                    inv = 1.0f / part;  // the largest draw weighs exactly 1: part >= 1
                }
                // arithmetic phase 2: PUCT (:266-280); per lane the best score, its lowest rank and how many of the lane's
                // children share it
                float bs = 0.f;
                uint32_t cnt = 0;
                const float sqrt_s = sqrtf(visit_sum);  // visit_sum >= 1 here
#if C6_SQRT_TAB
                const float* const sqt = mt.sqrt_tab;  // sqrtf(i), i < 1024: S - n is a small integer
#endif
C6_UNROLL(C6_PUCT_UNROLL)
                for (int k = lane; k - lane < C; k += 32) {  // every lane runs every round (warp vote inside)
                    const bool in = k < C;
                    const uint4 ch = st.child[0][in ? k : 0];
                    const int cn = in ? (int)ch.x : 0;
                    const float cp = __uint_as_float(ch.z);
                    float nz = 0.f;
                    if (NOISE) nz = st.lw[0][in ? k : 0] * inv;
                    const float prob = eps * nz + one_m_eps * cp;  // :269
                    float sc;
                    if (__all_sync(FULL, cn == 0)) {
                        // a round of unvisited children (the rule below the root): q = 0 (:77-79), sqrt(S - 0) = sqrt(S) for all of
                        // them and the division by 1 + 0 is the identity — the same bits without the two divisions and the root
                        sc = (turn < 0 ? -0.f : 0.f) + cpuct * (prob * sqrt_s);
                    } else {
                        const float cw = __uint_as_float(ch.y);
                        const float visit = (float)cn;
                        // q_value (:77-79) and u (:270).  The IEEE division's fast path is left for a subroutine whenever an
                        // operand is zero — and an unvisited child has W = 0 — so both divisions are fed stand-in operands in
                        // those lanes and the exact result (0, or the zero numerator itself) is selected afterwards: same
                        // values, no calls (r01k capture: 5 % of all instructions)
                        const bool seen = cn > 0 && cw != 0.f;
                        const float qd = (seen ? cw : 1.f) / (seen ? visit : 1.f);
                        const float q = cn > 0 ? (cw != 0.f ? qd : cw) : 0.f;  // +-0 / visit = +-0
#if C6_SQRT_TAB
                        const int sai = s_local - cn;  // >= 0
                        const float un = prob * (sai < 1024 ? sqt[sai] : sqrtf((float)sai));  // sqt[0] = +0 = S - n itself
#else
                        const float sa = visit_sum - visit;  // >= 0; sqrtf(0) would also leave the fast path
                        const float un = prob * (sa != 0.f ? sqrtf(sa != 0.f ? sa : 1.f) : sa);
#endif
                        const float ud = (un != 0.f ? un : 1.f) / (1.f + visit);
                        const float u = un != 0.f ? ud : un;  // :270; +-0 / (1 + visit) = +-0
                        sc = (turn < 0 ? -q : q) + cpuct * u;  // :272-280
                    }
                    if (in) {
                        st.lw[0][k] = sc;
                        if (bb == NIL || sc > bs) bs = sc, bb = (uint32_t)k, cnt = 1;
                        else if (sc == bs) cnt += 1;
                    }
                }
                // maximum over the warp: score, lowest rank holding it, number of children holding it — three integer warp
                // reductions (scores through order-preserving keys; a lane without children has key 0, below every score)
                {
                    const uint32_t my_key = bb == NIL ? 0u : float_key(bs);
                    const uint32_t top = __reduce_max_sync(FULL, my_key);
                    const bool mine = bb != NIL && my_key == top;
                    cnt = __reduce_add_sync(FULL, mine ? cnt : 0u);
                    bb = __reduce_min_sync(FULL, mine ? bb : NIL);
                    bs = key_float(top);
                }
                // exact ties: uniformly among the tied children (:291-295) or, for hand-checkable traces, the lowest rank
                if (tie_philox && cnt > 1) {
                    uint32_t j = tie_pick(cs.sim_index, (uint32_t)depth, D.seed, cs.game_id, cnt);
                    __syncwarp();
#pragma unroll 1
                    for (int i = 0; 32 * i < C; ++i) {
                        const int k = 32 * i + lane;
                        const unsigned bal = __ballot_sync(FULL, k < C && st.lw[0][k] == bs);
                        const uint32_t c = (uint32_t)__popc(bal);
                        if (j < c) {
                            bb = 32u * i + nth_set_bit(bal, j);
                            break;
                        }
                        j -= c;
                    }
                }
            }
            __syncwarp();  // cp.async results and st.empty of the other lanes
            child = st.ci[0][bb];
            {  // the chosen child's record is already staged: no second trip to memory for the next step
                const uint4 ch = st.child[0][bb];
                r.n = (int32_t)ch.x, r.w = __uint_as_float(ch.y), r.p = __uint_as_float(ch.z), r.link = ch.w;
            }
            chosen_bit = kth_empty(st, bb, lt, lane);
        }
        // Simulate::simulate_in (LC/policy/simulate/mod.rs:262-274)
        __syncwarp();  // every lane is done with the stage of this step
        if (lane == 0) {
            uint32_t* col = turn < 0 ? st.B : st.W;
            col[chosen_bit >> 5] |= 1u << (chosen_bit & 31);
        }
        L.key ^= ldz(&w.Z[(turn > 0 ? 256 : 0) + chosen_bit]);
        L.stones += 1;
        L.last_bit = chosen_bit;
        cur = child & Rm;
        depth += 1;
    }
    L.rec = cur;
    L.depth = depth;
    L.leaf_w = r.w, L.leaf_p = r.p;
    if (lane == 0) st.meta.sum_depth += (uint32_t)depth, st.meta.sims += 1;
    __syncwarp();
}

// update(): only the leaf's parent takes the value, every ancestor (leaf excluded) gets visit += 1.
// MODE 0: both halves (the reference's schedule); 1: the visits only, 2: the value only (multi-leaf mode: the visits are
// applied when the descent ends, the value when the leaf has been evaluated).
template <bool FUSED, int NOISE, int MODE = 0>
__device__ __forceinline__ void backup(const Dev& D, Warp& w, WarpStage<NOISE>& st, int depth, float v) {
    __syncwarp();
    for (int i = MODE == 2 ? depth - 1 + w.lane : w.lane; i < depth; i += 32) {
        if (i < 0) continue;
        uint32_t a;
        Rec r;
        if (FUSED && i < PATH_CACHE) {  // as read by this simulation's descent; nothing has written them since
            a = st.pidx[i], r = rec_of(st.prec[i]);
        } else {
            a = w.path(D)[i], r = load_rec(&w.rec[a]);
        }
        if (MODE != 2) r.n += 1;
        if (MODE != 1 && i == depth - 1) r.w = r.w + v;
        store_rec(&w.rec[a], r);
        if (FUSED && i == 0) st.rootrec = bits_of(r);
    }
    __syncwarp();
}

// full-board leaf (:338-348): no evaluator call, no visit increment, value = winner.  Reached a handful
// of times per game at most: the win check is kept out of line and out of the hot instruction stream.
__device__ __noinline__ float terminal_value(const uint32_t* sb, const uint32_t* sw) {
    uint32_t B[8], W[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) B[j] = sb[j], W[j] = sw[j];
    return (float)win_check(B, W);
}
template <bool FUSED, int NOISE>
__device__ __forceinline__ void terminal_leaf(const Dev& D, Warp& w, WarpStage<NOISE>& st, const Leaf& L) {
    __syncwarp();
    const float v = terminal_value(st.B, st.W);
    backup<FUSED>(D, w, st, L.depth, v);
}

// ------------------------------------------------------------------------------------------------
// leaf evaluators: value() for the warp, prior(bit) for the cell at bit position `bit` (already masked:
// only asked for empty cells).  Twins of the oracle's evaluators.
// ------------------------------------------------------------------------------------------------
// RandomEvaluator (:137-169) on Philox, fused with get_from (:218-234) and recover_way8
// (augment.rs:75-92): draw s of original cell (r,c) is what plane s shows at its image of (r,c), so the
// recovered prior needs exactly the 8 draws of its own cell.
struct EvPhilox {
    uint32_t k0, k1, sim;
    int lane;
    __device__ __forceinline__ void init(const Dev& D, const Cursor& cs, int lane_) { init(D, cs.game_id, cs.sim_index, lane_); }
    __device__ __forceinline__ void init(const Dev& D, uint32_t game_id, uint32_t sim_index, int lane_) {
        k0 = D.seed ^ DOMAIN_EVAL, k1 = game_id, sim = sim_index, lane = lane_;
    }
    // ((((((((0+v0)+v1)+..)+v7)/8 (:224): v0..3 are Philox block (sim, 225, 0, 0), v4..7 block (sim, 225, 1, 0).  Even lanes compute
    // the first block, odd lanes the second (one block's worth of instructions for the warp instead of two); lane 1's four draws
    // are then added in order.  Call with the whole warp.
    __device__ __forceinline__ float value() const {
        const Philox4 a = philox4x32_10(sim, CELLS, (uint32_t)(lane & 1), 0, k0, k1);
        const float x = u32_to_pm1(a.x), y = u32_to_pm1(a.y), z = u32_to_pm1(a.z), t = u32_to_pm1(a.w);
        float sum = (((0.f + x) + y) + z) + t;  // lane 0: the first half
        sum = __shfl_sync(FULL, sum, 0);
        sum = sum + __shfl_sync(FULL, x, 1), sum = sum + __shfl_sync(FULL, y, 1);
        sum = sum + __shfl_sync(FULL, z, 1), sum = sum + __shfl_sync(FULL, t, 1);
        return sum / 8.f;
    }
    // 24-bit draw number i (0..7) of the 192 bits in words q[0..5]: bits [24 i, 24 i + 24); the draw is x 2^-23 - 1 in [-1, 1)
    static __device__ __forceinline__ uint32_t bits24(const uint32_t (&q)[6], int i) {
        const int w = (3 * i) >> 2, sh = 8 * ((3 * i) & 3);
        const uint32_t x = sh == 0 ? q[w] : (sh == 8 ? q[w] >> 8 : __funnelshift_r(q[w], q[w + 1 < 6 ? w + 1 : 5], sh));
        return x & 0xFFFFFFu;
    }
    // recover_way8 order: t_i = p[2i] + flip(p[2i+1]), ((((0+t0)+t1)+t2)+t3)/8.  A pair sum t_i = (x 2^-23 - 1) + (y 2^-23 - 1) is a
    // multiple of 2^-23 below 2 in magnitude — exact in f32 — so it is formed in integers, (x + y - 2^24) 2^-23: the same bits
    // (the oracle adds the two f32 draws) with one conversion per pair instead of two.
    static __device__ __forceinline__ float fold8(const uint32_t (&q)[6]) {
        float total = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            total = total + (float)((int)(bits24(q, 2 * i) + bits24(q, 2 * i + 1)) - (1 << 24)) * (1.0f / 8388608.0f);
        return total / 8.f;
    }
    // Priors of the leaf's children by rank.  The children of ranks k and k + 32 (k = 64 j + lane) share THREE Philox blocks
    // (sim, 32 j + lane, 0..2, 0): their 384 bits are sixteen 24-bit draws, eight per child (a cell's eight draws are what the
    // eight symmetric planes show at its images: twin PhiloxRandomEvaluator in oracle/c6_oracle.hpp) — 1.5 blocks per child
    // instead of 2, nothing of a block unused.
    template <class Stage>
    __device__ __forceinline__ void fill_priors(Stage& st, int C) const {
#pragma unroll 1
        for (int k = lane; k < C; k += 64) {
            const uint32_t kb = (uint32_t)(((k >> 6) << 5) | lane);
            const Philox4 a = philox4x32_10(sim, kb, 0u, 0, k0, k1), b = philox4x32_10(sim, kb, 1u, 0, k0, k1);
            const uint32_t q0[6] = {a.x, a.y, a.z, a.w, b.x, b.y};
            st.prior[0][k] = fold8(q0);
            if (k + 32 < C) {
                const Philox4 c = philox4x32_10(sim, kb, 2u, 0, k0, k1);
                const uint32_t q1[6] = {b.z, b.w, c.x, c.y, c.z, c.w};
                st.prior[0][k + 32] = fold8(q1);
            }
        }
    }
};

// DetEval / NegEval: hash of the oriented board, evaluated on each of the 8 planes.
template <bool NEG>
struct EvHash {
    uint64_t h[8];
    int lane;
    __device__ __forceinline__ void init(const Leaf& L, const uint32_t* sB, const uint32_t* sW, const uint64_t* Z, int lane_) {
        lane = lane_;
        const int turn = turn_of(L.stones);
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            uint64_t k = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t bit = 32u * j + lane;
                int r = (int)(bit >> 4), c = (int)(bit & 15), rr, cc;
                sym_fwd(s, r, c, rr, cc);
                if ((sB[j] >> lane) & 1u) k ^= Z[rr * 16 + cc];
                if ((sW[j] >> lane) & 1u) k ^= Z[256 + rr * 16 + cc];
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) k ^= __shfl_xor_sync(FULL, k, off);
            h[s] = splitmix64(k ^ (turn < 0 ? 0xB1ACull : 0x3417Eull));
        }
    }
    __device__ __forceinline__ float value() const {
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) sum = sum + ((float)(uint32_t)(h[s] >> 40) * (1.0f / 8388608.0f) - 1.0f);
        return sum / 8.f;
    }
    template <class Stage>
    __device__ __forceinline__ void fill_priors(Stage& st, int C) const {
#pragma unroll 1
        for (int k = lane; k < C; k += 32) st.prior[0][k] = prior((uint32_t)st.cell[k]);
    }
    __device__ __forceinline__ float prior(uint32_t bit) const {
        int r = (int)(bit >> 4), c = (int)(bit & 15);
        float pv[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int rr, cc;
            sym_fwd(s, r, c, rr, cc);
            uint32_t x = (uint32_t)(splitmix64(h[s] + (uint64_t)(rr * 15 + cc) + 1) >> 40);
            pv[s] = NEG ? ((float)x * (1.0f / 8388608.0f) - 1.0f) : ((float)x * (1.0f / 16777216.0f));
        }
        float total = 0.f;
        total = total + (pv[0] + pv[1]), total = total + (pv[2] + pv[3]);
        total = total + (pv[4] + pv[5]), total = total + (pv[6] + pv[7]);
        return total / 8.f;
    }
};

// get_from() on an evaluator batch in the API layout (values[8], policies[8][225]) for this leaf.
struct EvBatch {
    const float* values;
    const float* pol;
    int lane;
    __device__ __forceinline__ float value() const {
        float myv = lane < 8 ? values[lane] : 0.f;
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) sum = sum + __shfl_sync(FULL, myv, s);  // iter().sum::<f32>() order (:224)
        return sum / 8.f;
    }
    template <class Stage>
    __device__ __forceinline__ void fill_priors(Stage& st, int C) const {
#pragma unroll 1
        for (int k = lane; k < C; k += 32) st.prior[0][k] = prior((uint32_t)st.cell[k]);
    }
    __device__ __forceinline__ float prior(uint32_t bit) const {
        int r = (int)(bit >> 4), c = (int)(bit & 15);
        float pv[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int rr, cc;
            sym_fwd(s, r, c, rr, cc);
            pv[s] = pol[s * CELLS + rr * 15 + cc];
        }
        float total = 0.f;  // recover_way8 order: ((((0+t0)+t1)+t2)+t3)/8, t_i = p[2i] + flip(p[2i+1])
        total = total + (pv[0] + pv[1]), total = total + (pv[2] + pv[3]);
        total = total + (pv[4] + pv[5]), total = total + (pv[6] + pv[7]);
        return (total / 8.f) * 1.0f;  // * (cell empty) mask (:227-232)
    }
};

// ------------------------------------------------------------------------------------------------
// expand (:350-377) for a non-full leaf, then update()
// ------------------------------------------------------------------------------------------------
// How a child of the leaf B finds the record of its board if another parent made it first
// (`entry(hash).or_insert`, :367-370; oracle/pool_model.cpp checks these rules against the map-based
// restatement on the CPU).  Write B = root + a_1 .. a_d, m = colour of the stone B's children add, X = B + c.
// Every expanded node kept by the collector contains the root, so another expanded parent of X is
// E = B - s' + c for a path stone s' of colour m:
//   (F)  s' = a_d (same colour as c): E = P + c is the sibling of B under its parent P, whose record the last
//        select step left in st.child; if it is expanded, X is entry rank_E(a_d) of its link run.  The common case.
//   (S1) s' an earlier path stone of colour m (possible from depth 3): E is looked up in the key table of
//        expanded nodes; X is entry rank_E(s') of its link run.
//   (S2) X itself expanded with no expanded parent left after a collection (flagged per stone count in
//        GameState::orph): looked up in the key table.
// An unexpanded node that nothing links any more carries no state (N = 0, W = 0, its prior is overwritten by
// the next expanding parent, :371), so a fresh record is equivalent.
//
// S1 + S2, out of line: rare (late-game deep paths), kept out of the hot instruction stream.  `cm` = the
// leaf's stones of colour m (a board in the warp's stage), `skip` = bit of a_d when rule F covers it (else NIL), `cell[k]` = bit position of the
// leaf's empty cell of rank k.  `mine` / `res`: bit i = this lane's child k = 32 i + lane exists / already has a
// record.  Returns the updated `res`; record indices go to pos[k].
__device__ __noinline__ unsigned resolve_rare(const Slot* idx, uint32_t Im, uint32_t epoch, const uint32_t* lnk, uint32_t Lm,
                                              const Rec* rec, uint32_t Rm, const uint64_t* Z, const GameState* gs, uint32_t* pos,
                                              const uint32_t* empty, const uint8_t* cell, const uint32_t* cm, uint64_t key,
                                              int stones, uint32_t zoff, uint32_t skip, unsigned mine, unsigned res, int lane,
                                              int depth) {
    if (depth >= 3) {
        const uint32_t* rootcol = zoff ? gs->W : gs->B;
        int before = 0;  // empty cells of the leaf in the words below j
        for (int j = 0; j < 8; ++j) {
            uint32_t others = cm[j] & ~rootcol[j];
            if ((skip >> 5) == (uint32_t)j) others &= ~(1u << (skip & 31));
            while (others) {  // warp-uniform
                const int b = __ffs((int)others) - 1;
                others &= others - 1;
                const uint32_t sp = 32u * j + (uint32_t)b;
                const uint32_t rsp = (uint32_t)(before + __popc(empty[j] & ((1u << b) - 1u)));
                const uint64_t ks = key ^ Z[zoff + sp];
                for (int i = 0; i < 8; ++i)
                    if (((mine & ~res) >> i) & 1u) {
                        const uint32_t cbit = cell[32 * i + lane];
                        const uint32_t e = xh_find(idx, Im, epoch, ks ^ Z[zoff + cbit]);
                        if (e != NIL) {
                            const uint32_t link = rec[e & Rm].link;
                            pos[i * 32 + lane] = lnk[(link + rsp - (cbit < sp ? 1u : 0u)) & Lm];
                            res |= 1u << i;
                        }
                    }
            }
            before += __popc(empty[j]);
        }
    }
    const uint32_t s1 = (uint32_t)stones + 1u;
    if ((gs->orph[s1 >> 5] >> (s1 & 31)) & 1u) {
        for (int i = 0; i < 8; ++i)
            if (((mine & ~res) >> i) & 1u) {
                const uint32_t x = xh_find(idx, Im, epoch, key ^ Z[zoff + cell[32 * i + lane]]);
                if (x != NIL) pos[i * 32 + lane] = x, res |= 1u << i;
            }
    }
    return res;
}

// STAGED: st.child still holds the children of the leaf's parent from the last select step (fused kernel);
// otherwise (the expansion runs in a later launch than the descent) they are staged again here.
// VMODE: backup mode once the leaf is expanded (0 = visits and value, 2 = value only: see backup)
template <bool STAGED, int VMODE = 0, class Ev, int NOISE>
__device__ __forceinline__ void expand_leaf(const Dev& D, Warp& w, Cursor& cs, const Leaf& L, const Ev& ev, float v,
                                            WarpStage<NOISE>& st) {
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1, Im = D.I - 1;
    const uint32_t lt = (1u << lane) - 1u;
    const int C = CELLS - L.stones;
    const int turn = turn_of(L.stones);
    const uint32_t zoff = turn > 0 ? 256u : 0u;
    const uint32_t link_base = (cs.lnk_head + 3u) & ~3u;  // link runs start on 16-byte boundaries (four links per cp.async)
    const uint32_t ec = cs.exp_count;
    const uint32_t rec_tail = st.meta.rec_tail, lnk_tail = st.meta.lnk_tail, epoch = st.meta.epoch;
    if (link_base - lnk_tail + (uint32_t)C > D.L) cs.err = ERR_LNK_FULL;
    if (cs.rec_head - rec_tail + (uint32_t)C > D.R) cs.err = ERR_REC_FULL;
    if (ec + 1 > D.E) cs.err = ERR_EXP_FULL;
    if (cs.err) return;
    // the key-table slot the leaf will be inserted at when the expansion is done: its first probe is in flight from here
    const uint32_t slot0 = hash_slot(L.key) & Im;
    const uint32_t slot0_epoch = __ldcg(&w.idx[slot0].epoch);
    // remember the board (st.B / st.W, contiguous) for the collector
    __syncwarp();
    if (lane < 16) w.expB(D)[(size_t)ec * 16 + lane] = lane < 8 ? st.B[lane] : st.W[lane - 8];
    // rule F applies when the last path stone has the colour of the stone the children add
    const bool sib = L.depth >= 1 && turn_of(L.stones - 1) == turn;
    const uint32_t lj = L.last_bit >> 5, lm = 1u << (L.last_bit & 31);
    if (!STAGED && sib) {  // the parent's children (its empties are the leaf's plus the last stone)
        const uint32_t prec = w.path(D)[L.depth - 1];
        stage_children(w, st, w.rec[prec].link, C + 1, lane, Rm, Lm);
        __syncwarp();
    }
    if (lane < 8) st.empty[lane] = ~(st.B[lane] | st.W[lane]) & (lane < 7 ? 0x7FFF7FFFu : 0x00007FFFu);
    __syncwarp();
    const uint64_t key = L.key;
    // Bit position of every child by rank (st.cell), byte-wise: lane l owns bits 8 l .. 8 l + 7 of the 256-bit empty mask
    // (byte l & 3 of word l >> 2); an exclusive prefix sum of the per-byte counts over the warp gives the rank of a lane's first
    // empty cell, and the lane writes its (at most eight) cells.  r0 = rank of the last path stone among the empties of a
    // sibling that lacks it (before the c < a_d correction): the empties below that bit, counted per word by lanes 0..7.
    uint32_t r0;
    {
        const uint32_t ew = st.empty[lane >> 2];
        uint32_t eb = (ew >> (8 * (lane & 3))) & 0xFFu;
        const int c = __popc(eb);
        int incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += o;
        }
        int at = incl - c;
        while (eb) {
            st.cell[at++] = (uint8_t)(8 * lane + (__ffs((int)eb) - 1));
            eb &= eb - 1;
        }
        const uint32_t e = lane < 8 ? st.empty[lane & 7] : 0u;
        const uint32_t below = lane < (int)lj ? e : (lane == (int)lj ? e & (lm - 1u) : 0u);
        r0 = __reduce_add_sync(FULL, (uint32_t)__popc(below));
    }
    __syncwarp();
    // bit i: this lane has child k = 32 i + lane
    const unsigned mine = ((1u << (C >> 5)) - 1u) | (lane < (C & 31) ? 1u << (C >> 5) : 0u);
    // A: rule F — the link of every shared child in flight; priors computed under the latency.  The sibling P + c
    // of the leaf under its parent P has rank k + [c > a_d] there.
    unsigned res = 0;  // bit i: child k already has a record, index in st.pos
    if (sib) {
#pragma unroll 1
        for (int k = lane; k < C; k += 32) {
            const uint32_t cbit = st.cell[k];
            const uint4 sv = st.child[0][k + (cbit > L.last_bit ? 1 : 0)];
            if ((int)sv.x >= 1) {
                cp_async4(&st.pos[0][k], &w.lnk[(sv.w + r0 - (cbit < L.last_bit ? 1u : 0u)) & Lm]);
                res |= 1u << (k >> 5);
            }
        }
    }
    ev.fill_priors(st, C);
    // rules S1 / S2 (warp-uniform test)
    {
        const uint32_t s1 = (uint32_t)L.stones + 1u;
        if (L.depth >= 3 || orphan_flag(*w.gs, st.meta, s1)) {
            res = resolve_rare(w.idx, Im, epoch, w.lnk, Lm, w.rec, Rm, D.Z, w.gs, &st.pos[0][0], st.empty, st.cell, turn < 0 ? st.B : st.W, key,
                               L.stones, zoff, sib ? L.last_bit : NIL, mine, res, lane, L.depth);
        }
    }
    cp_async_wait_all();
    // B: new records are laid out contiguously in row-major child order
    const unsigned isnew = mine & ~res;
    uint32_t base = cs.rec_head;
#pragma unroll 1
    for (int i = 0; 32 * i < C; ++i) {
        const bool nw = (isnew >> i) & 1u;
        const unsigned bal = __ballot_sync(FULL, nw);
        if (nw) st.pos[i][lane] = base + __popc(bal & lt);
        base += __popc(bal);
    }
    // C: records, priors, link run
#pragma unroll 1
    for (int k = lane; k < C; k += 32) {
        const uint32_t child = st.pos[0][k] & Rm;
        const float pr = st.prior[0][k];
        if ((isnew >> (k >> 5)) & 1u) {
            Rec nr;
            nr.n = 0, nr.w = 0.f, nr.p = pr, nr.link = NIL;
            store_rec(&w.rec[child], nr);
        } else {
            w.rec[child].p = pr;  // n_prob is overwritten by the last expanding parent (:371)
        }
        w.lnk[(link_base + (uint32_t)k) & Lm] = child;
    }
    __syncwarp();
    // the leaf itself: visit += 1, children attached (:353-357,374-377), findable by key from now on
    if (lane == 0) {
        Rec lr;
        if (STAGED) {  // the descent of this very simulation read the leaf's record: N = 0, W and P as in L
            lr.n = 0, lr.w = L.leaf_w, lr.p = L.leaf_p;
        } else {
            lr = load_rec(&w.rec[L.rec]);
        }
        lr.n += 1;
        lr.link = link_base & Lm;
        store_rec(&w.rec[L.rec], lr);
        if (STAGED && L.depth == 0) st.rootrec = bits_of(lr);  // the root itself was the leaf
        w.exp_rec(D)[ec] = L.rec;
        w.exp_key(D)[ec] = key;
        uint32_t pos = slot0;  // xh_insert with the first probe already answered (nothing else inserts in between)
        if (slot0_epoch == epoch) {
            pos = (pos + 1) & Im;
            while (__ldcg(&w.idx[pos].epoch) == epoch) pos = (pos + 1) & Im;
        }
        __stcg(reinterpret_cast<uint4*>(&w.idx[pos]), make_uint4((uint32_t)key, (uint32_t)(key >> 32), L.rec, epoch));
    }
    cs.exp_count = ec + 1;
    cs.lnk_head = link_base + (uint32_t)C;
    cs.rec_head = base;
    if (lane == 0) {
        WarpMeta& m = st.meta;
        m.sum_cleaf += (uint32_t)C, m.evals += 1;
        m.peak_rec = max(m.peak_rec, base - rec_tail);
        m.peak_lnk = max(m.peak_lnk, cs.lnk_head - lnk_tail);
    }
    backup<STAGED, NOISE, VMODE>(D, w, st, L.depth, v);
}

// ------------------------------------------------------------------------------------------------
// Waves: the schedule shared by the callback paths (K = 1: the reference's; the leaf waits for the evaluator between two
// launches) and the opt-in multi-leaf mode (K = Config::virtual_loss > 1; oracle twin: AlphaZero::search_wave).
//   front half: up to K descents.  Each applies the visits of update() (:400-406) at once — for K > 1 they are the virtual
//     loss the next descent of the wave sees — and parks its leaf in GameState::pend.  A descent that ends in a leaf already
//     waiting in this wave is discarded and ends the wave; a full-board leaf has its value at once (:338-348).
//   back half: the waiting leaves are expanded in descent order, each with the streams of its own simulation index, and each
//     adds its value to its parent (:394-398).
// ------------------------------------------------------------------------------------------------
template <int NOISE>
__device__ __forceinline__ void wave_descents(const Dev& D, Warp& w, Cursor& cs, WarpStage<NOISE>& st, const MathTab& mt, int& sims_this_move) {
    GameState& gs = *w.gs;
    const int lane = w.lane;
    uint32_t* const path0 = w.pth;
    int n_pend = 0;
#pragma unroll 1
    for (int j = 0; j < D.K && sims_this_move < D.num_sim; ++j) {
        w.pth = path0 + n_pend * 232;
        Leaf L;
        descend<NOISE, false>(D, w, cs, L, st, mt);
        const uint32_t index = cs.sim_index;
        cs.sim_index += 1;
        const bool terminal = L.stones == CELLS;
        bool dup = false;
        if (!terminal)
            for (int i = 0; i < n_pend; ++i) dup = dup || gs.pend[i].rec == L.rec;
        if (dup) {  // the discarded descent changed nothing; it is not a simulation
            if (lane == 0) st.meta.sims -= 1;
            break;
        }
        backup<false, NOISE, 1>(D, w, st, L.depth, 0.f);
        sims_this_move += 1;
        if (terminal) {
            const float v = terminal_value(st.B, st.W);
            backup<false, NOISE, 2>(D, w, st, L.depth, v);
            continue;
        }
        PendingLeaf& p = gs.pend[n_pend];
        if (lane < 8) p.B[lane] = st.B[lane], p.W[lane] = st.W[lane];
        if (lane == 0) p.key = L.key, p.rec = L.rec, p.depth = L.depth, p.stones = L.stones, p.last_bit = L.last_bit, p.sim_index = index, p.batch_pos = 0;
        __syncwarp();
        n_pend += 1;
    }
    w.pth = path0;
    if (lane == 0) gs.n_pend = n_pend;
    __syncwarp();
}

// make_ev(p, L): the evaluator of waiting leaf p (its board is in st.B / st.W when it is called)
template <int NOISE, class MakeEv>
__device__ __forceinline__ void wave_expansions(const Dev& D, Warp& w, Cursor& cs, WarpStage<NOISE>& st, MakeEv make_ev) {
    GameState& gs = *w.gs;
    const int lane = w.lane;
    uint32_t* const path0 = w.pth;
    const int n = gs.n_pend;
#pragma unroll 1
    for (int i = 0; i < n && !cs.err; ++i) {
        const PendingLeaf& p = gs.pend[i];
        __syncwarp();
        if (lane < 8) st.B[lane] = p.B[lane], st.W[lane] = p.W[lane];
        Leaf L;
        L.key = p.key, L.stones = p.stones, L.rec = p.rec, L.depth = p.depth, L.last_bit = p.last_bit, L.leaf_w = 0.f, L.leaf_p = 0.f;
        __syncwarp();
        w.pth = path0 + i * 232;
        const auto ev = make_ev(p, L);
        expand_leaf<false, 2>(D, w, cs, L, ev, ev.value(), st);
    }
    w.pth = path0;
    __syncwarp();
    if (lane == 0) gs.n_pend = 0;
    __syncwarp();
}

}  // namespace c6
