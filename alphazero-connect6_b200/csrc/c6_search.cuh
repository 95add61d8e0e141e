// The per-simulation hot path (one warp per game): select -> leaf evaluation -> expand -> backup.
//   select / maximum_from   LC/policy/alphazero_policy/mod.rs:243-323  (looped as in search() :417-420)
//   expand / get_from       LC/policy/alphazero_policy/mod.rs:218-234,331-381, augment.rs:60-92
//   update                  LC/policy/alphazero_policy/mod.rs:386-408
// Shaped by the ncu profile of the first version (profiles/r01_*): latency-bound on HBM round trips,
// so (1) everything a move needs from the game state lives in registers (`Cursor`) instead of being
// re-read and re-written in global memory every simulation, (2) the board-key inserts of an expansion
// are issued back to back and only their returned epochs are inspected, (3) expansion runs in two
// half-batches so that the live state fits the register budget of 5 blocks/SM without spills, and the
// priors of a batch are computed while its index probes are in flight.
#pragma once
#include "c6_device.cuh"

namespace c6 {

// Per-move copy of the mutable game bookkeeping, identical in every lane of the warp.
struct Cursor {
    uint32_t rec_head, lnk_head, exp_count, sim_index;
    uint32_t rec_tail, lnk_tail, epoch, game_id;
    uint32_t sum_ck, sum_cleaf, sum_depth, evals, sims;
    uint32_t peak_rec, peak_lnk;
    int err;
};

__device__ __forceinline__ void load_cursor(const GameState& g, Cursor& c) {
    c.rec_head = g.rec_head, c.lnk_head = g.lnk_head, c.exp_count = g.exp_count, c.sim_index = g.sim_index;
    c.rec_tail = g.rec_tail, c.lnk_tail = g.lnk_tail, c.epoch = g.epoch, c.game_id = g.game_id;
    c.sum_ck = 0, c.sum_cleaf = 0, c.sum_depth = 0, c.evals = 0, c.sims = 0;
    c.peak_rec = g.peak_rec, c.peak_lnk = g.peak_lnk;
    c.err = 0;
}
__device__ __forceinline__ void store_cursor(const Dev& D, GameState& g, const Cursor& c, int lane) {
    if (lane == 0) {
        g.rec_head = c.rec_head, g.lnk_head = c.lnk_head, g.exp_count = c.exp_count, g.sim_index = c.sim_index;
        g.sum_ck += c.sum_ck, g.sum_cleaf += c.sum_cleaf, g.sum_depth += c.sum_depth, g.evals += c.evals, g.sims += c.sims;
        g.peak_rec = c.peak_rec, g.peak_lnk = c.peak_lnk;
        if (c.err) {
            g.status = ST_ERROR, g.err = c.err;
            atomicAdd(D.gcount + 2, 1ull);
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// select
// ------------------------------------------------------------------------------------------------
// `gl` = per-warp shared scratch for the Gamma draws of one select step ([8][32] doubles), NOISE only.
template <bool NOISE>
__device__ __forceinline__ void descend(const Dev& D, Warp& w, Cursor& cs, Leaf& L, double* gl) {
    const GameState& g = *w.gs;
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) L.B[j] = g.B[j], L.W[j] = g.W[j];
    L.key = g.key;
    L.stones = g.stones;
    uint32_t cur = g.root;
    int depth = 0;
    const float eps = D.eps, one_m_eps = 1.f - D.eps, cpuct = D.cpuct;
    for (;;) {
        const Rec r = load_rec(&w.rec[cur]);
        if (r.n == 0) break;  // no children: leaf (:251-253)
        if (lane == 0) w.path[depth] = cur;
        const int C = CELLS - L.stones;
        const int turn = turn_of(L.stones);
        cs.sum_ck += (uint32_t)C;
        uint32_t chosen_bit = 0, child = 0;
        uint32_t e[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = ~(L.B[j] | L.W[j]) & valid_mask(j);
        if (C == 1) {  // single child: taken without scoring (:254-257)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (e[j]) chosen_bit = 32 * j + (__ffs((int)e[j]) - 1);
            child = w.lnk[r.link & Lm];
        } else {
            uint32_t ci[8];
            int cn[8];
            float cw[8], cp[8];
            int pre = 0, s_local = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ci[j] = NIL;
                if ((e[j] >> lane) & 1u) {
                    int k = pre + __popc(e[j] & ((1u << lane) - 1u));
                    ci[j] = w.lnk[(r.link + (uint32_t)k) & Lm];
                }
                pre += __popc(e[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                cn[j] = 0, cw[j] = 0.f, cp[j] = 0.f;
                if (ci[j] != NIL) {
                    Rec ch = load_rec(&w.rec[ci[j] & Rm]);
                    cn[j] = ch.n, cw[j] = ch.w, cp[j] = ch.p;
                    s_local += ch.n;
                }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s_local += __shfl_xor_sync(FULL, s_local, off);
            const float visit_sum = (float)s_local;  // :265
            double inv = 0.0;
            bool flat = false;
            if (NOISE) {
                // Dirichlet(alpha; C) drawn at every node with C >= 2 (:262,283): Gamma draws keyed
                // (sim, depth, child k), summed lane-partial then butterfly (fixed order), scaled.
                double part = 0.0;
                int pre2 = 0;
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const uint32_t ej = ~(L.B[j] | L.W[j]) & valid_mask(j);
                    double gj = 0.0;
                    if ((ej >> lane) & 1u) {
                        int k = pre2 + __popc(ej & ((1u << lane) - 1u));
                        gj = gamma_draw(cs.sim_index, (uint32_t)depth, (uint32_t)k, D.seed ^ DOMAIN_NOISE, cs.game_id, D.alpha);
                        part = part + gj;
                    }
                    gl[j * 32 + lane] = gj;
                    pre2 += __popc(ej);
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(FULL, part, off);
                flat = !(part > 0.0);
                inv = flat ? 0.0 : 1.0 / part;
            }
            float bs = 0.f;
            uint32_t bb = NIL, bc = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (ci[j] != NIL) {
                    float nz = 0.f;
                    if (NOISE) nz = flat ? (float)(1.0 / (double)C) : (float)(gl[j * 32 + lane] * inv);
                    float visit = (float)cn[j];
                    float q = cn[j] > 0 ? cw[j] / visit : 0.f;                  // q_value (:77-79)
                    float prob = eps * nz + one_m_eps * cp[j];                  // :269
                    float u = prob * sqrtf(visit_sum - visit) / (1.f + visit);  // :270
                    float sc = (turn < 0 ? -q : q) + cpuct * u;                 // :272-280
                    if (bb == NIL || sc > bs) bs = sc, bb = 32u * j + lane, bc = ci[j];
                }
            }
            // maximum; exact ties -> lowest row-major index (the reference picks uniformly, :291-295)
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                float os = __shfl_xor_sync(FULL, bs, off);
                uint32_t ob = __shfl_xor_sync(FULL, bb, off), oc = __shfl_xor_sync(FULL, bc, off);
                bool take = ob != NIL && (bb == NIL || os > bs || (os == bs && ob < bb));
                if (take) bs = os, bb = ob, bc = oc;
            }
            chosen_bit = bb, child = bc;
        }
        // Simulate::simulate_in (LC/policy/simulate/mod.rs:262-274)
        const uint32_t wj = chosen_bit >> 5, m = 1u << (chosen_bit & 31);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t mm = (j == (int)wj) ? m : 0u;
            if (turn < 0) L.B[j] |= mm; else L.W[j] |= mm;
        }
        L.key ^= w.Z[(turn > 0 ? 256 : 0) + chosen_bit];
        L.stones += 1;
        cur = child & Rm;
        depth += 1;
    }
    L.rec = cur;
    L.depth = depth;
    cs.sum_depth += (uint32_t)depth;
    cs.sims += 1;
    __syncwarp();
}

// update(): only the leaf's parent takes the value, every ancestor (leaf excluded) gets visit += 1.
__device__ __forceinline__ void backup(Warp& w, int depth, float v) {
    __syncwarp();
    for (int i = w.lane; i < depth; i += 32) {
        uint32_t a = w.path[i];
        Rec r = load_rec(&w.rec[a]);
        r.n += 1;
        if (i == depth - 1) r.w = r.w + v;
        store_rec(&w.rec[a], r);
    }
    __syncwarp();
}

// full-board leaf (:338-348): no evaluator call, no visit increment
__device__ __forceinline__ void terminal_leaf(Warp& w, const Leaf& L) {
    float v = (float)win_check(L.B, L.W);
    backup(w, L.depth, v);
}

// ------------------------------------------------------------------------------------------------
// leaf evaluators: value() for the warp, prior(j) for this lane's cell of word j (already masked:
// only asked for empty cells).  Twins of the oracle's evaluators.
// ------------------------------------------------------------------------------------------------
// RandomEvaluator (:137-169) on Philox, fused with get_from (:218-234) and recover_way8
// (augment.rs:75-92): draw s of original cell (r,c) is what plane s shows at its image of (r,c), so the
// recovered prior needs exactly the 8 draws of its own cell.
struct EvPhilox {
    uint32_t k0, k1, sim;
    int lane;
    __device__ __forceinline__ void init(const Dev& D, const Cursor& cs, int lane_) {
        k0 = D.seed ^ DOMAIN_EVAL, k1 = cs.game_id, sim = cs.sim_index, lane = lane_;
    }
    __device__ __forceinline__ float value() const {
        Philox4 va = philox4x32_10(sim, CELLS, 0, 0, k0, k1), vb = philox4x32_10(sim, CELLS, 1, 0, k0, k1);
        float sum = 0.f;
        sum = sum + u32_to_pm1(va.x), sum = sum + u32_to_pm1(va.y), sum = sum + u32_to_pm1(va.z), sum = sum + u32_to_pm1(va.w);
        sum = sum + u32_to_pm1(vb.x), sum = sum + u32_to_pm1(vb.y), sum = sum + u32_to_pm1(vb.z), sum = sum + u32_to_pm1(vb.w);
        return sum / 8.f;
    }
    __device__ __forceinline__ float prior(int j) const {
        uint32_t cell = (uint32_t)cell_of_bit(32u * j + lane);
        Philox4 a = philox4x32_10(sim, cell, 0, 0, k0, k1), b = philox4x32_10(sim, cell, 1, 0, k0, k1);
        float t0 = u32_to_pm1(a.x) + u32_to_pm1(a.y), t1 = u32_to_pm1(a.z) + u32_to_pm1(a.w);
        float t2 = u32_to_pm1(b.x) + u32_to_pm1(b.y), t3 = u32_to_pm1(b.z) + u32_to_pm1(b.w);
        float total = 0.f;
        total = total + t0, total = total + t1, total = total + t2, total = total + t3;
        return total / 8.f;
    }
};

// DetEval / NegEval: hash of the oriented board, evaluated on each of the 8 planes.
template <bool NEG>
struct EvHash {
    uint64_t h[8];
    int lane;
    __device__ __forceinline__ void init(const Leaf& L, const uint64_t* Z, int lane_) {
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
                if ((L.B[j] >> lane) & 1u) k ^= Z[rr * 16 + cc];
                if ((L.W[j] >> lane) & 1u) k ^= Z[256 + rr * 16 + cc];
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
    __device__ __forceinline__ float prior(int j) const {
        uint32_t bit = 32u * j + lane;
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
    __device__ __forceinline__ float prior(int j) const {
        uint32_t bit = 32u * j + lane;
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
template <class Ev>
__device__ __forceinline__ void expand_leaf(const Dev& D, Warp& w, Cursor& cs, const Leaf& L, const Ev& ev, float v) {
    const int lane = w.lane;
    const uint32_t Rm = D.R - 1, Lm = D.L - 1, Im = D.I - 1;
    const int C = CELLS - L.stones;
    const uint32_t zoff = turn_of(L.stones) > 0 ? 256u : 0u;
    const uint32_t epoch = cs.epoch;
    const uint32_t link_base = cs.lnk_head;
    const uint32_t ec = cs.exp_count;
    if (cs.lnk_head - cs.lnk_tail + (uint32_t)C > D.L) cs.err = ERR_LNK_FULL;
    if (cs.rec_head - cs.rec_tail + (uint32_t)C > D.R) cs.err = ERR_REC_FULL;
    if (ec + 1 > D.E) cs.err = ERR_EXP_FULL;
    if (cs.err) return;
    // remember the board for the collector, then only the empties are needed
    if (lane < 16) {
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (lane == j) word = L.B[j];
            if (lane == j + 8) word = L.W[j];
        }
        w.expB[(size_t)ec * 16 + lane] = word;
    }
    uint32_t e[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] = ~(L.B[j] | L.W[j]) & valid_mask(j);
    const uint64_t key = L.key;
    uint32_t base = cs.rec_head;
    int pre = 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        uint64_t ck[4];
        uint32_t pos[4];
        Slot seen[4];
        // A: first probe of every child key of this half, all loads in flight together
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = 4 * h + t;
            ck[t] = 0, pos[t] = 0;
            seen[t].key = 0, seen[t].idx = 0, seen[t].epoch = 0;
            if ((e[j] >> lane) & 1u) {
                ck[t] = key ^ w.Z[zoff + 32 * j + lane];
                pos[t] = hash_slot(ck[t]) & Im;
                seen[t] = load_slot(&w.idx[pos[t]]);
            }
        }
        // priors of this half: arithmetic under the probe latency
        float pr[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) pr[t] = ((e[4 * h + t] >> lane) & 1u) ? ev.prior(4 * h + t) : 0.f;
        // B: resolve (entry(hash).or_insert, :367-370)
        uint32_t child[4];
        unsigned newmask = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            child[t] = NIL;
            if ((e[4 * h + t] >> lane) & 1u) {
                uint32_t val;
                if (index_find(w.idx, Im, epoch, ck[t], pos[t], seen[t], val))
                    child[t] = val;
                else
                    newmask |= 1u << t;
            }
        }
        // new records are laid out contiguously in row-major child order
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            unsigned bal = __ballot_sync(FULL, (newmask >> t) & 1u);
            if ((newmask >> t) & 1u) child[t] = (base + __popc(bal & ((1u << lane) - 1u))) & Rm;
            base += __popc(bal);
        }
        // C: records / priors / link run; the key inserts are issued back to back
        uint32_t got[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int j = 4 * h + t;
            got[t] = 0;
            if ((e[j] >> lane) & 1u) {
                if ((newmask >> t) & 1u) {
                    Rec nr;
                    nr.n = 0, nr.w = 0.f, nr.p = pr[t], nr.link = NIL;
                    store_rec(&w.rec[child[t]], nr);
                    Slot want;
                    want.key = ck[t], want.idx = child[t], want.epoch = epoch;
                    got[t] = cas_slot(&w.idx[pos[t]], seen[t], want).epoch;
                } else {
                    w.rec[child[t]].p = pr[t];  // n_prob is overwritten by the last expanding parent (:371)
                }
                int k = pre + __popc(e[j] & ((1u << lane) - 1u));
                w.lnk[(link_base + (uint32_t)k) & Lm] = child[t];
            }
            pre += __popc(e[j]);
        }
        // a returned current epoch means another lane claimed the slot first: keep probing (rare)
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (((newmask >> t) & 1u) && got[t] == epoch)
                index_insert_after_conflict(w.idx, Im, epoch, ck[t], child[t], pos[t]);
    }
    __syncwarp();  // the insert paths above diverge; nothing below may overtake a lane still probing
    // the leaf itself: visit += 1, children attached (:353-357,374-377)
    if (lane == 0) {
        Rec lr = load_rec(&w.rec[L.rec]);
        lr.n += 1;
        lr.link = link_base & Lm;
        store_rec(&w.rec[L.rec], lr);
        w.exp_rec[ec] = L.rec;
        w.exp_key[ec] = key;
    }
    cs.exp_count = ec + 1;
    cs.lnk_head = link_base + (uint32_t)C;
    cs.rec_head = base;
    cs.sum_cleaf += (uint32_t)C;
    cs.evals += 1;
    cs.peak_rec = max(cs.peak_rec, base - cs.rec_tail);
    cs.peak_lnk = max(cs.peak_lnk, cs.lnk_head - cs.lnk_tail);
    backup(w, L.depth, v);
}

}  // namespace c6
