// DefaultPolicy — the reference's pure Monte-Carlo tree search with random rollouts — played by the GPU for many
// games at once.  One THREAD per game: the tree of a game is touched by one search at a time
// (search i+1 reads what search i wrote), so the parallelism is across games, and a thread carries its game's
// board in registers through the rollout exactly like the config-3b playout kernel (c6_microbench.cu).
//   DefaultPolicy::{init,select,expand,update,search,policy,next}   LC/policy/default_policy/mod.rs:121-271
//   Agent::play                                                     LC/agent/agent_impl/mod.rs:121-178
// (LC = /root/reference/Connect6/libconnect6/src).  What the reference keeps in `HashMap<u64, Node>`:
//   node  {visit, black_win, head, tail}  16 B  — `next_node: Vec<u64>` is a list of CHUNKS in push order
//   chunk {7 x (child | move << 24), next} 32 B  — one entry per expand(); the same child may be listed twice (:181).
//         One sector holds seven children: select() reads a chunk, then has the seven child nodes and the next chunk in
//         flight together — one memory round trip per seven children where a list of single entries took one per child
//         (a search scans ~150 entries in 3.9 select steps, and that dependent chain was what a game's time went into).
//         tail = chunk | fill << 24 (fill 1..7); free slots hold NIL
//   key   64-bit Zobrist board key -> node, open addressing; `map.insert` on an existing board RESETS that node
//         (visit = black_win = 0, no children, :174), `entry().or_insert` in init() does not (:123)
// Random choices are the Philox streams the oracle documents (oracle/c6_default_policy.hpp): key (seed ^ "DPOL", game),
// counters (stones at next(), search number, ...); `choose` on n items = item (u32 * n) >> 32 of the row-major list.
#include <algorithm>
#include <stdexcept>
#include <vector>

#include "c6_board.cuh"
#include "c6_engine.cuh"
#include "c6_rng.cuh"

namespace c6 {

constexpr uint32_t DOMAIN_DEFPOL = 0x44504F4Cu;  // "DPOL"

struct DpDev {
    int4* nodes;       // [n][ncap] {visit, black_win, head, tail}
    uint64_t* keys;    // [n][ncap]
    uint32_t* table;   // [n][tmask + 1] node index + 1, 0 = empty
    uint4* chunks;     // [n][ecap][2]
    uint32_t* path;    // [232][n] node indices of the current search (game-minor: coalesced across a warp)
    GameRecord* rec;   // [n]
    int32_t* root_visit;  // [n][225] or null: root node after the searches of each move
    int32_t* root_bw;
    const uint64_t* Z;
    unsigned long long* stats;  // searches, rollouts, rollout stones, select steps, children scanned, replaced, nodes, errors
    uint32_t ncap, tmask, ecap;
    int num_iter, max_moves;
    uint32_t seed, first_id;
    long long n;
};

struct DpTree {
    int4* nodes;
    uint64_t* keys;
    uint32_t* table;
    uint4* chunks;
    uint32_t tmask, n_nodes, n_ents;  // n_ents: chunks in use
    __device__ __forceinline__ uint32_t find(uint64_t key) const {
        uint32_t pos = (uint32_t)(key ^ (key >> 32)) & tmask;
        for (;;) {
            const uint32_t v = table[pos];
            if (v == 0) return NIL;
            if (keys[v - 1] == key) return v - 1;
            pos = (pos + 1) & tmask;
        }
    }
    __device__ __forceinline__ uint32_t create(uint64_t key) {  // key known to be absent
        const uint32_t id = n_nodes++;
        nodes[id] = make_int4(0, 0, (int)NIL, (int)NIL);
        keys[id] = key;
        uint32_t pos = (uint32_t)(key ^ (key >> 32)) & tmask;
        while (table[pos] != 0) pos = (pos + 1) & tmask;
        table[pos] = id + 1;
        return id;
    }
};

// select (:129-161): the LAST child with the maximal signed win ratio, if that ratio is non-zero
__device__ __forceinline__ bool dp_select(const DpTree& t, uint32_t node, int turn, uint32_t& child, uint32_t& bit,
                                          unsigned long long& scanned) {
    uint32_t ch = (uint32_t)t.nodes[node].z;
    if (ch == NIL) return false;
    bool any = false;
    float bp = 0.f;
    uint4 a = t.chunks[2 * (size_t)ch], b = t.chunks[2 * (size_t)ch + 1];
    for (;;) {
        // the next chunk and this chunk's seven child nodes are all requested before the first of them is looked at
        const uint32_t next = b.w;
        uint4 na = make_uint4(NIL, NIL, NIL, NIL), nb = na;
        if (next != NIL) na = t.chunks[2 * (size_t)next], nb = t.chunks[2 * (size_t)next + 1];
        const uint32_t e[7] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z};
        int2 cn[7];
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            cn[i] = make_int2(0, 0);
            if (e[i] != NIL) cn[i] = *reinterpret_cast<const int2*>(&t.nodes[e[i] & 0xFFFFFFu]);  // {visit, black_win}
        }
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            if (e[i] == NIL) continue;  // slots fill in order: everything behind a free slot is free
            // black_win / (1 + visit); a zero numerator would send the IEEE division into its subroutine, and most
            // children have no black win yet: stand-in operand, exact result (+0) selected afterwards
            float p = (cn[i].y != 0 ? (float)cn[i].y : 1.f) / (1.f + (float)cn[i].x);
            if (cn[i].y == 0) p = 0.f;
            if (turn > 0) p = -p;
            if (!any || !(p < bp)) any = true, bp = p, child = e[i] & 0xFFFFFFu, bit = e[i] >> 24;
            scanned += 1;
        }
        if (next == NIL) break;
        a = na, b = nb;
    }
    return any && bp != 0.f;
}

__device__ __forceinline__ void dp_place(uint32_t (&B)[8], uint32_t (&W)[8], uint32_t bit, int turn) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t m = ((int)(bit >> 5) == j) ? (1u << (bit & 31)) : 0u;
        if (turn < 0) B[j] |= m; else W[j] |= m;
    }
}
// the pick-th empty cell in row-major order
__device__ __forceinline__ uint32_t dp_pick(const uint32_t (&B)[8], const uint32_t (&W)[8], uint32_t pick) {
    uint32_t bit = 0;
    bool found = false;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t e = ~(B[j] | W[j]) & valid_mask(j);
        const uint32_t c = (uint32_t)__popc(e);
        if (!found) {
            if (pick < c) {
                bit = 32u * j + __fns(e, 0, (int)pick + 1);
                found = true;
            } else {
                pick -= c;
            }
        }
    }
    return bit;
}
// does the colour that just moved own a six?  (Only that colour can have gained one, so this is search() whenever
// the position before the stone had no winner.)
__device__ __forceinline__ bool dp_won(const uint32_t (&B)[8], const uint32_t (&W)[8], int turn) {
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (turn < 0) {
        run6_all(B, acc);
    } else {
        run6_all(W, acc);
    }
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) any |= acc[j];
    return any != 0;
}

// One warp per block: a block retires as soon as its 32 games are over instead of waiting for the longest of 128.
__global__ void __launch_bounds__(32) k_default_policy(DpDev D) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= D.n) return;
    DpTree t;
    t.nodes = D.nodes + (size_t)g * D.ncap, t.keys = D.keys + (size_t)g * D.ncap;
    t.table = D.table + (size_t)g * (D.tmask + 1), t.chunks = D.chunks + (size_t)g * D.ecap * 2;
    t.tmask = D.tmask, t.n_nodes = 0, t.n_ents = 0;
    uint32_t* path = D.path + g;
    const size_t pstride = (size_t)D.n;
    const uint32_t k0 = D.seed ^ DOMAIN_DEFPOL, k1 = D.first_id + (uint32_t)g;
    unsigned long long n_search = 0, n_roll = 0, n_stones = 0, n_sel = 0, n_scan = 0, n_repl = 0;

    // the game (Game::board): read once per search and written once per stone — kept in shared memory (word-major: no bank
    // conflicts) so that its sixteen registers are free for the loads a select step keeps in flight
    __shared__ uint32_t sroot[16][32];
    const int ln = (int)threadIdx.x;
#pragma unroll
    for (int j = 0; j < 16; ++j) sroot[j][ln] = 0;
    uint64_t rkey = 0;
    int rstones = 0, winner = 0, finished = 1, n_moves = 0, err = 0;
    GameRecord* out = D.rec + g;

    for (;;) {
        if (rstones == CELLS) break;                                // full board, no six: no move (winner None)
        if (D.max_moves > 0 && n_moves >= D.max_moves) { finished = 0; break; }
        // init (:121-124): entry().or_insert — an existing node keeps its statistics and children
        uint32_t root = t.find(rkey);
        if (root == NIL) root = t.create(rkey);
        for (uint32_t it = 0; it < (uint32_t)D.num_iter; ++it) {  // search (:229-249)
            if (t.n_nodes + 1 > D.ncap || t.n_ents + 1 > D.ecap) { err = 1; break; }
            uint32_t B[8], W[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
            uint64_t key = rkey;
            int stones = rstones, depth = 0;
            uint32_t node = root;
            path[0] = root;
            bool won = false;  // does the position at `node` have a winner?  (the root has none: the game goes on)
            for (;;) {
                uint32_t child = 0, bit = 0;
                n_sel += 1;
                if (!dp_select(t, node, turn_of(stones), child, bit, n_scan)) break;
                const int turn = turn_of(stones);
                dp_place(B, W, bit, turn);
                key ^= D.Z[(turn > 0 ? 256 : 0) + bit];
                stones += 1, depth += 1;
                node = child;
                path[(size_t)depth * pstride] = node;
                // the descent does not stop at a won position (:236-239) — remember whether one was passed: a six,
                // once on the board, stays
                won = won || dp_won(B, W, turn);
            }
            n_search += 1;
            if (won) continue;  // search_winner() != None: return without expanding (:241-243)
            if (stones == CELLS) continue;  // full board, no six: the reference's expand() would panic (:165); the search changes nothing
            // expand (:164-184)
            const int turn = turn_of(stones);
            const uint32_t ue = philox4x32_10((uint32_t)rstones, it, 0xFFFFFFFFu, 0, k0, k1).x;
            const uint32_t ebit = dp_pick(B, W, (uint32_t)(((unsigned long long)ue * (unsigned long long)(CELLS - stones)) >> 32));
            dp_place(B, W, ebit, turn);
            key ^= D.Z[(turn > 0 ? 256 : 0) + ebit];
            stones += 1, depth += 1;
            uint32_t child = t.find(key);
            if (child != NIL) {
                t.nodes[child] = make_int4(0, 0, (int)NIL, (int)NIL);  // HashMap::insert replaces the node
                n_repl += 1;
            } else {
                child = t.create(key);
            }
            {
                const uint32_t entry = child | (ebit << 24);
                int4 pn = t.nodes[node];
                const uint32_t tw = (uint32_t)pn.w;
                if (tw == NIL || (tw >> 24) == 7u) {  // first child, or the tail chunk is full: a new chunk
                    const uint32_t c = t.n_ents++;
                    t.chunks[2 * (size_t)c] = make_uint4(entry, NIL, NIL, NIL);
                    t.chunks[2 * (size_t)c + 1] = make_uint4(NIL, NIL, NIL, NIL);
                    if (tw != NIL) reinterpret_cast<uint32_t*>(&t.chunks[2 * (size_t)(tw & 0xFFFFFFu)])[7] = c; else pn.z = (int)c;
                    pn.w = (int)(c | (1u << 24));
                } else {
                    const uint32_t c = tw & 0xFFFFFFu, f = tw >> 24;
                    reinterpret_cast<uint32_t*>(&t.chunks[2 * (size_t)c])[f] = entry;
                    pn.w = (int)(c | ((f + 1u) << 24));
                }
                t.nodes[node] = pn;
            }
            path[(size_t)depth * pstride] = child;
            // update (:190-226): random rollout from the new position, then visit / black_win along the path
            n_roll += 1;
            int win = dp_won(B, W, turn) ? turn : 0;
            Philox4 r{0, 0, 0, 0};
            uint32_t n = 0;
            while (win == 0 && stones < CELLS) {
                if ((n & 3) == 0) r = philox4x32_10((uint32_t)rstones, it, n >> 2, 0, k0, k1);
                const uint32_t u = (n & 3) == 0 ? r.x : (n & 3) == 1 ? r.y : (n & 3) == 2 ? r.z : r.w;
                const uint32_t pb = dp_pick(B, W, (uint32_t)(((unsigned long long)u * (unsigned long long)(CELLS - stones)) >> 32));
                const int tn = turn_of(stones);
                dp_place(B, W, pb, tn);
                stones += 1, n += 1;
                if (dp_won(B, W, tn)) win = tn;
            }
            n_stones += n;
            if (win == 0) continue;  // nobody won: no update (:205-207)
            const int bw = win < 0 ? 1 : 0;
            for (int d = 0; d <= depth; ++d) {
                int4* pn = &t.nodes[path[(size_t)d * pstride]];
                int4 v = *pn;
                v.x += 1, v.y += bw;
                *pn = v;
            }
        }
        if (err) break;
        // policy (:252-261)
        uint32_t child = 0, bit = 0;
        n_sel += 1;
        if (!dp_select(t, root, turn_of(rstones), child, bit, n_scan)) {
            const uint32_t u = philox4x32_10((uint32_t)rstones, 0xFFFFFFFFu, 0, 0, k0, k1).x;
            uint32_t B[8], W[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
            bit = dp_pick(B, W, (uint32_t)(((unsigned long long)u * (unsigned long long)(CELLS - rstones)) >> 32));
        }
        if (D.root_visit) {
            const int4 rn = t.nodes[root];
            D.root_visit[(size_t)g * CELLS + n_moves] = rn.x;
            D.root_bw[(size_t)g * CELLS + n_moves] = rn.y;
        }
        // Agent::play: Game::set, record, is_game_end
        const int turn = turn_of(rstones);
        sroot[(turn > 0 ? 8 : 0) + (bit >> 5)][ln] |= 1u << (bit & 31);
        rkey ^= D.Z[(turn > 0 ? 256 : 0) + bit];
        rstones += 1;
        out->moves[n_moves++] = (uint8_t)cell_of_bit(bit);
        {
            uint32_t B[8], W[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
            if (dp_won(B, W, turn)) { winner = turn; break; }
        }
    }
    out->winner = (int8_t)winner;
    out->n_moves = (uint8_t)n_moves;
    out->finished = (uint8_t)(err ? 0 : finished);
    out->pad = (uint8_t)err;
    out->game_id = D.first_id + (uint32_t)g;
    atomicAdd(D.stats + 0, n_search), atomicAdd(D.stats + 1, n_roll), atomicAdd(D.stats + 2, n_stones);
    atomicAdd(D.stats + 3, n_sel), atomicAdd(D.stats + 4, n_scan), atomicAdd(D.stats + 5, n_repl);
    atomicAdd(D.stats + 6, (unsigned long long)t.n_nodes);
    if (err) atomicAdd(D.stats + 7, 1ull);
}

#define DP_CK(call)                                                                 \
    do {                                                                            \
        cudaError_t _e = (call);                                                    \
        if (_e != cudaSuccess) throw std::runtime_error(cudaGetErrorString(_e));    \
    } while (0)

void make_zobrist_table(uint64_t* Zb);  // c6_engine.cu

// n_games of DefaultPolicy self-play on `device`, in batches sized to the free HBM.  stats[8] as in DpDev::stats.
// Returns the milliseconds spent inside the kernel (CUDA events).
float default_policy_play(int num_iter, uint32_t seed, uint32_t first_game_id, long long n_games, int max_moves, int device,
                          GameRecord* out, int32_t* root_visit, int32_t* root_bw, unsigned long long* stats) {
    // num_iter == 0 panics in the reference (policy() looks the root up in a map nobody filled, mod.rs:131)
    if (num_iter < 1 || n_games < 0) throw std::runtime_error("default_policy_play: num_iter must be >= 1 and n_games >= 0");
    DP_CK(cudaSetDevice(device));
    for (int i = 0; i < 8; ++i) stats[i] = 0;
    if (n_games == 0) return 0.f;
    const int moves = (max_moves > 0 && max_moves < CELLS) ? max_moves : CELLS;
    // every search creates at most one node and one chunk of entries; init() creates at most one node per move
    const long long ecap64 = std::max<long long>(1, (long long)num_iter * moves);
    // child entries pack `child | move << 24`: node ids must stay below 2^24 (num_iter up to ~74,000 for whole games)
    if (ecap64 + moves + 2 >= (1ll << 24))
        throw std::runtime_error("default_policy_play: num_iter * moves must stay below 2^24 nodes per game (child ids are packed in 24 bits)");
    const uint32_t ecap = (uint32_t)ecap64;
    const uint32_t ncap = ecap + (uint32_t)moves + 2;
    uint32_t tsize = 64;
    while (tsize < 2 * ncap) tsize <<= 1;
    const size_t per_game = (size_t)ncap * (16 + 8) + (size_t)tsize * 4 + (size_t)ecap * 32 + 232 * 4 + sizeof(GameRecord) + 2 * CELLS * 4;
    size_t free_b = 0, total_b = 0;
    DP_CK(cudaMemGetInfo(&free_b, &total_b));
    long long batch = (long long)((free_b * 8 / 10) / per_game);
    batch = std::min<long long>(batch, 1 << 20);
    if (batch < 1) throw std::runtime_error("default_policy_play: not enough device memory for one game");
    batch = std::min(batch, n_games);
    DpDev D{};
    uint64_t hz[512];
    make_zobrist_table(hz);
    uint64_t* dZ = nullptr;
    cudaStream_t s;
    cudaEvent_t a, b;
    float total_ms = 0.f;
    DP_CK(cudaStreamCreate(&s));
    DP_CK(cudaEventCreate(&a));
    DP_CK(cudaEventCreate(&b));
    DP_CK(cudaMalloc(&dZ, sizeof hz));
    DP_CK(cudaMemcpy(dZ, hz, sizeof hz, cudaMemcpyHostToDevice));
    DP_CK(cudaMalloc(&D.nodes, (size_t)batch * ncap * 16));
    DP_CK(cudaMalloc(&D.keys, (size_t)batch * ncap * 8));
    DP_CK(cudaMalloc(&D.table, (size_t)batch * tsize * 4));
    DP_CK(cudaMalloc(&D.chunks, (size_t)batch * ecap * 32));
    DP_CK(cudaMalloc(&D.path, (size_t)batch * 232 * 4));
    DP_CK(cudaMalloc(&D.rec, (size_t)batch * sizeof(GameRecord)));
    DP_CK(cudaMalloc(&D.stats, 8 * sizeof(unsigned long long)));
    DP_CK(cudaMemset(D.stats, 0, 8 * sizeof(unsigned long long)));
    if (root_visit) {
        DP_CK(cudaMalloc(&D.root_visit, (size_t)batch * CELLS * 4));
        DP_CK(cudaMalloc(&D.root_bw, (size_t)batch * CELLS * 4));
    }
    D.Z = dZ, D.ncap = ncap, D.tmask = tsize - 1, D.ecap = ecap, D.num_iter = num_iter, D.max_moves = max_moves, D.seed = seed;
    for (long long done = 0; done < n_games; done += batch) {
        const long long nb = std::min(batch, n_games - done);
        D.n = nb, D.first_id = first_game_id + (uint32_t)done;
        DP_CK(cudaMemsetAsync(D.table, 0, (size_t)nb * tsize * 4, s));
        DP_CK(cudaMemsetAsync(D.rec, 0, (size_t)nb * sizeof(GameRecord), s));
        if (root_visit) {
            DP_CK(cudaMemsetAsync(D.root_visit, 0, (size_t)nb * CELLS * 4, s));
            DP_CK(cudaMemsetAsync(D.root_bw, 0, (size_t)nb * CELLS * 4, s));
        }
        DP_CK(cudaEventRecord(a, s));
        k_default_policy<<<(unsigned)((nb + 31) / 32), 32, 0, s>>>(D);
        DP_CK(cudaEventRecord(b, s));
        DP_CK(cudaGetLastError());
        DP_CK(cudaEventSynchronize(b));
        float ms = 0.f;
        DP_CK(cudaEventElapsedTime(&ms, a, b));
        total_ms += ms;
        DP_CK(cudaMemcpy(out + done, D.rec, (size_t)nb * sizeof(GameRecord), cudaMemcpyDeviceToHost));
        if (root_visit) {
            DP_CK(cudaMemcpy(root_visit + done * CELLS, D.root_visit, (size_t)nb * CELLS * 4, cudaMemcpyDeviceToHost));
            DP_CK(cudaMemcpy(root_bw + done * CELLS, D.root_bw, (size_t)nb * CELLS * 4, cudaMemcpyDeviceToHost));
        }
    }
    DP_CK(cudaMemcpy(stats, D.stats, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(D.nodes), cudaFree(D.keys), cudaFree(D.table), cudaFree(D.chunks), cudaFree(D.path), cudaFree(D.rec), cudaFree(D.stats);
    cudaFree(D.root_visit), cudaFree(D.root_bw), cudaFree(dZ);
    cudaEventDestroy(a), cudaEventDestroy(b), cudaStreamDestroy(s);
    if (stats[7]) throw std::runtime_error("default_policy_play: node pool of a game overflowed");
    return total_ms;
}

}  // namespace c6
