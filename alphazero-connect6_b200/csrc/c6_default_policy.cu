// DefaultPolicy — the reference's pure Monte-Carlo tree search with random rollouts — played by the GPU for many
// games at once.  One THREAD per game: the tree of a game is touched by one search at a time
// (search i+1 reads what search i wrote), so the parallelism is across games, and a thread carries its game's
// board in registers through the rollout exactly like the config-3b playout kernel (c6_microbench.cu).  The GPU keeps
// 75,776 such threads resident (128 registers each); each owns one tree slot and plays game after game in it.
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
#include <cstdlib>
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
    GameRecord* rec;   // [total]
    int32_t* root_visit;  // [total][225] or null: root node after the searches of each move
    int32_t* root_bw;
    const uint64_t* Z;
    unsigned long long* stats;  // searches, rollouts, rollout stones, select steps, children scanned, replaced, nodes, errors
    unsigned long long* next_game;  // games handed out so far
    uint32_t ncap, tmask, ecap;
    int num_iter, max_moves;
    uint32_t seed, first_id;
    long long n;      // tree slots (threads that own a tree: nodes / keys / table / chunks / path are [n][...])
    long long total;  // games to play; rec / root_visit / root_bw are [total][...]
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

__device__ __forceinline__ void dp_place(uint32_t (&B)[8], uint32_t (&W)[8], uint32_t bit, int turn) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t m = ((int)(bit >> 5) == j) ? (1u << (bit & 31)) : 0u;
        if (turn < 0) B[j] |= m; else W[j] |= m;
    }
}
// the pick-th empty cell in row-major order
__device__ __forceinline__ uint32_t dp_pick(const uint32_t (&B)[8], const uint32_t (&W)[8], uint32_t pick) {
    return pick_empty_cell(B, W, pick);
}
// does the colour that just moved own a six?  (Only that colour can have gained one, so this is search() whenever
// the position before the stone had no winner.)
__device__ __forceinline__ bool dp_won(const uint32_t (&B)[8], const uint32_t (&W)[8], int turn) {
    // the lanes of a warp are at different stone counts: the colour is SELECTED, not branched on (a branch would run the
    // check twice per pass, once for the lanes that just placed Black and once for those that placed White)
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, X[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) X[j] = turn < 0 ? B[j] : W[j];
    run6_all(X, acc);
    uint32_t any = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) any |= acc[j];
    return any != 0;
}

// One THREAD per game, one warp per block, and every lane a small state machine.
// Written as nested loops (game > stone > search > {select steps > children, rollout stones}) the 32 games of a warp spent
// most of their time waiting for each other: 7.7 of 32 lanes active per instruction (profiles/r02q), because SIMT runs every
// inner loop for the longest of its 32 trip counts.  Here one pass of the single loop gives every lane ONE unit of whatever
// its game needs next — one chunk of seven children of a select step, the expansion, one rollout stone, the backup — and the
// expensive part they have in common, placing a stone and checking its colour for a six, is one piece of code that the
// descent, the expansion and the rollout all run together.
// A lane whose game is over takes the next game from a counter and plays it in the same tree slot (the games of a warp last
// between ~60 and 225 stones; without this the warp ran at the pace of its longest game with a third of its lanes idle).
enum DpPhase : int { DP_GAME_BEGIN, DP_MOVE_BEGIN, DP_SEARCH_BEGIN, DP_SCAN, DP_EXPAND, DP_ROLL_WAIT, DP_ROLL, DP_BACKUP, DP_MOVE_END,
                     DP_GAME_END, DP_DONE };
enum DpStone : int { ST_NONE, ST_DESCEND, ST_EXPAND, ST_ROLL };
constexpr int DP_PATH_REGS = 8;  // path nodes kept in registers (deeper ones go through D.path)

__global__ void __launch_bounds__(32, 16) k_default_policy(DpDev D) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = slot < D.n;
    const long long gi = live ? slot : 0;  // lanes past the end run nothing but must stay for the warp votes
    DpTree t;
    t.nodes = D.nodes + (size_t)gi * D.ncap, t.keys = D.keys + (size_t)gi * D.ncap;
    t.table = D.table + (size_t)gi * (D.tmask + 1), t.chunks = D.chunks + (size_t)gi * D.ecap * 2;
    t.tmask = D.tmask, t.n_nodes = 0, t.n_ents = 0;
    uint32_t* path = D.path + gi;
    const size_t pstride = (size_t)D.n;
    const uint32_t k0 = D.seed ^ DOMAIN_DEFPOL;
    uint32_t k1 = 0;
    long long g = 0;  // the game this lane is playing
    unsigned long long n_stones = 0, n_scan = 0;
    uint32_t n_search = 0, n_roll = 0, n_sel = 0, n_repl = 0;  // per game: at most num_iter * 225 < 2^24

    // the game (Game::board): read once per search and written once per stone — kept in shared memory (word-major: no bank
    // conflicts) so that its sixteen registers are free for the loads a select step keeps in flight
    __shared__ uint32_t sroot[16][32];
    const int ln = (int)threadIdx.x;
    uint64_t rkey = 0, key = 0;
    int rstones = 0, winner = 0, finished = 1, n_moves = 0, err = 0;
    int phase = live ? DP_GAME_BEGIN : DP_DONE;
    uint32_t root = 0, it = 0, node = 0;
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0}, W[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t pth[DP_PATH_REGS] = {0, 0, 0, 0, 0, 0, 0, 0};
    int stones = 0, depth = 0, win = 0;
    bool won = false;                      // has the descent passed a position with a six?
    uint4 ca = make_uint4(NIL, NIL, NIL, NIL), cb = ca;  // DP_SCAN: the chunk to look at next
    bool any = false, final_select = false;              // select() in progress: best so far; is it policy()'s?
    float bp = 0.f;
    uint32_t child = 0, bit = 0;
    Philox4 rr{0, 0, 0, 0};
    uint32_t roll_n = 0;

    auto set_path = [&](int d, uint32_t nd) {
        bool in_regs = false;
#pragma unroll
        for (int i = 0; i < DP_PATH_REGS; ++i)
            if (i == d) pth[i] = nd, in_regs = true;
        if (!in_regs) path[(size_t)d * pstride] = nd;
    };

    // select() on `node` begins: first chunk in flight, or no children at all
    auto begin_select = [&](bool is_final) {
        n_sel += 1;
        any = false, bp = 0.f, final_select = is_final;
        const uint32_t ch = (uint32_t)t.nodes[node].z;
        if (ch == NIL) {
            phase = is_final ? DP_MOVE_END : DP_EXPAND;
        } else {
            ca = t.chunks[2 * (size_t)ch], cb = t.chunks[2 * (size_t)ch + 1];
            phase = DP_SCAN;
        }
    };

    // A rollout takes one Philox block per four stones.  Rollouts start only on passes whose number is a multiple of four, so
    // every rolling lane of the warp needs its next block on the same pass: one run of the Philox code per four passes, not
    // one per pass for the quarter of the lanes whose turn it is (a lane waits 1.5 passes on average for its start).
    uint32_t tick = 0;
    while (__any_sync(0xFFFFFFFFu, phase != DP_DONE)) {
        tick += 1;
        if (phase == DP_GAME_BEGIN) {  // the next game, on an empty tree
            g = (long long)atomicAdd(D.next_game, 1ull);
            if (g >= D.total) {
                phase = DP_DONE;
            } else {
                k1 = D.first_id + (uint32_t)g;
                uint4* tb = reinterpret_cast<uint4*>(t.table);
                for (uint32_t i = 0; i < (D.tmask + 1) / 4; ++i) tb[i] = make_uint4(0, 0, 0, 0);
                t.n_nodes = 0, t.n_ents = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) sroot[j][ln] = 0;
                rkey = 0, rstones = 0, winner = 0, finished = 1, n_moves = 0, err = 0;
                n_stones = 0, n_scan = 0, n_search = 0, n_roll = 0, n_sel = 0, n_repl = 0;
                phase = DP_MOVE_BEGIN;
            }
        }
        if (phase == DP_MOVE_BEGIN) {  // Agent::play's loop head, then init (:121-124): entry().or_insert keeps an existing node
            if (rstones == CELLS) {
                phase = DP_GAME_END;  // full board, no six: no move (winner None)
            } else if (D.max_moves > 0 && n_moves >= D.max_moves) {
                finished = 0, phase = DP_GAME_END;
            } else {
                root = t.find(rkey);
                if (root == NIL) root = t.create(rkey);
                it = 0, phase = DP_SEARCH_BEGIN;
            }
        }
        if (phase == DP_SEARCH_BEGIN) {  // search (:229-249), or policy (:252-261) once the budget is spent
            node = root;
            if (it == (uint32_t)D.num_iter) {
                begin_select(true);
            } else if (t.n_nodes + 1 > D.ncap || t.n_ents + 1 > D.ecap) {
                err = 1, phase = DP_GAME_END;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
                key = rkey, stones = rstones, depth = 0, won = false;  // the root has no winner: the game goes on
                pth[0] = root;
                begin_select(false);
            }
        }
        int stone = ST_NONE;  // the stone this pass places for the lane, if any
        uint32_t sbit = 0;
        // (the __syncwarp()s are what makes the lanes actually MEET at the shared pieces: without them the compiler threads each
        // phase's path through the code that follows and the warp keeps running them one after the other)
        __syncwarp();
        if (phase == DP_SCAN) {  // one chunk of select (:129-161): the LAST child with the maximal signed win ratio
            const uint32_t next = cb.w;
            uint4 na = make_uint4(NIL, NIL, NIL, NIL), nb = na;
            if (next != NIL) na = t.chunks[2 * (size_t)next], nb = t.chunks[2 * (size_t)next + 1];
            const uint32_t e[7] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z};
            int2 cn[7];
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                cn[i] = make_int2(0, 0);
                if (e[i] != NIL) cn[i] = *reinterpret_cast<const int2*>(&t.nodes[e[i] & 0xFFFFFFu]);  // {visit, black_win}
            }
            const int turn = turn_of(final_select ? rstones : stones);
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                if (e[i] == NIL) continue;  // slots fill in order: everything behind a free slot is free
                // black_win / (1 + visit); a zero numerator would send the IEEE division into its subroutine, and most
                // children have no black win yet: stand-in operand, exact result (+0) selected afterwards.  The stand-in goes
                // through an empty asm: the compiler otherwise sees that the quotient is discarded whenever the stand-in is
                // used and divides the raw numerator after all (r02v: the subroutine was 15 % of all instructions).
                float num = cn[i].y != 0 ? (float)cn[i].y : 1.f;
                asm volatile("" : "+f"(num));
                float p = num / (1.f + (float)cn[i].x);
                if (cn[i].y == 0) p = 0.f;
                if (turn > 0) p = -p;
                if (!any || !(p < bp)) any = true, bp = p, child = e[i] & 0xFFFFFFu, bit = e[i] >> 24;
                n_scan += 1;
            }
            if (next != NIL) {
                ca = na, cb = nb;
            } else if (final_select) {
                phase = DP_MOVE_END;
            } else if (any && bp != 0.f) {
                stone = ST_DESCEND, sbit = bit;  // the descent does not stop at a won position (:236-239)
            } else {
                phase = DP_EXPAND;
            }
        }
        if (phase == DP_EXPAND) {  // expand (:164-184) unless the search ends here
            n_search += 1;
            if (won || stones == CELLS) {
                // search_winner() != None: return without expanding (:241-243); a full board without a six: the reference's
                // expand() would panic (:165) — the search changes nothing
                it += 1, phase = DP_SEARCH_BEGIN;
            } else {
                const uint32_t ue = philox4x32_10((uint32_t)rstones, it, 0xFFFFFFFFu, 0, k0, k1).x;
                sbit = dp_pick(B, W, (uint32_t)(((unsigned long long)ue * (unsigned long long)(CELLS - stones)) >> 32));
                stone = ST_EXPAND;
            }
        } else if (phase == DP_ROLL || (phase == DP_ROLL_WAIT && (tick & 3u) == 0)) {  // update (:190-226): one stone of the random rollout
            phase = DP_ROLL;
            if (win != 0 || stones == CELLS) {
                phase = DP_BACKUP;
            } else {
                if ((roll_n & 3) == 0) rr = philox4x32_10((uint32_t)rstones, it, roll_n >> 2, 0, k0, k1);
                const uint32_t u = (roll_n & 3) == 0 ? rr.x : (roll_n & 3) == 1 ? rr.y : (roll_n & 3) == 2 ? rr.z : rr.w;
                sbit = dp_pick(B, W, (uint32_t)(((unsigned long long)u * (unsigned long long)(CELLS - stones)) >> 32));
                stone = ST_ROLL;
            }
        }
        // the part every kind of step shares: place the stone, has its colour a six now?
        __syncwarp();
        if (stone != ST_NONE) {
            const int turn = turn_of(stones);
            dp_place(B, W, sbit, turn);
            stones += 1;
            const bool six = dp_won(B, W, turn);
            if (stone == ST_ROLL) {
                roll_n += 1;
                if (six) win = turn;
            } else {
                key ^= D.Z[(turn > 0 ? 256 : 0) + sbit];
                depth += 1;
                if (stone == ST_DESCEND) {
                    node = child;
                    set_path(depth, node);
                    won = won || six;  // a six, once on the board, stays
                    begin_select(false);
                } else {
                    uint32_t c = t.find(key);
                    if (c != NIL) {
                        t.nodes[c] = make_int4(0, 0, (int)NIL, (int)NIL);  // HashMap::insert replaces the node
                        n_repl += 1;
                    } else {
                        c = t.create(key);
                    }
                    const uint32_t entry = c | (sbit << 24);
                    int4 pn = t.nodes[node];
                    const uint32_t tw = (uint32_t)pn.w;
                    if (tw == NIL || (tw >> 24) == 7u) {  // first child, or the tail chunk is full: a new chunk
                        const uint32_t nc = t.n_ents++;
                        t.chunks[2 * (size_t)nc] = make_uint4(entry, NIL, NIL, NIL);
                        t.chunks[2 * (size_t)nc + 1] = make_uint4(NIL, NIL, NIL, NIL);
                        if (tw != NIL) reinterpret_cast<uint32_t*>(&t.chunks[2 * (size_t)(tw & 0xFFFFFFu)])[7] = nc; else pn.z = (int)nc;
                        pn.w = (int)(nc | (1u << 24));
                    } else {
                        const uint32_t oc = tw & 0xFFFFFFu, f = tw >> 24;
                        reinterpret_cast<uint32_t*>(&t.chunks[2 * (size_t)oc])[f] = entry;
                        pn.w = (int)(oc | ((f + 1u) << 24));
                    }
                    t.nodes[node] = pn;
                    set_path(depth, c);
                    n_roll += 1;
                    win = six ? turn : 0, roll_n = 0, phase = DP_ROLL_WAIT;
                }
            }
        }
        if (phase == DP_BACKUP) {  // visit / black_win along the path — unless nobody won (:205-207)
            n_stones += roll_n;
            if (win != 0) {
                // every load of the path's statistics is in flight before the first store: one memory round trip for the
                // whole path (a load-add-store loop took one per node with the other 31 lanes waiting)
                const int bw = win < 0 ? 1 : 0;
                int2 v[DP_PATH_REGS];
#pragma unroll
                for (int d = 0; d < DP_PATH_REGS; ++d)
                    if (d <= depth) v[d] = *reinterpret_cast<const int2*>(&t.nodes[pth[d]]);
#pragma unroll
                for (int d = 0; d < DP_PATH_REGS; ++d)
                    if (d <= depth) *reinterpret_cast<int2*>(&t.nodes[pth[d]]) = make_int2(v[d].x + 1, v[d].y + bw);
                for (int d = DP_PATH_REGS; d <= depth; ++d) {
                    int2* pn = reinterpret_cast<int2*>(&t.nodes[path[(size_t)d * pstride]]);
                    int2 x = *pn;
                    x.x += 1, x.y += bw;
                    *pn = x;
                }
            }
            it += 1, phase = DP_SEARCH_BEGIN;
        }
        if (phase == DP_MOVE_END) {  // policy (:252-261) has scanned the root; Agent::play: Game::set, record, is_game_end
            if (!(any && bp != 0.f)) {
                const uint32_t u = philox4x32_10((uint32_t)rstones, 0xFFFFFFFFu, 0, 0, k0, k1).x;
#pragma unroll
                for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
                bit = dp_pick(B, W, (uint32_t)(((unsigned long long)u * (unsigned long long)(CELLS - rstones)) >> 32));
            }
            if (D.root_visit) {
                const int4 rn = t.nodes[root];
                D.root_visit[(size_t)g * CELLS + n_moves] = rn.x;
                D.root_bw[(size_t)g * CELLS + n_moves] = rn.y;
            }
            const int turn = turn_of(rstones);
            sroot[(turn > 0 ? 8 : 0) + (bit >> 5)][ln] |= 1u << (bit & 31);
            rkey ^= D.Z[(turn > 0 ? 256 : 0) + bit];
            rstones += 1;
            D.rec[g].moves[n_moves++] = (uint8_t)cell_of_bit(bit);
#pragma unroll
            for (int j = 0; j < 8; ++j) B[j] = sroot[j][ln], W[j] = sroot[8 + j][ln];
            if (dp_won(B, W, turn)) winner = turn, phase = DP_GAME_END; else phase = DP_MOVE_BEGIN;
        }
        if (phase == DP_GAME_END) {
            GameRecord* out = D.rec + g;
            out->winner = (int8_t)winner;
            out->n_moves = (uint8_t)n_moves;
            out->finished = (uint8_t)(err ? 0 : finished);
            out->pad = (uint8_t)err;
            out->game_id = D.first_id + (uint32_t)g;
            atomicAdd(D.stats + 0, (unsigned long long)n_search), atomicAdd(D.stats + 1, (unsigned long long)n_roll), atomicAdd(D.stats + 2, n_stones);
            atomicAdd(D.stats + 3, (unsigned long long)n_sel), atomicAdd(D.stats + 4, n_scan), atomicAdd(D.stats + 5, (unsigned long long)n_repl);
            atomicAdd(D.stats + 6, (unsigned long long)t.n_nodes);
            if (err) atomicAdd(D.stats + 7, 1ull);
            phase = DP_GAME_BEGIN;
        }
    }
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
    // Tree slots: one per thread the GPU keeps resident (a lane that finishes a game takes the next one into the same slot),
    // bounded by the free memory; records and root traces are per game.
    const size_t per_slot = (size_t)ncap * (16 + 8) + (size_t)tsize * 4 + (size_t)ecap * 32 + 232 * 4;
    const size_t per_game = sizeof(GameRecord) + (root_visit ? 2 * CELLS * 4 : 0);
    int sms = 0, blocks_per_sm = 0;
    DP_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    DP_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_default_policy, 32, 0));
    const long long batch = std::min<long long>(n_games, 1ll << 22);  // games per launch
    size_t free_b = 0, total_b = 0;
    DP_CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t for_slots = free_b * 8 / 10 > (size_t)batch * per_game ? free_b * 8 / 10 - (size_t)batch * per_game : 0;
    long long slots = std::min<long long>((long long)sms * std::max(blocks_per_sm, 1) * 32, (long long)(for_slots / per_slot));
    if (slots < 1) throw std::runtime_error("default_policy_play: not enough device memory for one game");
    if (const char* v = std::getenv("C6_DP_SLOTS"))  // tuning / tests: fewer slots than games makes every lane play several
        if (*v && std::atoll(v) > 0) slots = std::min<long long>(slots, std::atoll(v));
    slots = std::min(slots, batch);
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
    DP_CK(cudaMalloc(&D.nodes, (size_t)slots * ncap * 16));
    DP_CK(cudaMalloc(&D.keys, (size_t)slots * ncap * 8));
    DP_CK(cudaMalloc(&D.table, (size_t)slots * tsize * 4));
    DP_CK(cudaMalloc(&D.chunks, (size_t)slots * ecap * 32));
    DP_CK(cudaMalloc(&D.path, (size_t)slots * 232 * 4));
    DP_CK(cudaMalloc(&D.rec, (size_t)batch * sizeof(GameRecord)));
    DP_CK(cudaMalloc(&D.stats, 9 * sizeof(unsigned long long)));
    DP_CK(cudaMemset(D.stats, 0, 9 * sizeof(unsigned long long)));
    D.next_game = D.stats + 8;
    if (root_visit) {
        DP_CK(cudaMalloc(&D.root_visit, (size_t)batch * CELLS * 4));
        DP_CK(cudaMalloc(&D.root_bw, (size_t)batch * CELLS * 4));
    }
    D.Z = dZ, D.ncap = ncap, D.tmask = tsize - 1, D.ecap = ecap, D.num_iter = num_iter, D.max_moves = max_moves, D.seed = seed;
    for (long long done = 0; done < n_games; done += batch) {
        const long long nb = std::min(batch, n_games - done);
        D.n = std::min(slots, nb), D.total = nb, D.first_id = first_game_id + (uint32_t)done;
        DP_CK(cudaMemsetAsync(D.next_game, 0, sizeof(unsigned long long), s));
        DP_CK(cudaMemsetAsync(D.rec, 0, (size_t)nb * sizeof(GameRecord), s));
        if (root_visit) {
            DP_CK(cudaMemsetAsync(D.root_visit, 0, (size_t)nb * CELLS * 4, s));
            DP_CK(cudaMemsetAsync(D.root_bw, 0, (size_t)nb * CELLS * 4, s));
        }
        DP_CK(cudaEventRecord(a, s));
        k_default_policy<<<(unsigned)((D.n + 31) / 32), 32, 0, s>>>(D);
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
