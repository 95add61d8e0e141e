// ORACLE — TEST INFRASTRUCTURE ONLY (see c6_rng.hpp header note).
// CPU model of the device node pool (alphazero-connect6_b200/csrc/c6_search.cuh, c6_device.cuh):
// records + child-link runs, NO board-key index over the children.  A child of a node being
// expanded is shared with an existing record ("transposition", the reference's
// `entry(hash).or_insert`, alphazero_policy/mod.rs:367-370) by following links:
//   (F)  the sibling rule: if the last stone on the path has the colour of the stone about to be
//        placed, child B+c == (P+c)+a_d, where P+c is a sibling of B whose record the select step at
//        P has just read; if that sibling is expanded, its link run names the shared record;
//   (S1) the same through an earlier path stone s' of that colour: co-parent B-s'+c is looked up in
//        the small key table of EXPANDED nodes (one insert per simulation, not one per child);
//   (S2) expanded nodes that lost every expanded parent when the tree was collected after a real
//        move ("orphans", flagged per stone count by the collector) are looked up in that table.
// This program plays games with the map-based restatement (c6_oracle.hpp, the parity oracle) and
// with the model side by side and compares every move, root statistic and child visit count, so
// the rules above are checked on the CPU before the CUDA kernels are (which are then checked
// against the oracle on the GPU).  Usage: pool_model [games] [quick]
#include <bitset>
#include <cstdio>
#include <cstdlib>
#include <unordered_map>

#include "c6_oracle.hpp"

extern "C" void det_eval_callback(int player, float* values, float* policies, int len);
extern "C" void neg_eval_callback(int player, float* values, float* policies, int len);

using namespace c6o;

namespace {

constexpr uint32_t NIL = 0xFFFFFFFFu;
struct Rec {
    int32_t n = 0;
    float w = 0.f, p = 0.f;
    uint32_t link = NIL;
};
struct ExpEntry {
    Board b;
    uint64_t key;
    uint32_t rec;
    int stones;
};

inline int turn_of(int stones) { return stones == 0 ? -1 : ((((stones - 1) / 2) % 2 == 0) ? 1 : -1); }
inline uint64_t zkey(int turn, int cell) { return zobrist().z[turn > 0][cell]; }

struct Counters {
    uint64_t expansions = 0, fast_hits = 0, s1_runs = 0, s1_hits = 0, s2_runs = 0, s2_hits = 0, orphans = 0, deep = 0;
};

struct PoolModel {
    std::vector<Rec> rec;
    std::vector<uint32_t> lnk;
    std::vector<ExpEntry> exp;
    std::unordered_map<uint64_t, uint32_t> exph;  // key -> record, expanded nodes only
    std::bitset<BOARD_CAPACITY + 2> orph;
    Board root;
    uint64_t root_key = 0;
    uint32_t root_rec = 0;
    int root_stones = 0;
    HyperParameter param;
    AlphaZero evalbox;  // only get_from() is used
    Counters cnt;
    bool check_all_sources = true;
    // trace of the last move
    std::array<int32_t, BOARD_CAPACITY> last_child_visits{};
    int32_t last_root_visit = 0;
    float last_root_qsum = 0.f;

    PoolModel(Evaluator* e, HyperParameter p) : param(p), evalbox(e, p) {
        root = empty_board();
        rec.push_back(Rec{});
    }

    static std::vector<int> empties(const Board& b) {
        std::vector<int> v;
        for (int i = 0; i < BOARD_CAPACITY; ++i)
            if (b[i / BOARD_SIZE][i % BOARD_SIZE] == Player::None) v.push_back(i);
        return v;
    }
    static int rank_below(const Board& b, int cell) {  // empty cells of b before `cell` in row-major order
        int r = 0;
        for (int i = 0; i < cell; ++i) r += b[i / BOARD_SIZE][i % BOARD_SIZE] == Player::None;
        return r;
    }

    void search_once() {
        Board b = root;
        uint64_t key = root_key;
        int stones = root_stones;
        uint32_t cur = root_rec;
        std::vector<uint32_t> path;
        std::vector<int> cells;
        uint32_t parent_link = NIL;  // link run of the node the last step selected from (the "stage")
        Board parent_board = root;
        for (;;) {
            const Rec r = rec[cur];
            if (r.n == 0) break;
            path.push_back(cur);
            std::vector<int> em = empties(b);
            const int C = (int)em.size(), turn = turn_of(stones);
            int pick = 0;
            if (C > 1) {
                int32_t vs = 0;
                for (int k = 0; k < C; ++k) vs += rec[lnk[r.link + k]].n;
                const float visit_sum = (float)vs;
                float best = 0.f;
                for (int k = 0; k < C; ++k) {
                    const Rec& ch = rec[lnk[r.link + k]];
                    const float visit = (float)ch.n;
                    const float q = ch.n > 0 ? ch.w / visit : 0.f;
                    const float prob = param.epsilon * 0.f + (1.f - param.epsilon) * ch.p;
                    const float u = prob * std::sqrt(visit_sum - visit) / (1.f + visit);
                    const float sc = (turn < 0 ? -q : q) + param.c_puct * u;
                    if (k == 0 || sc > best) best = sc, pick = k;
                }
            }
            parent_link = r.link, parent_board = b;
            const int cell = em[pick];
            b[cell / BOARD_SIZE][cell % BOARD_SIZE] = turn < 0 ? Player::Black : Player::White;
            key ^= zkey(turn, cell);
            stones += 1;
            cells.push_back(cell);
            cur = lnk[r.link + pick];
        }
        const int depth = (int)path.size();
        float v;
        if (stones == BOARD_CAPACITY) {
            // full-board leaf (:338-348): value by winner relative to the side to move, no visit++
            const Player winner = search(b);
            const int turn = turn_of(stones);
            const Player tp = turn < 0 ? Player::Black : Player::White;
            v = winner == tp ? (float)turn : (winner != Player::None ? -(float)turn : 0.f);
            // note: the reference leaves node.value untouched (0) on a draw
        } else {
            v = expand(b, key, stones, cur, depth, cells, parent_link, parent_board);
        }
        for (int i = 0; i < depth; ++i) {
            rec[path[i]].n += 1;
            if (i == depth - 1) rec[path[i]].w = rec[path[i]].w + v;
        }
    }

    float expand(const Board& b, uint64_t key, int stones, uint32_t leaf, int depth, const std::vector<int>& cells,
                 uint32_t parent_link, const Board& parent_board) {
        cnt.expansions += 1;
        const int turn = turn_of(stones);
        const Player tp = turn < 0 ? Player::Black : Player::White;
        float value;
        FBoard prob;
        evalbox.get_from(tp, b, value, prob);
        const bool sib_mode = depth >= 1 && turn_of(stones - 1) == turn;
        const int a_d = depth >= 1 ? cells.back() : -1;
        // earlier path stones of the mover's colour (S1 candidates)
        std::vector<int> others;
        for (int i = 0; i < depth; ++i) {
            const int stone_no = root_stones + i;  // stones on the board before path stone i was placed
            if (turn_of(stone_no) == turn && !(sib_mode && i == depth - 1)) others.push_back(cells[i]);
        }
        if (depth >= 3) cnt.deep += 1;
        if (!others.empty()) cnt.s1_runs += 1;
        const bool s2 = orph.test(stones + 1);
        if (s2) cnt.s2_runs += 1;
        std::vector<int> em = empties(b);
        const uint32_t link_base = (uint32_t)lnk.size();
        lnk.resize(lnk.size() + em.size());
        for (size_t k = 0; k < em.size(); ++k) {
            const int c = em[k];
            uint32_t child = NIL;
            auto agree = [&](uint32_t found) {
                if (child != NIL && found != child) {
                    std::fprintf(stderr, "pool_model: two sources name different records for one child\n");
                    std::abort();
                }
                child = found;
            };
            if (sib_mode) {
                const uint32_t sib = lnk[parent_link + rank_below(parent_board, c)];
                if (rec[sib].n >= 1) {
                    agree(lnk[rec[sib].link + rank_below(b, a_d) - (c < a_d ? 1 : 0)]);
                    cnt.fast_hits += 1;
                }
            }
            if (child == NIL || check_all_sources)
                for (int sp : others) {
                    auto it = exph.find(key ^ zkey(turn, sp) ^ zkey(turn, c));
                    if (it != exph.end()) {
                        agree(lnk[rec[it->second].link + rank_below(b, sp) - (c < sp ? 1 : 0)]);
                        cnt.s1_hits += 1;
                    }
                }
            if ((child == NIL && s2) || check_all_sources) {
                auto it = exph.find(key ^ zkey(turn, c));
                if (it != exph.end()) {
                    if (!s2 && child == NIL) {
                        std::fprintf(stderr, "pool_model: unflagged orphan at %d stones\n", stones + 1);
                        std::abort();
                    }
                    agree(it->second);
                    if (s2) cnt.s2_hits += 1;
                }
            }
            if (child == NIL) {
                child = (uint32_t)rec.size();
                rec.push_back(Rec{0, 0.f, prob[c / BOARD_SIZE][c % BOARD_SIZE], NIL});
            } else {
                rec[child].p = prob[c / BOARD_SIZE][c % BOARD_SIZE];
            }
            lnk[link_base + k] = child;
        }
        rec[leaf].n += 1;
        rec[leaf].link = link_base;
        exph[key] = leaf;
        exp.push_back(ExpEntry{b, key, leaf, stones});
        return value;
    }

    static bool contains(const Board& big, const Board& small) {
        for (int r = 0; r < BOARD_SIZE; ++r)
            for (int c = 0; c < BOARD_SIZE; ++c)
                if (small[r][c] != Player::None && big[r][c] != small[r][c]) return false;
        return true;
    }

    // policy() + tail of next() + the collector
    std::optional<Pos> next() {
        for (int i = 0; i < param.num_simulation; ++i) search_once();
        const Rec r = rec[root_rec];
        last_child_visits.fill(-1);
        last_root_visit = r.n, last_root_qsum = r.w;
        if (r.n == 0) return std::nullopt;
        std::vector<int> em = empties(root);
        int32_t vs = 0;
        for (size_t k = 0; k < em.size(); ++k) vs += rec[lnk[r.link + k]].n;
        const float visit_sum = (float)vs;
        int best = 0;
        float bp = 0.f;
        for (size_t k = 0; k < em.size(); ++k) {
            const float visit = (float)rec[lnk[r.link + k]].n;
            last_child_visits[em[k]] = rec[lnk[r.link + k]].n;
            const float p = visit / (visit_sum - visit + 1.f);
            if (k == 0 || !(p < bp)) bp = p, best = (int)k;
        }
        const int cell = em[best], turn = turn_of(root_stones);
        root[cell / BOARD_SIZE][cell % BOARD_SIZE] = turn < 0 ? Player::Black : Player::White;
        root_key ^= zkey(turn, cell);
        root_stones += 1;
        root_rec = lnk[r.link + best];
        collect();
        return Pos(cell / BOARD_SIZE, cell % BOARD_SIZE);
    }

    void collect() {
        std::vector<ExpEntry> keep;
        for (auto& e : exp)
            if (contains(e.b, root)) keep.push_back(e);
        exp.swap(keep);
        exph.clear();
        for (auto& e : exp) exph[e.key] = e.rec;
        orph.reset();
        for (auto& e : exp) {
            if (e.stones == root_stones) continue;  // the root itself
            const int m = turn_of(e.stones - 1);    // colour of the stone placed last
            const Player mp = m < 0 ? Player::Black : Player::White;
            bool safe = false;
            for (int s = 0; s < BOARD_CAPACITY && !safe; ++s)
                if (e.b[s / BOARD_SIZE][s % BOARD_SIZE] == mp && root[s / BOARD_SIZE][s % BOARD_SIZE] == Player::None)
                    safe = exph.count(e.key ^ zkey(m, s)) != 0;
            if (!safe) orph.set(e.stones), cnt.orphans += 1;
        }
    }
};

struct Case {
    const char* name;
    Callback cb;
    int sims;
    float c_puct;
    int prefill;  // random stones placed (by both engines' game) before the search starts
    int max_moves;
};

int run_case(const Case& cs, int game, Counters& total) {
    HyperParameter hp;
    hp.num_simulation = cs.sims, hp.epsilon = 0.f, hp.c_puct = cs.c_puct;
    CppEval ev(cs.cb);
    AlphaZero az(&ev, hp);
    az.noise = NoiseMode::None, az.tie = TieBreak::LowestIndex;
    az.faithful = (game & 1) != 0;  // both cost models of the oracle: with and without its memory valve
    PoolModel pm(&ev, hp);
    Game g;
    // a legal prefix without a winner, identical for both sides
    uint64_t rs = splitmix64(0xC6ull * 977 + (uint64_t)game * 131 + (uint64_t)cs.prefill);
    for (int i = 0; i < cs.prefill;) {
        rs = splitmix64(rs);
        const int cell = (int)(rs % BOARD_CAPACITY);
        if (g.board[cell / BOARD_SIZE][cell % BOARD_SIZE] != Player::None) continue;
        Game t = g;
        t.set(Pos(cell / BOARD_SIZE, cell % BOARD_SIZE));
        if (t.is_game_end() != Player::None) continue;
        g = t;
        ++i;
    }
    pm.root = g.board, pm.root_key = zobrist_board(g.board), pm.root_stones = cs.prefill;
    int moves = 0;
    for (; cs.max_moves == 0 || moves < cs.max_moves; ++moves) {
        auto po = az.next(g);
        auto pmv = pm.next();
        if (po != pmv || az.last_root_visit != pm.last_root_visit ||
            std::memcmp(&az.last_root_qsum, &pm.last_root_qsum, 4) != 0 || az.last_child_visits != pm.last_child_visits) {
            std::fprintf(stderr, "MISMATCH case %s game %d move %d: oracle (%d,%d) N=%d, model (%d,%d) N=%d\n", cs.name, game, moves,
                         po ? po->first : -1, po ? po->second : -1, az.last_root_visit, pmv ? pmv->first : -1,
                         pmv ? pmv->second : -1, pm.last_root_visit);
            return 1;
        }
        if (!po) break;
        g.set(*po);
        if (g.is_game_end() != Player::None) break;
    }
    total.expansions += pm.cnt.expansions, total.fast_hits += pm.cnt.fast_hits, total.s1_runs += pm.cnt.s1_runs;
    total.s1_hits += pm.cnt.s1_hits, total.s2_runs += pm.cnt.s2_runs, total.s2_hits += pm.cnt.s2_hits;
    total.orphans += pm.cnt.orphans, total.deep += pm.cnt.deep;
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const int games = argc > 1 ? std::atoi(argv[1]) : 2;
    const bool quick = argc > 2;
    std::vector<Case> cases = {
        {"det_48", det_eval_callback, 48, 1.f, 0, 16},
        {"neg_48", neg_eval_callback, 48, 1.f, 0, 16},
        {"det_deep_lowc", det_eval_callback, 200, 0.05f, 0, 10},
        {"neg_deep_lowc", neg_eval_callback, 200, 0.05f, 3, 10},
        {"det_late", det_eval_callback, 120, 0.3f, 205, 0},
        {"neg_late", neg_eval_callback, 300, 0.2f, 214, 0},
        {"det_mid_odd", det_eval_callback, 150, 0.1f, 6, 12},
    };
    if (!quick) {
        cases.push_back({"det_800", det_eval_callback, 800, 1.f, 0, 6});
        cases.push_back({"neg_800_lowc", neg_eval_callback, 800, 0.1f, 2, 6});
        cases.push_back({"det_full_game_16", det_eval_callback, 16, 0.5f, 0, 0});
    }
    int bad = 0;
    for (auto& cs : cases) {
        Counters t;
        for (int gi = 0; gi < games; ++gi) bad += run_case(cs, gi, t);
        std::printf("%-18s expansions %8llu deep(>=3) %7llu  sibling hits %9llu  S1 runs %6llu hits %7llu  S2 runs %5llu hits %5llu  orphans flagged %4llu\n",
                    cs.name, (unsigned long long)t.expansions, (unsigned long long)t.deep, (unsigned long long)t.fast_hits,
                    (unsigned long long)t.s1_runs, (unsigned long long)t.s1_hits, (unsigned long long)t.s2_runs,
                    (unsigned long long)t.s2_hits, (unsigned long long)t.orphans);
    }
    std::printf(bad ? "pool_model: %d MISMATCHES\n" : "pool_model: all cases identical to the oracle\n", bad);
    return bad ? 1 : 0;
}
