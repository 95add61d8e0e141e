// ORACLE — TEST INFRASTRUCTURE ONLY (see c6_rng.hpp header note).
//
// CPU restatement of the reference's pure-MCTS policy, function by function:
//   /root/reference/Connect6/libconnect6/src/policy/default_policy/mod.rs:24-271   (Node, DefaultPolicy)
// Data structures as in the reference: a board-hash-keyed map of {visit, black_win, board, next_node},
// never pruned; `expand` INSERTS (mod.rs:174 — an existing node of that board is replaced by a fresh one,
// and the parent's `next_node` may list the same child more than once, :181); `select` is
// `max_by(partial_cmp)` (the LAST maximal child, :146-150) and stops unless that maximum is non-zero (:156).
//
// Parity unpinned at the `rand` boundary: the reference draws the expanded cell (:165), every rollout stone
// (:193) and the fallback move (:254) from the unseedable `thread_rng()`, and none of its tests pins a value.
// The seam that replaces them is a counter-based Philox4x32-10 stream, identical in the GPU engine:
//   key (seed ^ "DPOL", game_id);  m = stones on the board when next() was called, i = search number
//   expand   : counter (m, i, 0xFFFFFFFF, 0) word 0
//   rollout  : counter (m, i, n >> 2, 0) word n & 3 for rollout stone n
//   fallback : counter (m, 0xFFFFFFFF, 0, 0) word 0
// and `choose` on a list of n items = item (u32 * n) >> 32 (the list is the row-major empty cells in all
// three places: Simulate::from_game / deep_clone build it row-major and simulate_in removes order-preservingly).
#pragma once
#include "c6_oracle.hpp"

namespace c6o {

constexpr uint32_t DOMAIN_DEFPOL = 0x44504F4Cu;  // "DPOL"

struct DpNode {  // mod.rs:24-43
    int32_t visit = 0;
    int32_t black_win = 0;
    Board board;
    std::vector<uint64_t> next_node;
};

struct DpStats {
    uint64_t searches = 0, rollouts = 0, rollout_stones = 0, select_steps = 0, children_scanned = 0, replaced = 0;
};

struct DefaultPolicy : Policy {
    int num_iter = 50;  // mod.rs:104
    std::unordered_map<uint64_t, DpNode> map;
    uint32_t k0 = DOMAIN_DEFPOL, k1 = 0;
    uint32_t cur_m = 0, cur_i = 0;
    DpStats stats;

    DefaultPolicy() {}
    DefaultPolicy(int n, uint32_t seed, uint32_t game_id) : num_iter(n), k0(seed ^ DOMAIN_DEFPOL), k1(game_id) {}

    static uint64_t hash(const Board& b) { return siphash13_board(b); }  // mod.rs:55-59
    static size_t pick(uint32_t u, size_t n) { return (size_t)(((uint64_t)u * (uint64_t)n) >> 32); }

    void init(const Simulate& sim) {  // :121-124
        DpNode n;
        n.board = sim.board;
        map.emplace(hash(sim.board), std::move(n));  // entry().or_insert
    }

    static float prob(const DpNode& n, Player turn) {  // :135-140
        float x = (float)n.black_win / (1.f + (float)n.visit);
        return turn == Player::Black ? x : -x;
    }

    std::optional<Pos> select(const Simulate& sim) {  // :129-161
        const DpNode& tree_node = map.at(hash(sim.board));
        const DpNode* best = nullptr;
        float bp = 0.f;
        stats.select_steps += 1;
        for (uint64_t h : tree_node.next_node) {  // max_by: a later element wins ties
            const DpNode& c = map.at(h);
            float p = prob(c, sim.turn);
            stats.children_scanned += 1;
            if (!best || !(p < bp)) best = &c, bp = p;
        }
        if (best && bp != 0.f) return diff_board(best->board, sim.board);
        return std::nullopt;
    }

    Pos expand(const Simulate& sim) {  // :164-184
        const uint32_t u = philox4x32_10(cur_m, cur_i, 0xFFFFFFFFu, 0, k0, k1).v[0];
        Pos p = sim.possible[pick(u, sim.possible.size())];
        Simulate child = sim;
        child.simulate_in(p.first, p.second);
        const uint64_t hb = hash(child.board);
        DpNode fresh;
        fresh.board = child.board;
        if (map.count(hb)) stats.replaced += 1;
        map[hb] = std::move(fresh);  // HashMap::insert: replaces
        map.at(hash(sim.board)).next_node.push_back(hb);
        return p;
    }

    void update(const Simulate& sim, const std::vector<Pos>& path) {  // :190-226
        Simulate ro = sim.deep_clone();
        uint32_t n = 0;
        Philox4 blk{};
        stats.rollouts += 1;
        while (search(ro.board) == Player::None) {
            if (ro.possible.empty()) break;
            if ((n & 3) == 0) blk = philox4x32_10(cur_m, cur_i, n >> 2, 0, k0, k1);
            Pos p = ro.possible[pick(blk.v[n & 3], ro.possible.size())];
            ro.simulate_in(p.first, p.second);
            ++n;
        }
        stats.rollout_stones += n;
        const Player win = search(ro.board);
        if (win == Player::None) return;
        const int32_t bw = win == Player::Black ? 1 : 0;
        Simulate s = sim.deep_clone();
        auto upd = [&](const Simulate& x) {
            DpNode& node = map.at(hash(x.board));
            node.visit += 1;
            node.black_win += bw;
        };
        upd(s);
        for (auto it = path.rbegin(); it != path.rend(); ++it) {
            s.rollback_in(it->first, it->second);
            upd(s);
        }
    }

    void search_once(const Game& game) {  // :229-249
        Simulate sim = Simulate::from_game(game);
        init(sim);
        std::vector<Pos> path;
        while (auto pos = select(sim)) {
            path.push_back(*pos);
            sim.simulate_in(pos->first, pos->second);
        }
        stats.searches += 1;
        if (search(sim.board) != Player::None) return;
        // a descent that ends on a FULL board without a winner: the reference's expand() panics (`choose().unwrap()` on an empty
        // list, :165); restated as "this search changes nothing" — the engine does the same
        if (sim.possible.empty()) return;
        Pos p = expand(sim);
        path.push_back(p);
        sim.simulate_in(p.first, p.second);
        update(sim, path);
    }

    std::optional<Pos> policy(const Simulate& sim) {  // :252-261
        if (auto pos = select(sim)) return pos;
        // `.choose().unwrap()`: a full board panics in the reference; Agent::play never asks then unless the
        // game reached 225 stones without a winner
        if (sim.possible.empty()) std::abort();
        const uint32_t u = philox4x32_10(cur_m, 0xFFFFFFFFu, 0, 0, k0, k1).v[0];
        return sim.possible[pick(u, sim.possible.size())];
    }

    std::optional<Pos> next(const Game& game) override {  // :264-271
        uint32_t stones = 0;
        for (auto& row : game.board)
            for (auto c : row) stones += c != Player::None;
        cur_m = stones;
        // the reference panics here on a full board without a winner (`choose().unwrap()` on an empty list,
        // :165); Agent::play only gets there after 225 stones and no six — restated as "no move" (winner None)
        if (stones == BOARD_CAPACITY) return std::nullopt;
        for (cur_i = 0; (int)cur_i < num_iter; ++cur_i) search_once(game);
        return policy(Simulate::from_game(game));
    }
};

}  // namespace c6o
