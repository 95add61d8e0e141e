// Host-only pieces of the mirror (connect6_host.hpp) that need no GPU: RandomPolicy under Agent::play — a port of
// LC/policy/random_policy/tests.rs:7-25 (the game ends without an error, and a winner, if any, made the last move) run over
// many seeds — MultiPolicy dispatch by turn, and Game::set's error strings.
#include <cstdio>
#include <set>

#include "connect6_host.hpp"

using namespace connect6;

int main() {
    int bad = 0, wins = 0, draws = 0;
    std::set<std::vector<Pos>> distinct;
    for (uint64_t seed = 0; seed < 300; ++seed) {
        RandomPolicy policy(seed);
        PlayResult result;
        const std::string err = Agent(policy).play(result);
        if (!err.empty()) { std::printf("seed %llu: %s\n", (unsigned long long)seed, err.c_str()); ++bad; continue; }
        if (result.path.empty() || result.path.size() > BOARD_CAPACITY) ++bad;
        if (result.winner != Player::None) {
            ++wins;
            if (result.path.back().turn != result.winner) ++bad;  // tests.rs:19-23
        } else {
            ++draws;
            if (result.path.size() != BOARD_CAPACITY) ++bad;  // no winner: the board is full
        }
        // turn rule along the path: Black one stone, then two each (game_impl/mod.rs:125-131)
        for (size_t i = 0; i < result.path.size(); ++i) {
            const Player want = i == 0 ? Player::Black : ((((i - 1) / 2) % 2 == 0) ? Player::White : Player::Black);
            if (result.path[i].turn != want) { ++bad; break; }
        }
        std::vector<Pos> moves;
        for (const Path& p : result.path) moves.push_back(p.pos);
        distinct.insert(moves);
    }
    if (distinct.size() < 299) ++bad;  // different seeds, different games
    {  // OS-seeded default constructor plays a whole game too
        RandomPolicy policy;
        PlayResult result;
        if (!Agent(policy).play(result).empty() || result.path.empty()) ++bad;
    }
    {  // MultiPolicy: Black's policy is asked on Black's turns only
        struct Fixed : Policy {
            Pos p; int calls = 0;
            explicit Fixed(Pos q) : p(q) {}
            std::optional<Pos> next(const Game&) override { ++calls; return p; }
        } black(Pos(0, 0)), white(Pos(1, 1));
        MultiPolicy both(black, white);
        Game g;
        if (*both.next(g) != Pos(0, 0)) ++bad;
        if (!g.set(Pos(0, 0)).empty()) ++bad;
        if (*both.next(g) != Pos(1, 1) || black.calls != 1 || white.calls != 1) ++bad;
        if (g.set(Pos(0, 0)) != "already set position" || g.set(Pos(15, 0)) != "invalid position") ++bad;
    }
    std::printf("games=300 wins=%d draws=%d distinct=%zu bad=%d\n", wins, draws, distinct.size(), bad);
    return bad != 0;
}
