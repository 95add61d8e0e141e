// Host-side mirror of the reference's Rust API for the pieces that have no search in them
// (names, argument meaning and error strings follow the reference):
//   Player                 LC/game/player/mod.rs:8-61
//   Game, SetResult        LC/game/game_impl/mod.rs:26-253
//   Policy (trait)         LC/policy/mod.rs:45-48
//   CppPolicy              CB/libcppconnect6/src/cppbind/cpp_policy/mod.rs:8-43
//   Agent, Path, PlayResult LC/agent/agent_impl/mod.rs:34-178
// (LC = Connect6/libconnect6/src, CB = Connect6/cppbind).  Win detection uses the same bit-packed
// shift/AND routine as the kernels (c6_device.cuh: win_check), compiled for the host.
#pragma once
#include <array>
#include <cstdio>
#include <optional>
#include <random>
#include <string>
#include <utility>
#include <vector>

#include "c6_board.cuh"

namespace connect6 {

constexpr size_t BOARD_SIZE = 15;
constexpr size_t BOARD_CAPACITY = BOARD_SIZE * BOARD_SIZE;

enum class Player : int { Black = -1, None = 0, White = 1 };
inline Player switch_player(Player p) {
    return p == Player::Black ? Player::White : (p == Player::White ? Player::Black : Player::None);
}
inline Player player_from(int v) { return v == -1 ? Player::Black : (v == 1 ? Player::White : Player::None); }

using Board = std::array<std::array<Player, BOARD_SIZE>, BOARD_SIZE>;
using Pos = std::pair<size_t, size_t>;

// search() on a cell-array board
inline Player search(const Board& b) {
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0}, W[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t r = 0; r < BOARD_SIZE; ++r)
        for (size_t c = 0; c < BOARD_SIZE; ++c) {
            uint32_t bit = (uint32_t)(r * 16 + c);
            if (b[r][c] == Player::Black) B[bit >> 5] |= 1u << (bit & 31);
            if (b[r][c] == Player::White) W[bit >> 5] |= 1u << (bit & 31);
        }
    return player_from(c6::win_check(B, W));
}

struct SetResult {
    Player player = Player::None;
    int num_remain = 0;
    Pos position{0, 0};
};

class Game {
   public:
    Game() {
        for (auto& r : board_) r.fill(Player::None);
    }
    // "" on success, else Error::description() of the reference ("invalid position" / "already set position")
    std::string set(Pos pos, SetResult* out = nullptr) {
        size_t row = pos.first, col = pos.second;
        if (row >= BOARD_SIZE || col >= BOARD_SIZE) return "invalid position";
        if (board_[row][col] != Player::None) return "already set position";
        board_[row][col] = turn_;
        num_remain_ -= 1;
        if (out) *out = SetResult{turn_, num_remain_, pos};
        if (num_remain_ <= 0) {
            num_remain_ = 2;
            turn_ = switch_player(turn_);
        }
        return "";
    }
    const Board& get_board() const { return board_; }
    Player get_turn() const { return turn_; }
    int get_remain() const { return num_remain_; }
    Player is_game_end() const { return search(board_); }
    void print(FILE* f) const {
        std::string s = "0 A B C D E F G H I J K L M N O\n";
        for (size_t i = 0; i < BOARD_SIZE; ++i) {
            s.push_back((char)(0x61 + i));
            s.push_back(' ');
            for (size_t j = 0; j < BOARD_SIZE; ++j)
                s += board_[i][j] == Player::Black ? "X " : (board_[i][j] == Player::White ? "O " : "_ ");
            s.push_back('\n');
        }
        std::fputs(s.c_str(), f);
    }

   private:
    Player turn_ = Player::Black;
    int num_remain_ = 1;
    Board board_;
};

struct Policy {
    virtual ~Policy() {}
    virtual std::optional<Pos> next(const Game& game) = 0;
};

using PolicyCallback = void (*)(float* board, int* position);
class CppPolicy : public Policy {
   public:
    explicit CppPolicy(PolicyCallback cb) : cb_(cb) {}
    std::optional<Pos> callback(const Board& board) const {
        int res[2] = {-1, -1};
        float bf[BOARD_CAPACITY];
        for (size_t i = 0; i < BOARD_SIZE; ++i)
            for (size_t j = 0; j < BOARD_SIZE; ++j) bf[i * BOARD_SIZE + j] = (float)(int)board[i][j];
        cb_(bf, res);
        if (res[0] == -1) return std::nullopt;
        return Pos((size_t)res[0], (size_t)res[1]);
    }
    std::optional<Pos> next(const Game& game) override { return callback(game.get_board()); }

   private:
    PolicyCallback cb_;
};

// LC/policy/multi_policy/mod.rs:61-71 — one policy per colour
class MultiPolicy : public Policy {
   public:
    MultiPolicy(Policy& black, Policy& white) : black_(black), white_(white) {}
    std::optional<Pos> next(const Game& game) override {
        return game.get_turn() == Player::Black ? black_.next(game) : white_.next(game);
    }

   private:
    Policy &black_, &white_;
};

// LC/policy/io_policy/mod.rs:68-107 — a move typed as "aB" (row 'a'.., column 'A'..): up to 10 bytes per read, the
// alphabetic characters must be exactly two and in range, else "invalid input, retry".  End of input returns
// no move (the reference would spin on an exhausted reader).
class IoPolicy : public Policy {
   public:
    IoPolicy(FILE* in, FILE* out) : in_(in), out_(out) {}
    std::optional<Pos> next(const Game&) override {
        for (;;) {
            char buf[10];
            size_t got = 0;
            int c;
            while (got < sizeof buf && (c = std::fgetc(in_)) != EOF) {  // one line-buffered read() of at most 10 bytes
                buf[got++] = (char)c;
                if (c == '\n') break;
            }
            if (got == 0) return std::nullopt;
            char q[10];
            size_t n = 0;
            for (size_t i = 0; i < got; ++i)
                if ((buf[i] >= 'a' && buf[i] <= 'z') || (buf[i] >= 'A' && buf[i] <= 'Z')) q[n++] = buf[i];
            if (n == 2) {
                size_t row = (size_t)(unsigned char)q[0] - 0x61, col = (size_t)(unsigned char)q[1] - 0x41;  // wraps like usize
                if (row < BOARD_SIZE && col < BOARD_SIZE) return Pos(row, col);
            }
            std::fputs("invalid input, retry\n", out_);
            std::fflush(out_);
        }
    }

   private:
    FILE *in_, *out_;
};

// LC/policy/random_policy/mod.rs:31-49 — a uniformly chosen empty cell (`possible.choose(thread_rng())`), None on a full
// board.  Host-only, no search in it; seeded from the OS like thread_rng() unless a seed is given.
class RandomPolicy : public Policy {
   public:
    RandomPolicy() : rng_(std::random_device{}()) {}
    explicit RandomPolicy(uint64_t seed) : rng_(seed) {}
    std::optional<Pos> next(const Game& game) override {
        std::vector<Pos> possible;  // row-major, as Simulate::from_game builds it
        const Board& b = game.get_board();
        for (size_t r = 0; r < BOARD_SIZE; ++r)
            for (size_t c = 0; c < BOARD_SIZE; ++c)
                if (b[r][c] == Player::None) possible.emplace_back(r, c);
        if (possible.empty()) return std::nullopt;
        return possible[std::uniform_int_distribution<size_t>(0, possible.size() - 1)(rng_)];
    }

   private:
    std::mt19937_64 rng_;
};

struct Path {
    Player turn;
    Board board;
    Pos pos;
};
struct PlayResult {
    Player winner = Player::None;
    std::vector<Path> path;
};

class Agent {
   public:
    explicit Agent(Policy& policy, bool debug = false) : policy_(policy), debug_(debug) {}
    // returns "" (Ok) or the error description (Err); result in `out`
    std::string play(PlayResult& out) {
        Game& game = game_;
        out = PlayResult();
        for (;;) {
            if (debug_) game.print(stdout);
            auto pos = policy_.next(game);
            if (!pos) break;
            out.path.push_back(Path{game.get_turn(), game.get_board(), *pos});
            SetResult sr;
            std::string err = game.set(*pos, &sr);
            if (!err.empty()) return err;
            if (debug_)
                std::printf("%s (%c, %c), remain %d, 0.0 elapsed\n", sr.player == Player::Black ? "Black" : "White",
                            (char)(sr.position.first + 0x61), (char)(sr.position.second + 0x41), sr.num_remain);
            Player end = game.is_game_end();
            if (end != Player::None) {
                out.winner = end;
                break;
            }
        }
        if (debug_) game.print(stdout);
        return "";
    }

   private:
    Game game_;
    Policy& policy_;
    bool debug_;
};

}  // namespace connect6
