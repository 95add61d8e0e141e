// The bit-packed win detection of the kernels (c6_board.cuh: win_check, win_check_fast / run6_all), compiled for the host,
// against a cell-by-cell scan of search()'s rule (LC/game/search_winner/mod.rs:201-246: the colour whose run of six ENDS
// first in row-major order).  Random boards of several densities plus boards built around a run that touches the board
// edge or sits next to the row boundary, where a shift that carried across a row would show.  Also the n-th-empty-cell pick
// (nth_set_bit, pick_empty_cell) against a bit-by-bit walk.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "c6_board.cuh"

static int scan(const int (&cell)[15][15]) {
    int best = 1 << 30, win = 0;
    const int dr[4] = {0, 1, 1, 1}, dc[4] = {1, 0, 1, -1};
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) {
            if (!cell[r][c]) continue;
            for (int d = 0; d < 4; ++d) {
                int k = 0, rr = r, cc = c;
                while (k < 6 && rr >= 0 && rr < 15 && cc >= 0 && cc < 15 && cell[rr][cc] == cell[r][c]) rr -= dr[d], cc -= dc[d], ++k;
                if (k == 6 && r * 16 + c < best) best = r * 16 + c, win = cell[r][c];
            }
        }
    return win;
}

static long check(const int (&cell)[15][15], long& wins) {
    uint32_t B[8] = {0, 0, 0, 0, 0, 0, 0, 0}, W[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int r = 0; r < 15; ++r)
        for (int c = 0; c < 15; ++c) {
            const uint32_t bit = (uint32_t)(r * 16 + c);
            if (cell[r][c] < 0) B[bit >> 5] |= 1u << (bit & 31);
            if (cell[r][c] > 0) W[bit >> 5] |= 1u << (bit & 31);
        }
    const int want = scan(cell);
    wins += want != 0;
    return (c6::win_check(B, W) != want) + (c6::win_check_fast(B, W) != want);
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 200000;
    std::mt19937 rng(0xC6);
    long bad = 0, wins = 0, boards = 0;
    int cell[15][15];
    for (int it = 0; it < n; ++it, ++boards) {  // random boards: empty fraction 1/3 .. 2/3
        const int empties = 1 + it % 3;
        for (auto& row : cell)
            for (int& v : row) {
                const int x = (int)(rng() % (uint32_t)(2 + empties));
                v = x == 0 ? -1 : (x == 1 ? 1 : 0);
            }
        bad += check(cell, wins);
    }
    // five in a row wrapped around a row end (must NOT win), and six touching each edge (must win), on a sparse background
    const int dr[4] = {0, 1, 1, 1}, dc[4] = {1, 0, 1, -1};
    for (int d = 0; d < 4; ++d)
        for (int r = 0; r < 15; ++r)
            for (int c = 0; c < 15; ++c)
                for (int len = 5; len <= 7; ++len)
                    for (int col = -1; col <= 1; col += 2, ++boards) {
                        for (auto& row : cell)
                            for (int& v : row) v = (rng() % 5 == 0) ? ((rng() & 1) ? 1 : -1) : 0;
                        // a run of `len` that is cut by the board edge keeps whatever part fits; the linear bit index continues
                        // into the neighbouring row, which is exactly what a careless shift would accept
                        for (int k = 0; k < len; ++k) {
                            const int lin = (r + k * dr[d]) * 15 + c + k * dc[d];  // cell-major wrap on purpose
                            if (lin >= 0 && lin < 225) cell[lin / 15][lin % 15] = col;
                        }
                        bad += check(cell, wins);
                    }
    // nth_set_bit / pick_empty_cell (the random-empty-cell pick of the playout and DefaultPolicy kernels) against a bit-by-bit walk
    long picks = 0;
    for (int it = 0; it < 400000; ++it) {
        uint32_t x = rng();
        if (it % 5 == 0) x &= rng();
        if (it % 7 == 0) x |= rng();
        if (it % 11 == 0) x = 1u << (rng() & 31);
        if (!x) continue;
        const uint32_t n = rng() % (uint32_t)__builtin_popcount(x);
        uint32_t want = 0, k = n;
        for (int b = 0; b < 32; ++b)
            if ((x >> b) & 1u) {
                if (k == 0) { want = (uint32_t)b; break; }
                --k;
            }
        bad += c6::nth_set_bit(x, n) != want;
        ++picks;
    }
    for (int it = 0; it < 100000; ++it) {
        uint32_t B[8], W[8];
        int order[256], empties = 0;
        for (int j = 0; j < 8; ++j) {
            const uint32_t a = rng() & (it % 3 ? rng() : 0xFFFFFFFFu), b = rng() & rng() & ~a;
            B[j] = a & c6::valid_mask(j), W[j] = b & c6::valid_mask(j);
        }
        for (int bit = 0; bit < 256; ++bit) {
            const int j = bit >> 5;
            if (((c6::valid_mask(j) >> (bit & 31)) & 1u) && !(((B[j] | W[j]) >> (bit & 31)) & 1u)) order[empties++] = bit;
        }
        if (!empties) continue;
        for (int rep = 0; rep < 4; ++rep, ++picks) {
            const uint32_t p = rep == 0 ? 0u : (rep == 1 ? (uint32_t)empties - 1u : rng() % (uint32_t)empties);
            bad += c6::pick_empty_cell(B, W, p) != (uint32_t)order[p];
        }
    }
    std::printf("boards=%ld wins=%ld picks=%ld bad=%ld\n", boards, wins, picks, bad);
    return bad != 0;
}
