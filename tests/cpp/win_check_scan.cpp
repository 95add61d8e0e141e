// The bit-packed win detection of the kernels (c6_board.cuh: win_check, win_check_fast / run6_all), compiled for the host,
// against a cell-by-cell scan of search()'s rule (LC/game/search_winner/mod.rs:201-246: the colour whose run of six ENDS
// first in row-major order).  Random boards of several densities plus boards built around a run that touches the board
// edge or sits next to the row boundary, where a shift that carried across a row would show.
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
    std::printf("boards=%ld wins=%ld bad=%ld\n", boards, wins, bad);
    return bad != 0;
}
