"""From packed game records to training tuples — the consumer immediately after self-play.

The reference's replay buffers take `(winner, [(turn, board, (row, col)), ...])` and store one
`(winner, player, board, position)` sample per stone (AlphaZero/py_weighted/buffer.py:22-31,
AlphaZero/cpp_weighted/ReplayBuffer.hpp:22-29), through 912-byte `Path` objects.  The engine's records are
packed (240 B per game, one byte per stone); `training_tuples` expands them straight into the arrays a buffer
samples from, vectorised, without materialising `Path`s.  `as_game_results` gives the reference's nested
form for code written against `Buffer.push_game`.
"""
import numpy as np

RECORD_BYTES = 240
CELLS = 225


def _fields(records):
    rec = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, RECORD_BYTES)
    winner = rec[:, 0].view(np.int8).astype(np.int32)
    n_moves = rec[:, 1].astype(np.int32)
    moves = rec[:, 8:8 + 232]
    return winner, n_moves, moves


def stone_colours(max_len=CELLS):
    """Colour of stone number m (Black 1 stone, then 2 each: LC/game/game_impl/mod.rs:125-131,158-167)."""
    m = np.arange(max_len)
    return np.where(m == 0, -1, np.where(((m - 1) // 2) % 2 == 0, 1, -1)).astype(np.int32)


def training_tuples(records):
    """(winners[N], players[N], boards[N, 225] int8, positions[N]) with one row per stone of every game:
    `boards[i]` is the position BEFORE the stone (agent_impl/mod.rs:139-143), `players[i]` who placed it."""
    winner, n_moves, moves = _fields(records)
    total = int(n_moves.sum())
    colours = stone_colours()
    winners = np.repeat(winner, n_moves)
    game_of = np.repeat(np.arange(len(n_moves)), n_moves)
    start = np.cumsum(n_moves) - n_moves
    stone = np.arange(total) - np.repeat(start, n_moves)
    players = colours[stone]
    positions = moves[game_of, stone].astype(np.int32)
    boards = np.zeros((total, CELLS), dtype=np.int8)
    # board before stone s of a game = all earlier stones of that game: a running scatter per game
    order = np.argsort(stone, kind="stable")
    cur = np.zeros((len(n_moves), CELLS), dtype=np.int8)
    lo = 0
    counts = np.bincount(stone, minlength=1)
    for s, cnt in enumerate(counts):  # at most 225 vectorised steps, each over all games that are that long
        idx = order[lo:lo + cnt]
        g = game_of[idx]
        boards[idx] = cur[g]
        cur[g, positions[idx]] = colours[s]
        lo += cnt
    return winners, players, boards, positions


def as_game_results(records):
    """[(winner, [(turn, board[225] list, (row, col)), ...]), ...] — what pyconnect6.self_play returns."""
    winner, n_moves, moves = _fields(records)
    colours = stone_colours()
    out = []
    for g in range(len(n_moves)):
        board = [0] * CELLS
        path = []
        for s in range(int(n_moves[g])):
            cell = int(moves[g, s])
            path.append((int(colours[s]), list(board), (cell // 15, cell % 15)))
            board[cell] = int(colours[s])
        out.append((int(winner[g]), path))
    return out
