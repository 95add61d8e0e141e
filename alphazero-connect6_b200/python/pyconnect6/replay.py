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


class PackedBuffer:
    """Drop-in for the reference's replay `Buffer` (AlphaZero/py_weighted/buffer.py:4-44: same constructor, `push_game`,
    `sample`, `clear_half`, `__len__`, same retention rule: the newest `max_size` samples) that keeps samples packed —
    (game, stone) pairs over 240-byte records — and expands only what `sample` returns.  `push_records` takes the engine's
    packed records directly; `push_game` still accepts the nested `(winner, [(turn, board, (row, col)), ...])` form."""

    def __init__(self, max_size, board_size, num_sample):
        self.max_size, self.board_size, self.num_sample = max_size, board_size, num_sample
        self._games = []                       # (winner, moves uint8[n])
        self._index = np.zeros((0, 2), np.int64)  # live samples, oldest first: (game slot, stone number)

    def __len__(self):
        return len(self._index)

    def _append(self, winner, moves):
        slot = len(self._games)
        self._games.append((int(winner), np.asarray(moves, dtype=np.uint8)))
        new = np.stack([np.full(len(moves), slot, np.int64), np.arange(len(moves), dtype=np.int64)], axis=1)
        self._index = np.concatenate([self._index, new])
        if len(self._index) > self.max_size:
            self._index = self._index[-self.max_size:]

    def push_records(self, records):
        winner, n_moves, moves = _fields(records)
        for g in range(len(n_moves)):
            self._append(winner[g], moves[g, :n_moves[g]])

    def push_game(self, game_result):
        win, path = game_result
        self._append(win, [row * self.board_size + col for (_, _, (row, col)) in path])

    def _expand(self, slot, stone):
        winner, moves = self._games[slot]
        colours = stone_colours(max(len(moves), 1))
        board = [0] * CELLS
        for s in range(stone):
            board[int(moves[s])] = int(colours[s])
        return winner, int(colours[stone]), board, int(moves[stone])

    def contents(self):
        """Every live sample, oldest first, in the reference buffer's element format (winner, player, board, position)."""
        return [self._expand(int(g), int(s)) for g, s in self._index]

    def sample(self):
        index_set = np.random.choice(len(self._index), self.num_sample, replace=False)
        wins, players, boards, policies = [], [], [], []
        for i in index_set:
            win, player, board, pos = self._expand(int(self._index[i, 0]), int(self._index[i, 1]))
            wins.append(win), players.append(player), boards.append(board), policies.append(pos)
        return wins, players, boards, policies

    def clear_half(self):
        self._index = self._index[len(self._index) // 2:]


def training_tuples_device(records, device=0):
    """The tuples of `training_tuples` as float32 torch tensors in HBM — (winners[N], players[N], boards[N, 225], positions[N]),
    the layout of the reference trainer's `ReplayBuffer::sample` (cpp_weighted/ReplayBuffer.hpp:38-65) — expanded on the GPU
    from the packed records (uint8 array / bytes, or a CUDA uint8 tensor that never leaves the device)."""
    import torch
    from . import _capi
    L = _capi.product()
    dev = torch.device("cuda", device)
    if isinstance(records, torch.Tensor):
        rec = records.to(dev).contiguous().view(torch.uint8).reshape(-1, RECORD_BYTES)
    else:
        rec = torch.from_numpy(np.ascontiguousarray(records, dtype=np.uint8).reshape(-1, RECORD_BYTES).copy()).to(dev)
    n_games = rec.shape[0]
    n_moves = rec[:, 1].to(torch.int64)
    offsets = (torch.cumsum(n_moves, 0) - n_moves).contiguous()
    rows = int(n_moves.sum().item())
    winners = torch.empty(rows, dtype=torch.float32, device=dev)
    players = torch.empty_like(winners)
    positions = torch.empty_like(winners)
    boards = torch.empty((rows, CELLS), dtype=torch.float32, device=dev)
    if n_games and rows:
        stream = torch.cuda.current_stream(dev).cuda_stream
        if L.c6_expand_records_device(rec.data_ptr(), offsets.data_ptr(), n_games, winners.data_ptr(), players.data_ptr(), boards.data_ptr(),
                                      positions.data_ptr(), stream) != 0:
            raise RuntimeError(L.c6_last_error().decode())
    return winners, players, boards, positions
