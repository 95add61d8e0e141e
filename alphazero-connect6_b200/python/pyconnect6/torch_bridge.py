"""Zero-copy evaluator bridge: a PyTorch module consumes the engine's leaf batch where it lies in HBM.

The reference's trainer callback (AlphaZero/cpp_weighted/main.cpp:16-37) receives host pointers, copies
them to the device, runs the net and copies the results back — per leaf of 8 boards.  Here the engine hands
the evaluator DEVICE pointers for a whole wave of leaves (`c6_device_callback`, include/connect6_b200.h) and
the CUDA stream its own kernels are ordered on; the bridge wraps them as torch tensors (no copy) and runs
the module on that stream, so nothing synchronises and nothing crosses PCIe.

    from pyconnect6 import torch_bridge
    net = torch_bridge.WeightedPolicy(num_block=10).cuda().eval()
    records, stats = torch_bridge.self_play_torch(torch_bridge.module_evaluator(net), n_games=8192, num_simulation=800)
"""
import ctypes as C
import os

import numpy as np

from . import _capi

CELLS = 225


class _DeviceBuffer:
    """Minimal __cuda_array_interface__ holder over a raw device pointer (torch.as_tensor adopts it without a copy)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


def device_tensor(ptr, shape, device):
    import torch
    return torch.as_tensor(_DeviceBuffer(ptr, shape), device=device)


def module_evaluator(module):
    """evaluate(player, boards[len, 15, 15]) -> (values[len], policies[len, 225]) from a module with the
    reference's forward(player[len], board[len, 225]) -> (policy, value) signature (cpp_weighted/model.hpp:52-95)."""
    import torch

    @torch.no_grad()
    def evaluate(player, boards):
        n = boards.shape[0]
        p = torch.full((n,), float(player), device=boards.device, dtype=boards.dtype)
        policy, value = module(p, boards.reshape(n, CELLS))
        return value.reshape(n), policy.reshape(n, CELLS)

    return evaluate


def self_play_torch(evaluate, n_games, num_simulation=800, epsilon=0.25, dirichlet_alpha=0.03, c_puct=1.0, seed=None, device=0,
                    slots=None, log2_records=18, max_moves=0, first_id=0, chunk=None):
    """`n_games` of AlphaZero self-play on `device` with a torch evaluator fed through device pointers.
    `chunk` bounds the boards handed to `evaluate` per call (None: one call per side to move and pipeline group).
    Returns ((n_games, 240) uint8 packed records, Stats)."""
    import torch
    L = _capi.product()
    dev = torch.device("cuda", device)
    err = []
    handle = [None]

    def cb(player, d_values, d_policies, length, stream, user):
        try:
            with torch.cuda.device(dev), torch.cuda.stream(torch.cuda.ExternalStream(stream, device=dev)):
                step = chunk or length
                for lo in range(0, length, step):
                    n = min(step, length - lo)
                    pol = device_tensor(d_policies + lo * CELLS * 4, (n, CELLS), dev)
                    val = device_tensor(d_values + lo * 4, (n,), dev)
                    v, p = evaluate(int(player), pol.view(n, 15, 15))
                    val.copy_(v.reshape(n).to(torch.float32))
                    pol.copy_(p.reshape(n, CELLS).to(torch.float32))  # stream-ordered after every read of the boards
        except BaseException as e:  # noqa: BLE001 - must not unwind through the C frames
            err.append(e)
            if handle[0]:
                L.c6_engine_cancel(handle[0])

    ccb = _capi.DEVICE_CALLBACK(cb)
    if seed is None:
        seed = int.from_bytes(os.urandom(4), "little")
    cfg = _capi.make_config(num_simulation=num_simulation, epsilon=epsilon, dirichlet_alpha=dirichlet_alpha, c_puct=c_puct, seed=seed,
                            device=device, slots=min(slots or n_games, n_games), log2_records=log2_records,
                            evaluator=_capi.EVAL_DEVICE_CALLBACK, max_moves=max_moves)
    e = L.c6_engine_create(C.byref(cfg))
    if not e:
        raise RuntimeError(L.c6_last_error().decode())
    handle[0] = e
    try:
        if L.c6_engine_start(e, first_id, n_games) != 0:
            raise RuntimeError(L.c6_last_error().decode())
        rc = L.c6_engine_run_device_callback(e, ccb, None)
        if err:
            raise RuntimeError("evaluator failed") from err[0]
        if rc != 0:
            raise RuntimeError(L.c6_last_error().decode())
        rec = (_capi.Record * n_games)()
        st = _capi.Stats()
        if L.c6_engine_records(e, rec, n_games) != 0 or L.c6_engine_stats(e, C.byref(st)) != 0:
            raise RuntimeError(L.c6_last_error().decode())
    finally:
        handle[0] = None
        L.c6_engine_destroy(e)
    return np.frombuffer(rec, dtype=np.uint8).reshape(n_games, C.sizeof(_capi.Record)).copy(), st


def WeightedPolicy(num_block=10, channels=256):
    """The policy/value net of the reference's libtorch trainer (AlphaZero/cpp_weighted/model.hpp:9-95) as a
    torch.nn.Module with the same forward(player, board) -> (policy, value) contract: 2 input planes (player,
    board), 3x3 conv + BN + ReLU stem, num_block - 1 residual blocks, policy head 1x1 conv(2) + BN + ReLU + FC +
    softmax, value head 1x1 conv(1) + BN(225) + ReLU + FC(256) + ReLU + FC(1) + tanh.  Random-init here: it
    stands for the user's evaluator in BASELINE.json configs[4]."""
    import torch
    from torch import nn

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv0 = nn.Conv2d(channels, channels, 3, padding=1)
            self.norm0 = nn.BatchNorm2d(channels)
            self.conv1 = nn.Conv2d(channels, channels, 3, padding=1)
            self.norm1 = nn.BatchNorm2d(channels)

        def forward(self, x):
            y = torch.relu(self.norm0(self.conv0(x)))
            return torch.relu(self.norm1(self.conv1(y)) + x)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.Conv2d(2, channels, 3, padding=1)
            self.norm = nn.BatchNorm2d(channels)
            self.tower = nn.Sequential(*[Block() for _ in range(num_block - 1)])
            self.policy_conv = nn.Conv2d(channels, 2, 1)
            self.policy_norm = nn.BatchNorm2d(2)
            self.policy_linear = nn.Linear(2 * CELLS, CELLS)
            self.value_conv = nn.Conv2d(channels, 1, 1)
            self.value_norm = nn.BatchNorm1d(CELLS)
            self.value_linear1 = nn.Linear(CELLS, 256)
            self.value_linear2 = nn.Linear(256, 1)

        def forward(self, player, board):
            n = board.shape[0]
            planes = torch.stack([player.reshape(n, 1).expand(n, CELLS), board.reshape(n, CELLS)], dim=1).reshape(n, 2, 15, 15)
            x = self.tower(torch.relu(self.norm(self.conv(planes))))
            p = torch.relu(self.policy_norm(self.policy_conv(x))).reshape(n, 2 * CELLS)
            p = torch.softmax(self.policy_linear(p), dim=1)
            v = torch.relu(self.value_norm(self.value_conv(x).reshape(n, CELLS)))
            v = torch.tanh(self.value_linear2(torch.relu(self.value_linear1(v))))
            return p, v

    return Net()
