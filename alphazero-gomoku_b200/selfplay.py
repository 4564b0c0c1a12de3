"""Batched self-play on one B200: G concurrent games, search and leaf evaluation on the device.

``SelfPlay.step()`` is one ply of ``play_game_and_collect`` (train.py:360-412) for all G games:
``MCTS.run`` (FILL -> EVAL -> COMMIT rounds of up to 32 leaves per game), temperature sampling,
``do_move``, and - for games that ended - outcome labelling, 8-fold symmetry expansion into the
example buffer and a restart from the empty board with a cleared tree (train.py:674-694).
The only host involvement is one 16-byte read per round (leaf count / games still running).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib, ptr
from .engine import POS_WORDS, SearchEngine
from .nn_engine import NetEngine

ROW = 901          # floats per example: planes 3*225, pi 225, z
TRUNK_FLOPS = {64: 2 * 225 * 9 * 64 * 64, 128: 2 * 225 * 9 * 128 * 128}   # per position per 3x3 layer


class SelfPlay:
    def __init__(self, model, rule: int = 0, n_games: int = 2048, n_sims: int = 800, cpuct: float = 1.0,
                 queue_len: int = 32, node_capacity: int = 8192, noise: bool = True, alpha: float = 0.05,
                 eps: float = 0.15, noise_plies: int = 10, temp_threshold: float = 10.0, max_moves: int = 225,
                 use_symmetries: bool = True, example_capacity: int = 1 << 20, seed: int = 12345, device="cuda:0",
                 game_base: int = 0):
        self.device = torch.device(device)
        self.G, self.n_sims, self.temp_threshold, self.max_moves = n_games, n_sims, float(temp_threshold), max_moves
        self.use_symmetries = use_symmetries
        self.engine = SearchEngine(rule, n_games, cpuct=cpuct, queue_len=queue_len, node_capacity=node_capacity,
                                   noise=noise, alpha=alpha, eps=eps, noise_plies=noise_plies, seed=seed, device=device,
                                   game_base=game_base)
        net = model.net if hasattr(model, "net") else model
        self.net = NetEngine(len(net.res_blocks), net.channels, self.device, max_batch=n_games * queue_len)
        self.net.load_state_dict(net.state_dict())
        self.channels, self.n_layers = net.channels, 2 * len(net.res_blocks)
        check(lib.azg_selfplay_enable(self.engine._h, max_moves))
        dev = self.device
        self.probs = torch.empty((n_games * queue_len, 225), dtype=torch.float32, device=dev)
        self.noise = torch.zeros((n_games, 225), dtype=torch.float64, device=dev) if noise else None
        self.actions = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.done = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.winners = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.capacity = example_capacity
        self.examples = torch.empty((example_capacity, ROW), dtype=torch.float32, device=dev) if example_capacity else None
        empty = torch.zeros((n_games, POS_WORDS), dtype=torch.int32, device=dev)
        empty[:, 16] = 1          # player 1 to move
        empty[:, 17] = -1         # no last move
        self.empty_roots = empty
        self.engine.set_roots(empty, clear_tree=True)
        self.draw = 0
        # counters
        self.total_sims = 0
        self.total_evals = 0
        self.total_rounds = 0
        self.total_launches = 0
        self.games_finished = 0
        self.last_pi = None

    def close(self):
        self.engine.close()
        self.net.close()

    # ------------------------------------------------------------------ one ply for every game
    def search(self):
        """MCTS.run for all games; returns (pi, visits) on the device."""
        eng = self.engine
        eng.begin(self.n_sims)
        launches = 1
        if self.noise is not None:
            eng._sync_stream()
            check(lib.azg_selfplay_noise(eng._h, self.draw, ptr(self.noise)))
            launches += 1
        while True:
            n_leaves, n_more, _ = eng.fill()
            launches += 2
            if n_leaves > 0:
                self.net.forward_leaves(eng, self.probs)
                eng.commit(self.probs, self.noise)
                launches += 1 + self.n_layers + 2 + 1           # stem, 3x3 layers, two head kernels, commit
                self.total_evals += n_leaves
                self.total_rounds += 1
            if n_more == 0:
                break
        self.total_sims += self.G * self.n_sims
        self.total_launches += launches + 1
        return eng.result()

    def step(self):
        eng = self.engine
        pi, visits = self.search()
        self.last_pi = pi
        self.draw += 1
        check(lib.azg_selfplay_choose(eng._h, ptr(pi), C.c_float(self.temp_threshold), self.draw, ptr(self.actions)))
        status = eng.advance(self.actions, gc=True, reserve=self.n_sims + self.n_sims // self.engine.queue_len + 8)
        check(lib.azg_selfplay_finish(eng._h, ptr(status), self.max_moves, int(self.use_symmetries), ptr(self.examples),
                                      self.capacity, ptr(self.cursor), ptr(self.done), ptr(self.winners)))
        eng.set_roots(self.empty_roots, mask=self.done, clear_tree=True)
        self.total_launches += 4
        return status

    # ------------------------------------------------------------------ examples
    def n_examples(self) -> int:
        return min(int(self.cursor.item()), self.capacity)

    def drain_examples(self) -> torch.Tensor:
        """Rows produced since the last drain, float32[n, 901] (planes 675, pi 225, z) on the device."""
        n = self.n_examples()
        out = self.examples[:n].clone()
        self.cursor.zero_()
        return out

    @staticmethod
    def split(rows: torch.Tensor):
        """-> (states [n,3,15,15], pis [n,225], zs [n,1]) as ReplayBuffer.sample returns them (train.py:287-293)."""
        return rows[:, :675].reshape(-1, 3, 15, 15), rows[:, 675:900], rows[:, 900:901]
