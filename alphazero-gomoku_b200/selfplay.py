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
PACKED_WORDS = 244  # uint32 per packed ply: stones 16, side to move, z, pi 225, pad (include/azgomoku_b200.h)


def expand_examples(packed: torch.Tensor, use_symmetries: bool = True) -> torch.Tensor:
    """Packed plies int32[n, 244] (device) -> example rows float32[n * 8, 901] in the reference's symmetry order
    (train.py:405-410): what ``play_game_and_collect`` appends for these plies."""
    n = int(packed.shape[0])
    k = 8 if use_symmetries else 1
    out = torch.empty((n * k, ROW), dtype=torch.float32, device=packed.device)
    if n:
        src = packed.contiguous()
        with torch.cuda.device(packed.device):
            check(lib.azg_examples_expand(ptr(src), n, int(use_symmetries), ptr(out), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def trunk_flops(channels: int) -> int:
    """Algorithmic FLOPs (2 x MAC) of one 3x3 trunk layer on one position (225 real pixels)."""
    return 2 * 225 * 9 * channels * channels


TRUNK_FLOPS = {c: trunk_flops(c) for c in (64, 128, 256)}


class SelfPlay:
    def __init__(self, model, rule: int = 0, n_games: int = 2048, n_sims: int = 800, cpuct: float = 1.0,
                 queue_len: int = 32, node_capacity: int = 8192, noise: bool = True, alpha: float = 0.05,
                 eps: float = 0.15, noise_plies: int = 10, temp_threshold: float = 10.0, max_moves: int = 225,
                 use_symmetries: bool = True, example_capacity: int = 1 << 20, seed: int = 12345, device="cuda:0",
                 game_base: int = 0, max_games: int | None = None, packed_examples: bool = False):
        """``max_games``: play exactly that many games to completion (train.py:671-694 plays
        ``games_per_iteration`` full games): finished slots restart only while fewer than ``max_games`` games
        have been started, afterwards they retire (``active`` mask) - no game is cut off or counted twice.
        ``packed_examples``: keep finished games as one 976-byte record per ply (stones, side, z, pi) and expand the
        8 symmetries only when rows are asked for (``drain_examples``) - ``drain_packed`` is what ranks exchange;
        ``example_capacity`` then counts plies."""
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
        # kernels per evaluated round: stem, one per 3x3 layer, head kernels (the 1x1 head convs are fused into the
        # last layer up to 128 channels), commit
        self._round_launches = 1 + self.n_layers + (1 if net.channels <= 128 and self.n_layers else 2) + 1
        check(lib.azg_selfplay_enable(self.engine._h, max_moves))
        dev = self.device
        self.probs = torch.empty((n_games * queue_len, 225), dtype=torch.float32, device=dev)
        self.noise = torch.zeros((n_games, 225), dtype=torch.float64, device=dev) if noise else None
        self.actions = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.done = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.winners = torch.empty(n_games, dtype=torch.int32, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        self.capacity = example_capacity
        self.packed = bool(packed_examples)
        if not example_capacity:
            self.examples = None
        elif self.packed:
            self.examples = torch.empty((example_capacity, PACKED_WORDS), dtype=torch.int32, device=dev)
        else:
            self.examples = torch.empty((example_capacity, ROW), dtype=torch.float32, device=dev)
        empty = torch.zeros((n_games, POS_WORDS), dtype=torch.int32, device=dev)
        empty[:, 16] = 1          # player 1 to move
        empty[:, 17] = -1         # no last move
        self.empty_roots = empty
        self.engine.set_roots(empty, clear_tree=True)
        self.draw = 0
        self.max_games = max_games
        if max_games is not None:
            self.active = (torch.arange(n_games, device=dev) < max_games).to(torch.int32)
            self.started = torch.tensor([min(n_games, max_games)], dtype=torch.int64, device=dev)
            check(lib.azg_selfplay_set_active(self.engine._h, ptr(self.active)))
        else:
            self.active = None
        # counters
        self.total_sims = 0
        self.total_evals = 0
        self.total_rounds = 0
        self.total_launches = 0
        self.games_finished = 0
        self.last_pi = None
        self._graph_rounds = 0

    def close(self):
        self.engine.close()
        self.net.close()

    # ------------------------------------------------------------------ one ply for every game
    def search(self):
        """MCTS.run for all games; returns (pi, visits) on the device."""
        self.search_begin()
        while self.search_round():
            pass
        return self.engine.result()

    # the run in its asynchronous pieces (PipelinedSelfPlay interleaves them for two game groups)
    def _restart(self):
        """Restart finished slots from the empty board with a cleared tree (train.py:674-694); with ``max_games``
        only while games remain to be started - the other finished slots retire.  Device-side, no host sync."""
        eng = self.engine
        if self.max_games is None:
            eng.set_roots(self.empty_roots, mask=self.done, clear_tree=True)
            return
        done = self.done != 0
        restart = done & ((self.started + torch.cumsum(self.done, 0)) <= self.max_games)
        self.started += restart.sum()
        self.active.copy_(((self.active != 0) & (~done | restart)).to(torch.int32))
        eng.set_roots(self.empty_roots, mask=restart.to(torch.int32), clear_tree=True)

    def games_running(self) -> int:
        """Slots still playing (host sync); always G without ``max_games``."""
        return self.G if self.active is None else int(self.active.sum().item())

    def search_begin(self):
        eng = self.engine
        eng.begin(self.n_sims, mask=self.active)
        self._launches = 1
        if self.noise is not None:
            eng._sync_stream()
            check(lib.azg_selfplay_noise(eng._h, self.draw, ptr(self.noise)))
            self._launches += 1
        eng.fill_async()
        self._launches += 2

    def search_round(self) -> bool:
        """Consume the outstanding fill: evaluate + commit its leaf batch and queue the next fill.
        Returns False when the run is complete."""
        eng = self.engine
        n_leaves, n_more, _ = eng.read_counters()
        if n_leaves > 0:
            self.net.forward_leaves(eng, self.probs)
            eng.commit(self.probs, self.noise)
            self._launches += self._round_launches
            self.total_evals += n_leaves
            self.total_rounds += 1
        if n_more == 0:
            self.total_sims += self.G * self.n_sims
            self.total_launches += self._launches + 1
            return False
        eng.fill_async()
        self._launches += 2
        return True

    def step(self):
        if self._graph_rounds:
            return self._step_graph()
        pi, visits = self.search()
        return self.finish_step(pi)

    # ------------------------------------------------------------------ whole ply as one CUDA graph
    def enable_graph(self, extra_rounds: int = 1):
        """Capture one full ply - begin, noise, a FIXED number of FILL -> leaf evaluation -> COMMIT rounds,
        result, move choice, advance, game-end handling, restart - into a CUDA graph.  A run of n
        simulations queues at most n + n/queue_len + 1 leaves, i.e. needs at most
        ceil(that / queue_len) + 1 rounds; rounds after a game (or all games) finished are no-ops
        (every kernel reads its work count from device memory), so no host synchronisation is left in
        the ply.  Results are identical to the host-driven loop."""
        q = self.engine.queue_len
        if q < 2:
            raise ValueError("graph mode needs queue_len >= 2 (with a queue of 1 every simulation is its own round)")
        # every flush parks one simulation that queues a second leaf, so n simulations queue up to n*q/(q-1) leaves
        leaves = -(-self.n_sims * q // (q - 1)) + 1
        self._graph_rounds = (leaves + q - 1) // q + extra_rounds
        self._graph = None
        self._graph_plies = 0

    def _ply_async(self):
        """Everything of one ply, without host synchronisation (capturable)."""
        eng = self.engine
        eng.begin(self.n_sims, mask=self.active)
        if self.noise is not None:
            eng._sync_stream()
            check(lib.azg_selfplay_noise(eng._h, self.draw, ptr(self.noise)))
        for _ in range(self._graph_rounds):
            eng.fill_async()
            self.net.forward_leaves(eng, self.probs)
            eng.commit(self.probs, self.noise)
        check(lib.azg_search_result(eng._h, ptr(self._pi), ptr(self._visits)))
        check(lib.azg_selfplay_choose(eng._h, ptr(self._pi), C.c_float(self.temp_threshold), self.draw, ptr(self.actions)))
        check(lib.azg_search_advance(eng._h, ptr(self.actions), 1, self._reserve(), ptr(self._status)))
        self._finish(self._status)
        self._restart()

    def _finish(self, status):
        eng = self.engine
        if self.packed:
            check(lib.azg_selfplay_finish_packed(eng._h, ptr(status), self.max_moves, ptr(self.examples), self.capacity,
                                                 ptr(self.cursor), ptr(self.done), ptr(self.winners)))
        else:
            check(lib.azg_selfplay_finish(eng._h, ptr(status), self.max_moves, int(self.use_symmetries), ptr(self.examples),
                                          self.capacity, ptr(self.cursor), ptr(self.done), ptr(self.winners)))

    def _reserve(self) -> int:
        """Free nodes a slab needs for another run: one per simulation plus one per flush (see enable_graph)."""
        q = max(self.engine.queue_len, 2)
        return -(-self.n_sims * q // (q - 1)) + 8

    GRAPH_CHECK_EVERY = 16

    def _check_graph_run(self):
        """The replayed graph never looks at the counters: every few plies make sure no run was truncated by the
        fixed round count and no game hit an engine error (full slab, depth overflow, no float64 prior slot)."""
        _, n_more, _ = self.engine.read_counters()
        st = self.engine.stats()
        if n_more != 0 or st["games_in_error"] != 0:
            raise RuntimeError(f"graph-mode self-play: {n_more} games still had simulations to run after "
                               f"{self._graph_rounds} rounds, {st['games_in_error']} games in error (bits {st['error_bits']:#x})")

    def _step_graph(self):
        """Captured once, replayed every ply (the RNG streams are keyed on device-side counters)."""
        if self._graph is None:
            self._pi = torch.empty((self.G, 225), dtype=torch.float32, device=self.device)
            self._visits = torch.empty((self.G, 225), dtype=torch.int32, device=self.device)
            self._status = torch.empty(self.G, dtype=torch.int32, device=self.device)
            g = torch.cuda.CUDAGraph()
            self.net.profile(False)
            with torch.cuda.graph(g):
                self.engine._sync_stream()
                self._ply_async()
            self._graph = g
        self._graph.replay()
        self._graph_plies += 1
        if self._graph_plies % self.GRAPH_CHECK_EVERY == 1:
            self._check_graph_run()
        self.last_pi = self._pi
        self.total_sims += self.G * self.n_sims
        self.total_rounds += self._graph_rounds
        self.total_launches += 2 + self._graph_rounds * (2 + self._round_launches) + 5
        return self._status

    def finish_step(self, pi):
        eng = self.engine
        self.last_pi = pi
        check(lib.azg_selfplay_choose(eng._h, ptr(pi), C.c_float(self.temp_threshold), self.draw, ptr(self.actions)))
        status = eng.advance(self.actions, gc=True, reserve=self._reserve())
        self._finish(status)
        self._restart()
        self.total_launches += 4
        return status

    # ------------------------------------------------------------------ examples
    def n_examples(self) -> int:
        return min(int(self.cursor.item()), self.capacity)

    def drain_packed(self) -> torch.Tensor:
        """Packed plies of the games finished since the last drain, int32[n, 244] (``packed_examples`` mode)."""
        if not self.packed:
            raise RuntimeError("drain_packed needs SelfPlay(packed_examples=True)")
        n = self.n_examples()
        out = self.examples[:n].clone()
        self.cursor.zero_()
        return out

    def drain_examples(self) -> torch.Tensor:
        """Rows produced since the last drain, float32[n, 901] (planes 675, pi 225, z) on the device."""
        if self.packed:
            return expand_examples(self.drain_packed(), self.use_symmetries)
        n = self.n_examples()
        out = self.examples[:n].clone()
        self.cursor.zero_()
        return out

    @staticmethod
    def split(rows: torch.Tensor):
        """-> (states [n,3,15,15], pis [n,225], zs [n,1]) as ReplayBuffer.sample returns them (train.py:287-293)."""
        return rows[:, :675].reshape(-1, 3, 15, 15), rows[:, 675:900], rows[:, 900:901]


class PipelinedSelfPlay:
    """Two half-size ``SelfPlay`` groups on two CUDA streams.  While one group's leaf batch is in the
    tensor-core trunk, the other group's FILL kernel (a warp per game, no shared memory, latency
    bound) runs next to it on the same SMs, so the tree walk disappears from the critical path.
    Results per game are identical to the unpipelined driver (games are independent and the RNG is
    keyed by the global game id)."""

    def __init__(self, model, n_games: int = 2048, game_base: int = 0, device="cuda:0", **kw):
        if n_games % 2:
            raise ValueError("n_games must be even")
        self.device = torch.device(device)
        self.G = n_games
        h = n_games // 2
        cap = kw.pop("example_capacity", 1 << 20)
        self.halves = [SelfPlay(model, n_games=h, game_base=game_base + i * h, device=device, example_capacity=cap // 2, **kw)
                       for i in range(2)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(2)]
        self.n_sims = self.halves[0].n_sims

    def close(self):
        for s in self.halves:
            s.close()

    def step(self):
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            st.wait_stream(cur)
        for sp, st in zip(self.halves, self.streams):
            with torch.cuda.stream(st):
                sp.search_begin()
        active = [True, True]
        while any(active):
            for i, (sp, st) in enumerate(zip(self.halves, self.streams)):
                if active[i]:
                    with torch.cuda.stream(st):
                        active[i] = sp.search_round()
        out = []
        for sp, st in zip(self.halves, self.streams):
            with torch.cuda.stream(st):
                pi, _ = sp.engine.result()
                out.append(sp.finish_step(pi))
        for st in self.streams:
            cur.wait_stream(st)
        return torch.cat(out)

    # aggregated views used by bench.py / tests
    @property
    def last_pi(self):
        return torch.cat([s.last_pi for s in self.halves])

    @property
    def actions(self):
        return torch.cat([s.actions for s in self.halves])

    @property
    def done(self):
        return torch.cat([s.done for s in self.halves])

    def counters(self) -> dict:
        keys = ("total_sims", "total_evals", "total_rounds", "total_launches")
        return {k: sum(getattr(s, k) for s in self.halves) for k in keys}

    def stats(self) -> dict:
        a, b = (s.engine.stats() for s in self.halves)
        out = {k: a[k] + b[k] for k in a}
        out["max_nodes"] = max(a["max_nodes"], b["max_nodes"])
        out["error_bits"] = a["error_bits"] | b["error_bits"]
        return out

    def drain_examples(self) -> torch.Tensor:
        return torch.cat([s.drain_examples() for s in self.halves])
