"""Batched search engine: Python face of the azg_* search entry points.

One ``SearchEngine`` holds G concurrent games of one rule on one B200, each with
its own HBM-resident tree slab, and advances all of them in lock step:

    begin -> [ fill -> evaluate leaves -> commit ]* -> result -> advance

which is the reference's ``MCTS.run`` (mcts/new_mcts_alpha.py:77-97) for G games
at once.  Device buffers are torch tensors; all compute is in the CUDA library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import azg_config, check, lib, ptr

GOMOKU, PENTE = 0, 1
RULE_BY_NAME = {"gomoku": GOMOKU, "pente": PENTE}
POS_WORDS = 24          # azg_pos as int32[24]


def rule_of(game_class_or_name) -> int:
    """Map the reference's ``game_class`` argument (a class, an instance or a name) to a rule id."""
    name = game_class_or_name if isinstance(game_class_or_name, str) else getattr(
        game_class_or_name, "__name__", type(game_class_or_name).__name__)
    name = name.lower()
    if name not in RULE_BY_NAME:
        raise ValueError(f"Unsupported rules: {name}. Only 'gomoku' and 'pente' are supported.")
    return RULE_BY_NAME[name]


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Rules:
    """Batched rule kernels over packed positions (int32[n,24] device tensors)."""

    def __init__(self, rule: int, device="cuda:0"):
        self.rule = rule
        self.device = torch.device(device)

    def pack(self, boards, players, lasts=None, caps=None, plies=None) -> torch.Tensor:
        """boards int8[n,225] etc. (numpy or tensors) -> packed positions on the device."""
        dev = self.device
        b = torch.as_tensor(np.ascontiguousarray(boards) if isinstance(boards, np.ndarray) else boards).to(dev, torch.int8).reshape(-1, 225).contiguous()
        n = b.shape[0]
        as_i32 = lambda x: None if x is None else torch.as_tensor(x).to(dev, torch.int32).contiguous()
        pl, la, ca, pi = as_i32(players), as_i32(lasts), as_i32(caps), as_i32(plies)
        out = torch.zeros((n, POS_WORDS), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.azg_rules_pack(ptr(b), ptr(pl), ptr(la), ptr(ca), ptr(pi), ptr(out), n, _stream()))
        return out

    def unpack(self, pos: torch.Tensor):
        n = pos.shape[0]
        dev = pos.device
        boards = torch.empty((n, 225), dtype=torch.int8, device=dev)
        players = torch.empty(n, dtype=torch.int32, device=dev)
        lasts = torch.empty(n, dtype=torch.int32, device=dev)
        caps = torch.empty((n, 2), dtype=torch.int32, device=dev)
        plies = torch.empty(n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            check(lib.azg_rules_unpack(ptr(pos), ptr(boards), ptr(players), ptr(lasts), ptr(caps), ptr(plies), n, _stream()))
        return boards, players, lasts, caps, plies

    def play(self, pos: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
        """In-place do_move on every position; returns the status bits (int32[n])."""
        n = pos.shape[0]
        status = torch.empty(n, dtype=torch.int32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.azg_rules_play(self.rule, ptr(pos), ptr(actions.to(torch.int32).contiguous()), ptr(status), n, _stream()))
        return status

    def status(self, pos: torch.Tensor) -> torch.Tensor:
        n = pos.shape[0]
        status = torch.empty(n, dtype=torch.int32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.azg_rules_status(self.rule, ptr(pos), ptr(status), n, _stream()))
        return status

    def legal(self, pos: torch.Tensor) -> torch.Tensor:
        n = pos.shape[0]
        out = torch.empty((n, 225), dtype=torch.float32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.azg_rules_legal(ptr(pos), ptr(out), n, _stream()))
        return out

    def encode(self, pos: torch.Tensor) -> torch.Tensor:
        n = pos.shape[0]
        out = torch.empty((n, 3, 15, 15), dtype=torch.float32, device=pos.device)
        with torch.cuda.device(pos.device):
            check(lib.azg_rules_encode(ptr(pos), ptr(out), n, _stream()))
        return out


class SearchEngine:
    def __init__(self, rule: int, n_games: int, cpuct: float = 1.0, queue_len: int = 32, node_capacity: int = 8192,
                 noise: bool = False, alpha: float = 0.03, eps: float = 0.03, noise_plies: int = 10, seed: int = 12345,
                 device="cuda:0", game_base: int = 0, fast_warps: int = 0, virtual_loss: int = 1):
        """``fast_warps`` = 0: the reference's algorithm, exact visit counts.  1..16: the NON-PARITY fast mode - that many
        warps walk each game's tree concurrently under a virtual loss (see tree.cu); for latency, not for parity."""
        if not torch.cuda.is_available():
            raise _lib.AzgError("no CUDA device: azgomoku_b200 has no CPU fallback")
        self.device = torch.device(device)
        self.rule, self.G, self.queue_len = rule, n_games, queue_len
        cfg = azg_config(device=self.device.index or 0, rule=rule, n_games=n_games, queue_len=queue_len,
                         node_capacity=node_capacity, noise_on=int(noise), noise_plies=noise_plies, game_base=int(game_base),
                         cpuct=float(cpuct), alpha=float(alpha), eps=float(eps), seed=seed, fast_warps=int(fast_warps),
                         virtual_loss=int(virtual_loss))
        h = C.c_void_p()
        check(lib.azg_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.rules = Rules(rule, self.device)
        self.max_leaves = n_games * queue_len

    def close(self):
        if getattr(self, "_h", None):
            lib.azg_destroy(self._h)
            self._h = None

    __del__ = close

    # -- helpers
    def _sync_stream(self):
        """Every engine call runs on the engine's own device and on torch's current stream there."""
        idx = self.device.index or 0
        if torch.cuda.current_device() != idx:
            torch.cuda.set_device(idx)
        check(lib.azg_set_stream(self._h, _stream()))

    @property
    def memory_bytes(self) -> int:
        return int(lib.azg_memory_bytes(self._h))

    def set_roots(self, pos: torch.Tensor, mask: torch.Tensor | None = None, clear_tree: bool = True):
        self._sync_stream()
        m = None if mask is None else mask.to(self.device, torch.int32).contiguous()
        check(lib.azg_set_roots(self._h, ptr(pos), ptr(m), int(clear_tree)))

    def clear(self, mask: torch.Tensor | None = None):
        self._sync_stream()
        m = None if mask is None else mask.to(self.device, torch.int32).contiguous()
        check(lib.azg_set_roots(self._h, None, ptr(m), 1))

    def roots(self) -> torch.Tensor:
        self._sync_stream()
        out = torch.empty((self.G, POS_WORDS), dtype=torch.int32, device=self.device)
        check(lib.azg_get_roots(self._h, ptr(out)))
        return out

    def begin(self, n_sims: int, plies: torch.Tensor | None = None, mask: torch.Tensor | None = None):
        """Start a run; games with mask == 0 (if given) sit it out."""
        self._sync_stream()
        p = None if plies is None else plies.to(self.device, torch.int32).contiguous()
        m = None if mask is None else mask.to(self.device, torch.int32).contiguous()
        check(lib.azg_search_begin_masked(self._h, ptr(p), int(n_sims), ptr(m)))

    def fill(self):
        """-> (n_leaves, n_more, n_roots) after a host sync: batch size, games that need another
        fill after this batch, games whose root is in the batch."""
        self._sync_stream()
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(lib.azg_search_fill(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def fill_async(self):
        self._sync_stream()
        check(lib.azg_search_fill(self._h, None, None, None))

    def read_counters(self):
        """-> (n_leaves, n_more, n_roots) of the last ``fill_async`` (synchronises this engine's stream only)."""
        self._sync_stream()
        a, b, c = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        check(lib.azg_search_read_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def counters_ptr(self) -> int:
        return int(lib.azg_search_counters(self._h))

    def leaf_planes(self, n_leaves: int) -> torch.Tensor:
        self._sync_stream()
        out = torch.empty((max(n_leaves, 1), 3, 15, 15), dtype=torch.float32, device=self.device)
        check(lib.azg_search_leaf_planes(self._h, ptr(out)))
        return out[:n_leaves]

    def commit(self, probs: torch.Tensor, noise: torch.Tensor | None = None):
        self._sync_stream()
        assert probs.dtype == torch.float32 and probs.is_contiguous()
        if noise is not None:
            assert noise.dtype == torch.float64 and noise.shape == (self.G, 225) and noise.is_contiguous()
        check(lib.azg_search_commit(self._h, ptr(probs), ptr(noise)))

    def result(self):
        self._sync_stream()
        pi = torch.empty((self.G, 225), dtype=torch.float32, device=self.device)
        visits = torch.empty((self.G, 225), dtype=torch.int32, device=self.device)
        check(lib.azg_search_result(self._h, ptr(pi), ptr(visits)))
        return pi, visits

    def advance(self, actions: torch.Tensor, gc: bool = True, reserve: int = 0) -> torch.Tensor:
        """Play actions (>= 0) on the roots and sweep dead nodes; ``reserve`` > 0 drops the tree of
        any game that could not hold another run of that many new nodes."""
        self._sync_stream()
        a = actions.to(self.device, torch.int32).contiguous()
        status = torch.empty(self.G, dtype=torch.int32, device=self.device)
        check(lib.azg_search_advance(self._h, ptr(a), int(gc), int(reserve), ptr(status)))
        return status

    def stats(self) -> dict:
        self._sync_stream()
        out = (C.c_uint64 * 8)()
        check(lib.azg_search_stats(self._h, out))
        keys = ("sims", "visits", "evals", "live_nodes", "max_nodes", "games_in_error", "error_bits", "dropped_trees")
        return {k: int(out[i]) for i, k in enumerate(keys)}

    # -- the reference's run() for all games, evaluator supplied by the caller
    def run(self, n_sims: int, evaluate, plies: torch.Tensor | None = None, noise: torch.Tensor | None = None):
        """``evaluate(planes f32[L,3,15,15] on device) -> probs f32[L,225] on device``.
        Returns (pi, visits).  This is the generic path (any evaluator, one host sync per
        round); ``selfplay.SelfPlay`` drives the fused on-device evaluator instead."""
        self.begin(n_sims, plies)
        while True:
            n_leaves, n_more, _ = self.fill()
            if n_leaves > 0:
                planes = self.leaf_planes(n_leaves)
                probs = evaluate(planes)
                self.commit(probs.contiguous(), noise)
            if n_more == 0:
                break
        return self.result()
