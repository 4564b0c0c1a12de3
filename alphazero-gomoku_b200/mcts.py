"""Drop-in ``MCTS`` (same constructor, ``run``, ``symmetries``, ``clear_tree``) on the B200 engine.

Mirrors the public surface of the reference's ``mcts.new_mcts_alpha.MCTS``
(mcts/new_mcts_alpha.py:12-97): one instance per game, the tree persists between
``run`` calls until ``clear_tree()``, ``run`` does not mutate the game object, and
any object with ``predict(X) -> (probs, values)`` can be injected as ``nn_model``.
The search itself (selection, expansion, backup, transposition table, deferred
evaluation queue) runs in the CUDA library; this class is a G = 1 face of
``SearchEngine``.

Evaluator paths:
* ``nn_model`` with ``predict_device`` (this package's ``PyTorchModel``): leaf planes
  never leave the GPU;
* anything else (the reference's own ``PyTorchModel``, test fakes): planes are copied
  to the host, ``predict`` is called exactly as the reference calls it
  (new_mcts_alpha.py:160-161) and the probabilities are copied back.
"""
from __future__ import annotations

import numpy as np
import torch

from .engine import SearchEngine, rule_of


def game_fields(game_state):
    """(board int8[225], player, last, caps, plies) from a reference-style game object."""
    board = np.asarray(game_state.board).astype(np.int8).reshape(-1)
    last = getattr(game_state, "last_move", None)
    caps = getattr(game_state, "captures", None)
    return (board, int(game_state.current_player), -1 if last is None else int(last[0]) * 15 + int(last[1]),
            (0, 0) if caps is None else (int(caps[1]), int(caps[2])), len(getattr(game_state, "move_history", ())))


class MCTS:
    def __init__(self, game_class, n_simulations, nn_model, cpuct=1.0, batch_size=32, dirichlet_alpha=0.03,
                 epsilon=0.03, apply_dirichlet_n_first_moves=10, add_dirichlet_noise=True,
                 node_capacity=65536, device="cuda:0", gc=True, fast_warps=0, virtual_loss=1):
        """The reference's constructor arguments first (same names, order, defaults); the rest are engine extras.
        ``fast_warps`` > 0 selects the non-parity fast mode (concurrent simulations under a virtual loss): several
        times lower latency per move, visit counts no longer those of the reference."""
        self.game_class = game_class
        self.n_simulations = n_simulations
        self.nn_model = nn_model
        self.cpuct = cpuct
        self.batch_size = batch_size
        self.dirichlet_alpha = dirichlet_alpha
        self.epsilon = epsilon
        self.apply_dirichlet_n_first_moves = apply_dirichlet_n_first_moves
        self.add_dirichlet_noise = add_dirichlet_noise
        self.action_size = 225
        self.gc = gc
        self.rule = rule_of(game_class)
        self.engine = SearchEngine(self.rule, 1, cpuct=cpuct, queue_len=batch_size, node_capacity=node_capacity,
                                   noise=add_dirichlet_noise, alpha=dirichlet_alpha, eps=epsilon,
                                   noise_plies=apply_dirichlet_n_first_moves, device=device, fast_warps=fast_warps,
                                   virtual_loss=virtual_loss)
        self.device = self.engine.device
        self.last_visits = None         # int32[225] root visit counts of the last run (N[root] in the reference)
        self.n_evals = 0
        self._fresh = True
        self._probs = None

    # ------------------------------------------------------------------ reference helpers
    def symmetries(self, state, pi):
        """Eight dihedral images, reference order (new_mcts_alpha.py:42-56)."""
        size = state.shape[1]
        grid = np.asarray(pi).reshape(size, size)
        out = []
        for k in range(4):
            s = np.rot90(state, k, axes=(1, 2))
            g = np.rot90(grid, k)
            out.append((s, g.flatten()))
            out.append((np.flip(s, axis=2), np.flip(g, axis=1).flatten()))
        return out

    def clear_tree(self):
        """new_mcts_alpha.py:58-72."""
        self.engine.clear()
        self._fresh = True

    def search(self, game_state, move_number):
        """The reference's per-simulation recursion (new_mcts_alpha.py:102-151) has no host-side counterpart:
        simulations run inside ``azg_fill_kernel``, ``n_simulations`` per ``run``."""
        raise NotImplementedError("MCTS.search is internal to the CUDA engine; call run(game_state, move_number)")

    # ------------------------------------------------------------------ evaluation
    def _evaluate(self, planes: torch.Tensor) -> torch.Tensor:
        fast = getattr(self.nn_model, "predict_device", None)
        if fast is not None:
            probs, _ = fast(planes)
            return probs
        X = planes.cpu().numpy()
        probs, _values = self.nn_model.predict(X)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(probs, dtype=np.float32).reshape(len(X), -1))).to(self.device)

    # ------------------------------------------------------------------ run
    def run(self, game_state, move_number):
        """Visit distribution float32[225] of the root after ``n_simulations`` (new_mcts_alpha.py:77-97)."""
        eng = self.engine
        board, player, last, caps, plies = game_fields(game_state)
        pos = eng.rules.pack(board[None, :], [player], [last], [list(caps)], [plies])
        eng.set_roots(pos, clear_tree=False)
        if self.gc and not self._fresh:
            eng.advance(torch.full((1,), -1, dtype=torch.int32, device=self.device), gc=True,
                        reserve=self.n_simulations + self.n_simulations // max(self.batch_size, 1) + 8)
        self._fresh = False
        eng.begin(self.n_simulations, torch.tensor([int(move_number)], dtype=torch.int32, device=self.device))
        net = self._device_net()
        if net is not None:
            self._run_on_device(net, move_number)
        else:
            self._run_with_host_model(move_number)
        pi, visits = eng.result()
        self.last_visits = visits[0].cpu().numpy()
        return pi[0].cpu().numpy()

    def _root_noise(self, n_roots, move_number):
        """numpy's global generator is advanced exactly when the reference advances it: once per run whose
        root is evaluated while noise applies (new_mcts_alpha.py:170-172)."""
        if not (n_roots and self.add_dirichlet_noise and move_number < self.apply_dirichlet_n_first_moves):
            return None
        d = np.random.dirichlet([self.dirichlet_alpha] * self.action_size)
        return torch.from_numpy(np.ascontiguousarray(d, dtype=np.float64)[None, :]).to(self.device)

    def _run_with_host_model(self, move_number):
        """Any ``predict``-style evaluator: one host round trip per queue flush, as in the reference."""
        eng = self.engine
        while True:
            n_leaves, n_more, n_roots = eng.fill()
            if n_leaves > 0:
                probs = self._evaluate(eng.leaf_planes(n_leaves))
                eng.commit(probs.contiguous(), self._root_noise(n_roots, move_number))
                self.n_evals += n_leaves
            if n_more == 0:
                break

    def _device_net(self):
        """The CUDA leaf evaluator of this package's ``PyTorchModel`` (None for any other evaluator)."""
        ensure = getattr(self.nn_model, "_ensure_engine", None)
        if ensure is None:
            return None
        net = ensure()
        same_gpu = (net.device.index or 0) == (self.device.index or 0)
        return net if net.max_batch >= self.batch_size and same_gpu else None

    ROUNDS_PER_SYNC = 8

    def _run_on_device(self, net, move_number):
        """Latency path: the leaf batch never leaves the GPU and the host looks at the counters
        only every ``ROUNDS_PER_SYNC`` rounds; rounds launched after the run has finished are no-ops (every
        kernel takes its work count from device memory).  The first round is synchronous because the root
        can only be evaluated there, and the reference draws its Dirichlet noise on the host."""
        eng = self.engine
        if self._probs is None:
            self._probs = torch.empty((self.batch_size, 225), dtype=torch.float32, device=self.device)
        n_leaves, n_more, n_roots = eng.fill()
        if n_leaves > 0:
            net.forward_leaves(eng, self._probs)
            eng.commit(self._probs, self._root_noise(n_roots, move_number))
        while n_more:
            for _ in range(self.ROUNDS_PER_SYNC):
                eng.fill_async()
                net.forward_leaves(eng, self._probs)
                eng.commit(self._probs, None)
            _, n_more, _ = eng.read_counters()
        self.n_evals = eng.stats()["evals"]
