"""Oracle: deferred-evaluation PUCT search (TEST INFRASTRUCTURE ONLY).

Restates mcts/new_mcts_alpha.py of the reference as an explicit-stack state
machine (the reference recurses).  The load-bearing quirks are kept on purpose
and each is cited:

* new leaves are queued and given a placeholder node with a uniform-legal prior;
  the simulation returns 0 (new_mcts_alpha.py:114-132);
* when the queue reaches ``queue_len`` it is evaluated at once and the leaf that
  filled it is NOT returned from - selection continues at that leaf (:121-135);
* evaluation overwrites P and zeroes N and W of every queued node (:176-180);
* the network value is stored but never backed up; only terminal -1 / 0 move up
  the path (:106-112, :128-132, :146-151);
* priors are masked by legality without renormalising, with a uniform-legal
  fallback under 1e-8 (:164-168);
* root Dirichlet noise only when the root itself is evaluated in this run (:171-174).

Pinned against the reference by ``oracle/make_golden.py``.
"""
from __future__ import annotations

import math

import numpy as np

from . import rules


class Search:
    def __init__(self, rule, n_sims, model, cpuct=1.0, queue_len=32, alpha=0.03, eps=0.03,
                 noise_plies=10, noise=True, noise_fn=None):
        self.rule = rule
        self.n_sims = n_sims
        self.model = model
        self.cpuct = cpuct
        self.queue_len = queue_len
        self.alpha = alpha
        self.eps = eps
        self.noise_plies = noise_plies
        self.noise = noise
        # noise_fn(n) -> float64[n]; defaults to numpy's global generator like the reference (:172)
        self.noise_fn = noise_fn or (lambda n: np.random.dirichlet([self.alpha] * n))
        self.reset()

    def reset(self):
        """new_mcts_alpha.py:58-72."""
        self.P, self.Nv, self.W, self.legal = {}, {}, {}, {}
        self.queue = []          # (key, encoded planes, legal mask at queue time)
        self.root_key = None
        self.n_evals = 0
        self.n_batches = 0
        self.n_visits = 0        # node visits (search() entries in the reference)

    # ------------------------------------------------------------------ evaluation
    def _flush(self, ply):
        """new_mcts_alpha.py:156-185."""
        if not self.queue:
            return
        X = np.stack([q[1] for q in self.queue], axis=0).astype(np.float32)
        probs, _values = self.model.predict(X)
        self.n_evals += len(self.queue)
        self.n_batches += 1
        for (key, _x, ok), p in zip(self.queue, probs):
            p = p.flatten() * ok
            if np.sum(p) < 1e-8:
                p = ok / np.sum(ok)
            if self.noise and key == self.root_key and ply < self.noise_plies:
                d = self.noise_fn(len(p))
                p = (1 - self.eps) * p + self.eps * d
                p /= np.sum(p)
            self.P[key] = p
            self.Nv[key] = np.zeros_like(p, dtype=np.float32)
            self.W[key] = np.zeros_like(p, dtype=np.float32)
            self.legal[key] = ok
        self.queue = []

    # ------------------------------------------------------------------ one simulation
    def _simulate(self, root: rules.Position, ply):
        pos = root.copy()
        path = []                                   # (key, action) from the root down
        while True:
            self.n_visits += 1
            key = pos.key()
            if rules.game_over(pos):                # :106-112
                v = 0 if rules.winner(pos) == 0 else -1
                break
            if key not in self.P:                   # :114-132
                ok = rules.legal_mask(pos)
                self.queue.append((key, rules.encode(pos), ok))
                if len(self.queue) >= self.queue_len:
                    self._flush(ply)
                if key not in self.P:
                    self.P[key] = ok / np.sum(ok)
                    self.Nv[key] = np.zeros_like(ok, dtype=np.float32)
                    self.W[key] = np.zeros_like(ok, dtype=np.float32)
                    self.legal[key] = ok
                    v = 0
                    break
                # else: fall through and select at the freshly evaluated leaf
            ok = self.legal[key]                    # :135-140
            n, w, p = self.Nv[key], self.W[key], self.P[key]
            root_n = math.sqrt(np.sum(n))
            score = w / (1 + n) + self.cpuct * p * root_n / (1 + n)
            score = np.where(ok == 1, score, -1e9)
            a = int(np.argmax(score))
            path.append((key, a))
            rules.play(pos, a)                      # :141-144
        for key, a in reversed(path):               # :146-151
            v = -v
            self.W[key][a] += v
            self.Nv[key][a] += 1

    # ------------------------------------------------------------------ public
    def run(self, root: rules.Position, ply) -> np.ndarray:
        """new_mcts_alpha.py:77-97: visit distribution of the root after n_sims."""
        self.root_key = root.key()
        for _ in range(self.n_sims):
            self._simulate(root, ply)
        self._flush(ply)
        counts = self.Nv[self.root_key]
        total = np.sum(counts)
        if total > 0:
            return counts / total
        ok = self.legal[self.root_key]
        return ok / np.sum(ok)


def dihedral8(planes: np.ndarray, pi: np.ndarray):
    """The 8 board symmetries in the reference's order (new_mcts_alpha.py:42-56):
    for k in 0..3: rot90 by k, then that rotation mirrored left-right."""
    n = planes.shape[1]
    grid = pi.reshape(n, n)
    out = []
    for k in range(4):
        s = np.rot90(planes, k, axes=(1, 2))
        g = np.rot90(grid, k)
        out.append((s, g.flatten()))
        out.append((np.flip(s, axis=2), np.flip(g, axis=1).flatten()))
    return out
