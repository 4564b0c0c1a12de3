"""Oracle: one self-play game and its training examples (TEST INFRASTRUCTURE ONLY).

Restates train.py:252-266 (temperature sampling) and train.py:360-412
(``play_game_and_collect``) over ``oracle.rules`` / ``oracle.search``.
"""
from __future__ import annotations

import numpy as np

from . import rules
from .search import Search, dihedral8


def temper(pi: np.ndarray, temp: float) -> np.ndarray:
    """train.py:252-259."""
    if temp <= 0:
        return pi
    z = np.log(pi + 1e-15) / temp
    e = np.exp(z - np.max(z))
    return e / np.sum(e)


def pick(pi: np.ndarray, temp: float, choice=None) -> int:
    """train.py:262-266.  ``choice(n, p)`` defaults to numpy's global generator."""
    if temp == 0:
        return int(np.argmax(pi))
    p = temper(pi, temp)
    choice = choice or (lambda n, p: np.random.choice(n, p=p))
    return int(choice(len(p), p))


def play_one(search: Search, pos: rules.Position, temp_fn, max_plies=225, expand=True, choice=None):
    """train.py:360-412: returns (examples, winner); examples are
    (planes f32[3,15,15], pi f32[225], z)."""
    rows = []
    ply = 0
    while True:
        planes = rules.encode(pos)
        pi = search.run(pos, pos.plies)
        keep = pi.copy()
        a = pick(pi, temp_fn(ply), choice)
        if rules.legal_mask(pos)[a] != 1.0:
            a = int(np.argmax(pi))
        rows.append((planes, keep, pos.player))
        rules.play(pos, a)
        ply += 1
        if rules.game_over(pos) or ply >= max_plies:
            break
    won = rules.winner(pos)
    out = []
    for planes, pi, who in rows:
        z = 0.0 if won == 0 else (1.0 if won == who else -1.0)
        if expand:
            for s, g in dihedral8(planes, pi):
                out.append((s.astype(np.float32), g.astype(np.float32), z))
        else:
            out.append((planes.astype(np.float32), pi.astype(np.float32), z))
    return out, won
