"""Oracle: Gomoku / Pente rules on a flat 225-cell board (TEST INFRASTRUCTURE ONLY).

Restates games/gomoku.py and games/pente.py of the reference as free functions
over one small state record.  Cells are indexed ``a = r*15 + c`` exactly like the
reference's action index (gomoku.py:46-55).  Pinned against the reference by
``oracle/make_golden.py`` (bit-exact on every ply of the committed traces).
"""
from __future__ import annotations

import numpy as np

N = 15            # board edge (gomoku.py:20, pente.py:12)
A = N * N         # action count (gomoku.py:43)
GOMOKU, PENTE = 0, 1

# line directions for five-in-a-row (gomoku.py:171, pente.py:213)
LINE_DIRS = ((1, 0), (0, 1), (1, 1), (1, -1))
# custodial-capture directions, in the reference's order (pente.py:124-129)
CAPTURE_DIRS = ((1, 0), (-1, 0), (0, 1), (0, -1), (1, 1), (-1, -1), (1, -1), (-1, 1))


class Position:
    """Board + side to move + the path-local extras the reference carries.

    ``last`` is the action index of the most recent stone (-1 == the
    reference's ``last_move is None``); ``caps[p-1]`` is Pente's
    ``captures[p]`` (pente.py:19).  ``plies`` mirrors ``len(move_history)``.
    """

    __slots__ = ("rule", "cells", "player", "last", "caps", "plies")

    def __init__(self, rule: int = GOMOKU):
        self.rule = rule
        self.cells = np.zeros(A, dtype=np.int8)
        self.player = 1
        self.last = -1
        self.caps = [0, 0]
        self.plies = 0

    def copy(self) -> "Position":
        q = Position(self.rule)
        q.cells = self.cells.copy()
        q.player = self.player
        q.last = self.last
        q.caps = list(self.caps)
        q.plies = self.plies
        return q

    def key(self) -> bytes:
        """Transposition key: board bytes + side to move (new_mcts_alpha.py:190-197).
        Captures are NOT part of the key."""
        return self.cells.tobytes() + bytes([self.player])


def on_board(r: int, c: int) -> bool:
    return 0 <= r < N and 0 <= c < N


def play(pos: Position, action: int) -> bool:
    """Place a stone for the side to move.  Returns False and leaves ``pos``
    untouched for an off-board or occupied cell (gomoku.py:66-70, pente.py:58-62)."""
    if not (0 <= action < A) or pos.cells[action] != 0:
        return False
    me = pos.player
    foe = 3 - me
    pos.cells[action] = me
    pos.last = action
    pos.plies += 1
    if pos.rule == PENTE:
        r, c = divmod(action, N)
        # pente.py:131-150: me, foe, foe, me along each of the eight rays
        for dr, dc in CAPTURE_DIRS:
            r3, c3 = r + 3 * dr, c + 3 * dc
            if not on_board(r3, c3):
                continue            # r1,r2 lie between, so they are on board too
            a1 = (r + dr) * N + (c + dc)
            a2 = (r + 2 * dr) * N + (c + 2 * dc)
            a3 = r3 * N + c3
            if pos.cells[a1] == foe and pos.cells[a2] == foe and pos.cells[a3] == me:
                pos.cells[a1] = 0
                pos.cells[a2] = 0
                pos.caps[me - 1] += 1
    pos.player = foe
    return True


def play_rc(pos: Position, r: int, c: int) -> bool:
    """(r, c) entry point with the reference's range check (gomoku.py:67)."""
    if not on_board(r, c):
        return False
    return play(pos, r * N + c)


def winner(pos: Position) -> int:
    """0 / 1 / 2, judged only through the last stone (gomoku.py:155-193,
    pente.py:199-233).  Overlines (6+) win.  Pente checks the capture count of the
    last stone's owner first (pente.py:209)."""
    if pos.last < 0:
        return 0
    who = int(pos.cells[pos.last])
    if who == 0:
        return 0
    if pos.rule == PENTE and pos.caps[who - 1] >= 5:
        return who
    r, c = divmod(pos.last, N)
    for dr, dc in LINE_DIRS:
        run = 1
        for sgn in (1, -1):
            rr, cc = r + sgn * dr, c + sgn * dc
            while on_board(rr, cc) and pos.cells[rr * N + cc] == who:
                run += 1
                rr += sgn * dr
                cc += sgn * dc
        if run >= 5:
            return who
    return 0


def legal_mask(pos: Position) -> np.ndarray:
    """float32[225], 1.0 on empty cells (gomoku.py:109-121, pente.py:164-172)."""
    return (pos.cells == 0).astype(np.float32)


def game_over(pos: Position) -> bool:
    """Winner through the last move, or no empty cell (gomoku.py:195-197)."""
    return winner(pos) != 0 or not bool((pos.cells == 0).any())


def encode(pos: Position) -> np.ndarray:
    """float32[3,15,15]: side-to-move stones, opponent stones, constant ones
    (gomoku.py:130-150, pente.py:180-194)."""
    b = pos.cells.reshape(N, N)
    out = np.empty((3, N, N), dtype=np.float32)
    out[0] = b == pos.player
    out[1] = b == (3 - pos.player)
    out[2] = 1.0
    return out


def from_board(rule: int, board, player: int, last=None, caps=(0, 0), plies: int = 0) -> Position:
    """Build a Position from the reference's game-object fields."""
    pos = Position(rule)
    pos.cells = np.asarray(board).astype(np.int8).reshape(A).copy()
    pos.player = int(player)
    pos.last = -1 if last is None else int(last[0]) * N + int(last[1])
    pos.caps = [int(caps[0]), int(caps[1])]
    pos.plies = int(plies)
    return pos
