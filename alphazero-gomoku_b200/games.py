"""Drop-in ``Gomoku`` / ``Pente`` game objects (games/gomoku.py, games/pente.py) whose rule
evaluation - ``do_move`` with captures, ``check_winner``, ``is_game_over`` - runs in the CUDA rule
kernels through ``azg_rules_play_host``.  Same attributes and method names as the reference's
classes, so callers (train.py, play.py, players) are unchanged.  These single-game objects exist
for interface compatibility; throughput paths keep positions on the device (``engine.Rules``)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from ._lib import check, lib, ptr


class _Game:
    RULE = 0

    def __init__(self, size: int = 15):
        if size != 15:
            raise ValueError("azgomoku_b200 rule kernels are specialised for the 15x15 board")
        self.size = size
        self.board = np.zeros((size, size), dtype=np.int8)
        self.current_player = 1
        self.move_history: List[Tuple[int, int]] = []
        self.last_move: Optional[Tuple[int, int]] = None
        self._status = 0

    # ---- reference helpers (gomoku.py:42-55)
    @property
    def action_size(self) -> int:
        return self.size * self.size

    def action_to_move(self, action: int) -> Tuple[int, int]:
        return divmod(int(action), self.size)

    def move_to_action(self, move: Tuple[int, int]) -> int:
        return int(move[0] * self.size + move[1])

    def _caps(self):
        return [0, 0]

    def _set_caps(self, caps):
        pass

    def _device_step(self, action: int):
        boards = np.ascontiguousarray(self.board.reshape(1, -1).astype(np.int8))
        players = np.array([self.current_player], np.int32)
        lasts = np.array([-1 if self.last_move is None else self.last_move[0] * 15 + self.last_move[1]], np.int32)
        caps = np.array([self._caps()], np.int32)
        plies = np.array([len(self.move_history)], np.int32)
        acts = np.array([action], np.int32)
        status = np.zeros(1, np.int32)
        check(lib.azg_rules_play_host(self.RULE, 0, ptr(boards), ptr(players), ptr(lasts), ptr(caps), ptr(plies), ptr(acts), ptr(status), 1))
        return boards.reshape(15, 15), int(players[0]), int(lasts[0]), caps[0].tolist(), int(status[0])

    def do_move(self, move: Tuple[int, int]) -> bool:
        """gomoku.py:60-78 / pente.py:57-79: False (no raise) for an off-board or occupied cell."""
        r, c = int(move[0]), int(move[1])
        action = r * 15 + c if (0 <= r < 15 and 0 <= c < 15) else -1
        board, player, last, caps, status = self._device_step(action)
        if status & 8:
            return False
        before = self.board
        self.board = board.astype(before.dtype)
        self._on_captured(before, self.board, (r, c))
        self.current_player = player
        self.last_move = (r, c)
        self.move_history.append((r, c))
        self._set_caps(caps)
        self._status = status
        return True

    def _on_captured(self, before, after, move):
        pass

    def _query(self, want_legal=False, want_planes=False):
        """Status bits (and optionally the legal mask / encoded planes) of the current position from the
        CUDA rule kernels (azg_rules_query_host)."""
        boards = np.ascontiguousarray(self.board.reshape(1, -1).astype(np.int8))
        players = np.array([self.current_player], np.int32)
        lasts = np.array([-1 if self.last_move is None else self.last_move[0] * 15 + self.last_move[1]], np.int32)
        caps = np.array([self._caps()], np.int32)
        plies = np.array([len(self.move_history)], np.int32)
        status = np.zeros(1, np.int32)
        legal = np.zeros((1, 225), np.float32) if want_legal else None
        planes = np.zeros((1, 3, 15, 15), np.float32) if want_planes else None
        check(lib.azg_rules_query_host(self.RULE, 0, ptr(boards), ptr(players), ptr(lasts), ptr(caps), ptr(plies), ptr(status),
                                       ptr(legal), ptr(planes), 1))
        self._status = int(status[0])
        return legal, planes

    def _refresh(self):
        self._query()

    def check_winner(self) -> int:
        self._refresh()
        return self._status & 3

    def get_winner(self) -> int:
        return self.check_winner()

    def is_game_over(self) -> bool:
        self._refresh()
        return bool(self._status & 4)

    def get_legal_moves(self):
        e = np.where(self.board == 0)
        return list(zip(e[0].tolist(), e[1].tolist()))

    def has_legal_moves(self) -> bool:
        return bool(np.any(self.board == 0))

    def get_valid_moves(self) -> np.ndarray:
        return self._query(want_legal=True)[0][0]

    def get_state(self) -> np.ndarray:
        return self.board.copy()

    def get_encoded_state(self) -> np.ndarray:
        return self._query(want_planes=True)[1][0]

    def display(self) -> None:
        """Text rendering of the position (the reference prints a coloured grid; same information)."""
        marks = ".XO"
        print("    " + " ".join(f"{c + 1:2}" for c in range(self.size)))
        for r in range(self.size):
            print(f"{r + 1:2}   " + "  ".join(marks[int(v)] for v in self.board[r]))
        print(f"to move: {marks[self.current_player]}  last: {self.last_move}")


class Gomoku(_Game):
    RULE = 0

    def clone(self) -> "Gomoku":
        g = Gomoku(self.size)
        g.board = self.board.copy()
        g.current_player = int(self.current_player)
        g.move_history = list(self.move_history)
        g.last_move = None if self.last_move is None else tuple(self.last_move)
        return g

    def undo_move(self) -> None:
        if not self.move_history:
            return
        r, c = self.move_history.pop()
        self.board[r, c] = 0
        self.current_player = 3 - self.current_player
        self.last_move = self.move_history[-1] if self.move_history else None


class Pente(_Game):
    RULE = 1

    def __init__(self, size: int = 15):
        super().__init__(size)
        self.captures = {1: 0, 2: 0}
        self.capture_history: List[List[Tuple[int, int]]] = []

    def _caps(self):
        return [self.captures[1], self.captures[2]]

    def _set_caps(self, caps):
        self.captures = {1: int(caps[0]), 2: int(caps[1])}

    def _on_captured(self, before, after, move):
        gone = np.argwhere((before != 0) & (after == 0))
        self.capture_history.append([(int(r), int(c)) for r, c in gone])

    def clone(self) -> "Pente":
        g = Pente(self.size)
        g.board = self.board.copy()
        g.current_player = int(self.current_player)
        g.last_move = None if self.last_move is None else tuple(self.last_move)
        g.captures = dict(self.captures)
        g.move_history = list(self.move_history)
        g.capture_history = [list(x) for x in self.capture_history]
        return g

    def undo_move(self) -> None:
        if not self.move_history:
            return
        self.current_player = 3 - self.current_player
        r, c = self.move_history.pop()
        captured = self.capture_history.pop()
        self.board[r, c] = 0
        for rr, cc in captured:
            # bug-compatible with pente.py:100-103: the reference restores captured stones in the
            # CAPTURER's colour; undo is not on the search path (the search clones), kept identical
            self.board[rr, cc] = self.current_player
        self.captures[self.current_player] -= len(captured) // 2
        self.last_move = self.move_history[-1] if self.move_history else None
