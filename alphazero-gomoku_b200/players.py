"""Drop-in AlphaZero ``Player`` on the B200 engine.

Interface contract (players/player_alpha.py:7-77 of the reference, loaded by name from play.py:19-30):
``Player(rules, board_size, n_simulations, c_puct, model_path, nn_model)``; ``play(board, turn_number,
last_opponent_move) -> (row, col)``; attributes ``n_simulations`` and ``model_path`` are read by
play_loop.py.  Only Gomoku is accepted, as in the reference.  The search tree lives on the GPU and is kept
between calls (the reference never clears it either).
"""
from __future__ import annotations

import numpy as np

from .games import Gomoku
from .mcts import MCTS
from .network import PyTorchModel

DEFAULT_SNAPSHOT = "models/snapshot_iter140_20260109_190822.pt"     # the reference's default path


def _as_board(board) -> np.ndarray:
    """The caller passes either a list of rows or a game object (player_alpha.py:59-62)."""
    cells = board if isinstance(board, list) else board.board
    return np.array(cells, dtype=int)


def _side_to_move(turn_number: int) -> int:
    """Player 1 moves on even turns (player_alpha.py:65)."""
    return 2 if turn_number % 2 else 1


class Player:
    def __init__(self, rules="gomoku", board_size=15, n_simulations=5000, c_puct=1.0, model_path=DEFAULT_SNAPSHOT,
                 nn_model=PyTorchModel, fast_warps=0, batch_size=32):
        """Reference arguments first (player_alpha.py:26-27).  ``fast_warps`` > 0 / a longer ``batch_size`` select the engine's
        non-parity fast search (lower latency per move, visit counts no longer the reference's); the defaults are exact."""
        self.rules, self.board_size = rules.lower(), board_size
        self.n_simulations, self.c_puct, self.model_path = n_simulations, c_puct, model_path
        self.net = nn_model(board_size=board_size)
        if model_path is None:
            print("[PlayerAlpha] no snapshot given: playing with random weights")
        else:
            print(f"[PlayerAlpha] loading {model_path}")
            self.net.load(model_path)
        self.net.net.eval()
        if self.rules != "gomoku":
            raise ValueError(f"Unsupported rules: {self.rules}. Only 'gomoku' is supported.")
        self.game_class = Gomoku
        self.mcts = MCTS(game_class=Gomoku, n_simulations=n_simulations, nn_model=self.net, cpuct=c_puct,
                         add_dirichlet_noise=False, fast_warps=fast_warps, batch_size=batch_size)

    def play(self, board, turn_number, last_opponent_move):
        """Search from the given position and answer with the most visited move."""
        position = self.game_class(size=self.board_size)
        position.board = _as_board(board)
        position.current_player = _side_to_move(turn_number)
        position.last_move = last_opponent_move
        visits = self.mcts.run(position, turn_number)
        best = int(np.argmax(visits))
        return best // self.board_size, best % self.board_size
