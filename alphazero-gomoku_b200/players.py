"""Drop-in ``Player`` (players/player_alpha.py:7-77): same constructor and ``play`` signature, the
search and the network run on the B200 engine."""
from __future__ import annotations

import numpy as np

from .games import Gomoku
from .mcts import MCTS
from .network import PyTorchModel


class Player:
    def __init__(self, rules="gomoku", board_size=15, n_simulations=5000, c_puct=1.0,
                 model_path="models/snapshot_iter140_20260109_190822.pt", nn_model=PyTorchModel):
        self.rules = rules.lower()
        self.board_size = board_size
        self.n_simulations = n_simulations
        self.c_puct = c_puct
        self.model_path = model_path
        self.net = nn_model(board_size=self.board_size)
        if model_path is not None:
            print(f"[PlayerAlpha] loading model: {model_path}")
            self.net.load(model_path)
        else:
            print("[PlayerAlpha] WARNING: no model given, using random weights")
        self.net.net.eval()
        if self.rules != "gomoku":                     # player_alpha.py:35-36
            raise ValueError(f"Unsupported rules: {self.rules}. Only 'gomoku' is supported.")
        self.game_class = Gomoku
        self.mcts = MCTS(game_class=self.game_class, n_simulations=self.n_simulations, nn_model=self.net,
                         cpuct=self.c_puct, add_dirichlet_noise=False)

    def play(self, board, turn_number, last_opponent_move):
        """player_alpha.py:51-77: rebuild the position, search, return the argmax move (r, c)."""
        game = self.game_class(size=self.board_size)
        if isinstance(board, list):
            game.board = np.array(board, dtype=int)
        else:
            game.board = np.array(board.board, dtype=int).copy()
        game.current_player = 1 if turn_number % 2 == 0 else 2
        game.last_move = last_opponent_move
        pi = self.mcts.run(game, turn_number)
        return divmod(int(np.argmax(pi)), self.board_size)
