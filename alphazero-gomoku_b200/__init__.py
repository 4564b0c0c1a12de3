"""azgomoku_b200 - B200-native batched AlphaZero self-play engine for Gomoku / Pente.

Drop-in for the hot path of shirongcan/AlphaZero-Gomoku:
``MCTS`` (mcts/new_mcts_alpha.py), ``PyTorchModel`` (network.py), the game rules
(games/gomoku.py, games/pente.py) and the self-play driver (train.py:360-412),
all executed by hand-written sm_100a CUDA behind the C ABI in
``include/azgomoku_b200.h``.  Import name: ``alphazero_gomoku_b200`` (the
directory is ``alphazero-gomoku_b200``; the top-level alias module maps it).
"""
from ._lib import AzgError, LIB_PATH, lib  # noqa: F401  (loads the CUDA library or raises)
from .engine import GOMOKU, PENTE, Rules, SearchEngine, rule_of  # noqa: F401
from .mcts import MCTS  # noqa: F401

__all__ = ["AzgError", "LIB_PATH", "lib", "GOMOKU", "PENTE", "Rules", "SearchEngine", "rule_of", "MCTS"]
