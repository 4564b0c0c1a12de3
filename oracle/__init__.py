"""CPU oracle for the B200 self-play engine.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker or the timed CPU baseline.  The product path
(``alphazero-gomoku_b200/``) never imports this package and fails loudly when
its CUDA library is missing.

The modules restate, in numpy / torch-CPU, the algorithm of the reference
(shirongcan/AlphaZero-Gomoku) for the one hot path this repository replaces:

* ``rules``  - Gomoku / Pente board rules (games/gomoku.py, games/pente.py)
* ``search`` - deferred-evaluation PUCT search (mcts/new_mcts_alpha.py)
* ``net``    - fp32 policy/value ResNet forward (network.py)
* ``selfplay`` - one self-play game + symmetry expansion (train.py:252-266, 360-412)

Pinning: the reference ships no tests or golden vectors ("parity unpinned" by
the reference's own suite).  The oracle is instead pinned against OUTPUTS OF THE
REFERENCE ITSELF, produced in the build container by importing
``/root/reference`` unmodified: ``oracle/make_golden.py`` generated the
fixtures under ``tests/golden/`` and asserted, while generating, that every
restated function agrees with the reference bit for bit on the same inputs.
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures
without needing the reference.
"""
