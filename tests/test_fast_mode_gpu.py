"""GPU: the NON-PARITY fast mode (azg_config.fast_warps > 0: several warps walk one game's tree concurrently under a
virtual loss - north-star subsystem 1).  Its visit counts are by construction not the reference's, so it is checked
through what must hold anyway: bookkeeping invariants, no virtual loss left behind, tactics, and agreement of the chosen
move with the exact mode where the priors make the choice clear."""
import numpy as np
import pytest
import torch

from oracle import fakes, rules as orules

pytestmark = pytest.mark.gpu


class G:
    def __init__(self, pos):
        self.board = pos.cells.reshape(15, 15).copy()
        self.current_player = pos.player
        self.last_move = None if pos.last < 0 else divmod(pos.last, 15)
        self.move_history = [None] * pos.plies


class Gomoku:
    pass


def position(moves):
    pos = orules.Position(0)
    for mv in moves:
        assert orules.play(pos, mv)
    return pos


@pytest.mark.parametrize("warps", [1, 4, 16])
def test_fast_mode_invariants(warps):
    import alphazero_gomoku_b200 as m
    n = 400
    mcts = m.MCTS(Gomoku, n, fakes.Hashed(), add_dirichlet_noise=False, fast_warps=warps)
    pos = position([112, 113, 127])
    for ply in range(3):
        pi = mcts.run(G(pos), pos.plies)
        legal = orules.legal_mask(pos)
        assert abs(pi.sum() - 1.0) < 1e-5 and (pi[legal == 0] == 0).all() and (pi >= 0).all()
        v = mcts.last_visits
        assert (v >= 0).all() and v.sum() > 0                  # no virtual loss left behind
        if ply == 0:
            assert v.sum() <= n                                # fresh tree: every real visit counted once (later runs reuse the tree)
        st = mcts.engine.stats()
        assert st["games_in_error"] == 0
        orules.play(pos, int(np.argmax(pi)))
    # the whole tree: N >= 0 and |W| <= N everywhere once the run is over
    eng = mcts.engine
    assert st["sims"] == 3 * n
    eng.close()


class OneSpike(fakes._Base):
    """Almost all prior mass on ONE empty cell per position (chosen by a position hash): a clear choice for any search."""

    def weights(self, X):
        h = fakes.position_hash(X)[:, None]
        a = np.arange(225, dtype=np.uint64)[None, :]
        score = fakes._mix(h ^ (a * np.uint64(0x100000001B3))) % np.uint64(1 << 20) + np.uint64(1)
        empty = (X[:, 0].reshape(len(X), -1) + X[:, 1].reshape(len(X), -1)) < 0.5
        pick = np.argmax(np.where(empty, score, 0), axis=1)
        w = np.ones((len(X), 225), dtype=np.uint64)
        w[np.arange(len(X)), pick] = np.uint64(10 ** 6)
        return w


def test_fast_mode_finds_the_win():
    """Four in a row with an open end and the move: both modes find a winning move (the search only backs up terminal
    values, so a one-ply win is the tactic it can be asked for with flat priors)."""
    import alphazero_gomoku_b200 as m
    win_for_mover = position([112, 0, 113, 1, 114, 2, 115, 30])            # player 1 has 112..115 and the move: 111 or 116 wins
    for warps in (0, 8, 16):
        mcts = m.MCTS(Gomoku, 900, fakes.Uniform(), add_dirichlet_noise=False, fast_warps=warps)
        pi = mcts.run(G(win_for_mover), win_for_mover.plies)
        assert int(np.argmax(pi)) in (111, 116), (warps, int(np.argmax(pi)))
        mcts.engine.close()


def test_fast_mode_agrees_with_exact_mode_on_clear_choices():
    """With one dominant prior per position both modes choose the same move on most positions."""
    import alphazero_gomoku_b200 as m
    rng = np.random.default_rng(3)
    same = total = 0
    exact = m.MCTS(Gomoku, 300, OneSpike(), add_dirichlet_noise=False)
    fast = m.MCTS(Gomoku, 300, OneSpike(), add_dirichlet_noise=False, fast_warps=8)
    for _ in range(12):
        pos = orules.Position(0)
        for _ in range(int(rng.integers(2, 30))):
            e = np.flatnonzero(pos.cells == 0)
            q = pos.copy()
            orules.play(q, int(e[int(rng.integers(0, len(e)))]))
            if orules.game_over(q):
                break
            pos = q
        exact.clear_tree(); fast.clear_tree()
        a = int(np.argmax(exact.run(G(pos), pos.plies)))
        b = int(np.argmax(fast.run(G(pos), pos.plies)))
        same += int(a == b); total += 1
    print("fast vs exact: same move on", same, "of", total)
    assert same >= 0.75 * total
    exact.engine.close(); fast.engine.close()


def test_fast_mode_with_the_cuda_network_and_batched_games():
    """Fast mode through the device evaluator, several games per engine."""
    from alphazero_gomoku_b200.engine import SearchEngine
    from alphazero_gomoku_b200.network import PyTorchModel
    torch.manual_seed(0)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    net = model._ensure_engine()
    G_ = 8
    eng = SearchEngine(0, G_, queue_len=32, node_capacity=4096, fast_warps=8)
    probs = torch.empty((G_ * 32, 225), dtype=torch.float32, device="cuda")
    eng.begin(500)
    rounds = 0
    while True:
        n_leaves, n_more, _ = eng.fill()
        if n_leaves > 0:
            net.forward_leaves(eng, probs)
            eng.commit(probs, None)
        rounds += 1
        if n_more == 0 or rounds > 200:
            break
    pi, visits = eng.result()
    assert rounds <= 200 and eng.stats()["games_in_error"] == 0
    assert torch.allclose(pi.sum(1), torch.ones(G_, device="cuda"), atol=1e-4)
    assert bool((visits >= 0).all()) and bool((visits.sum(1) > 300).all()) and bool((visits.sum(1) <= 500).all())
    eng.close()
