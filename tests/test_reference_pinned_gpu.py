"""GPU: rows a17 / f2 / f3 / a7 of SURVEY section 8 against fixtures WRITTEN BY THE REFERENCE
(oracle/make_golden.py build_game / build_arena / build_train) and against the oracle on whole games."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden
from oracle import fakes, rules as orules, selfplay as oselfplay
from oracle.search import Search

pytestmark = pytest.mark.gpu


def unpack_planes(bits):
    return np.unpackbits(bits, axis=1)[:, :675].astype(np.float32).reshape(-1, 3, 15, 15)


# ------------------------------------------------------------------------------------------------ a17
@pytest.mark.parametrize("name", ["g_t0", "p_t0", "g_cut", "g_temp_noise"])
def test_host_play_game_and_collect_equals_reference(name):
    """train.play_game_and_collect over the drop-in MCTS and game objects reproduces the reference's whole
    game (train.py:360-412) example for example: planes, pi, z, symmetry order, winner - including the
    temperature schedule and root noise drawn from numpy's seeded generator."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200 import games, train as tr
    z = load_golden("selfplay_games.npz")
    rule, n_sims, max_moves, thr, syms, noise, winner = (int(v) for v in z[f"{name}/cfg"])
    temp_fn = (lambda mn: 0.0) if thr < 0 else (lambda mn: max(0.0, 1.0 - mn / thr))
    cls = games.Pente if rule else games.Gomoku
    mcts = m.MCTS(cls, n_sims, fakes.BY_NAME[str(z[f"{name}/model"][0])](), cpuct=1.0, dirichlet_alpha=0.3, epsilon=0.25,
                  apply_dirichlet_n_first_moves=6, add_dirichlet_noise=bool(noise))
    game = cls(15)
    np.random.seed(int(z["seed"][0]))
    rows, won = tr.play_game_and_collect(mcts, game, temp_fn, max_moves=max_moves, use_symmetries=bool(syms))
    assert won == winner and len(rows) == len(z[f"{name}/z"])
    assert [r * 15 + c for r, c in game.move_history] == z[f"{name}/moves"].tolist()
    assert all(r[0].dtype == np.float32 and r[1].dtype == np.float32 for r in rows)
    assert np.array_equal(np.stack([r[0] for r in rows]), unpack_planes(z[f"{name}/planes_bits"]))
    assert np.array_equal(np.stack([r[1] for r in rows]), z[f"{name}/pi"])
    assert np.array_equal(np.array([r[2] for r in rows], dtype=np.float32), z[f"{name}/z"])
    mcts.engine.close()


@pytest.mark.parametrize("rule,max_moves", [(0, 225), (1, 225), (0, 9)])
def test_device_selfplay_whole_games_equal_oracle(rule, max_moves):
    """The batched device driver (SelfPlay: search, move choice, game end, labels, 8 symmetries, restart) against
    oracle.selfplay.play_one game by game.  The oracle searches with the SAME network outputs (the CUDA
    evaluator through its numpy API; outputs do not depend on the batch shape) and is handed the one
    random quantity of each game - the ply-0 move, sampled at temperature 1 by the device's Philox stream;
    from ply 1 on the temperature is 0, so every later move, every example row and the winner must agree."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(8)
    model = PyTorchModel(n_res_blocks=2, channels=64, device="cuda:0")
    G, S, THR = 8, 64, 1e-9
    sp = SelfPlay(model, rule=rule, n_games=G, n_sims=S, noise=False, temp_threshold=THR, max_moves=max_moves,
                  node_capacity=8192, example_capacity=1 << 16, seed=5)
    first_move = [None] * G
    pending = set(range(G))          # slots whose FIRST game is still running (restarted slots keep playing; ignored)
    blocks = {}
    consumed = 0

    class DeviceNet:
        def predict(self, X):
            return model.predict(X)

    def oracle_game(g):
        s = Search(rule, S, DeviceNet(), cpuct=1.0, queue_len=32, noise=False)
        rows, won = oselfplay.play_one(s, orules.Position(rule), lambda ply: max(0.0, 1.0 - ply / THR), max_plies=max_moves,
                                       expand=True, choice=lambda n, p: first_move[g])
        flat = np.concatenate([np.stack([r[0] for r in rows]).reshape(len(rows), -1), np.stack([r[1] for r in rows]),
                               np.array([[r[2]] for r in rows], np.float32)], axis=1)
        return flat, won

    for step in range(max_moves + 2):
        if not pending:
            break
        sp.step()
        acts, done = sp.actions.cpu().numpy(), sp.done.cpu().numpy()
        for g in range(G):
            if first_move[g] is None:
                first_move[g] = int(acts[g])
        n = sp.n_examples()
        region = sp.examples[consumed:n].cpu().numpy()      # the rows of the games that ended in this step, one block per game
        consumed = n
        for g in [g for g in range(G) if done[g] and g in pending]:
            want, won = oracle_game(g)
            assert int(sp.winners[g].item()) == won, (g, step)
            assert len(want) == 8 * (step + 1), (g, step, len(want))            # same game length
            hits = [off for off in range(0, len(region) - len(want) + 1, 8) if np.array_equal(region[off:off + len(want)], want)]
            assert hits, f"game {g} (ended at step {step}): its example rows differ from the oracle's"
            blocks[g] = len(want)
            pending.discard(g)
    assert not pending and len(blocks) == G
    assert all(ln % 8 == 0 and ln // 8 <= max_moves for ln in blocks.values())
    sp.close()


# ------------------------------------------------------------------------------------------------ a7
def test_clear_tree_equals_fresh_tree():
    """MCTS.clear_tree (new_mcts_alpha.py:58-72): after clearing, a run at any position gives the visit
    counts of a brand-new tree (oracle, fresh Search), not those of the reused one."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200 import games
    mcts = m.MCTS(games.Gomoku, 150, fakes.Hashed(), add_dirichlet_noise=False)
    game = games.Gomoku(15)
    pos = orules.Position(0)
    reused = Search(0, 150, fakes.Hashed(), noise=False)
    for ply in range(4):
        pi = mcts.run(game, ply)
        assert np.array_equal(pi, reused.run(pos, ply))
        a = int(np.argmax(pi))
        game.do_move(divmod(a, 15))
        orules.play(pos, a)
    with_reuse = reused.run(pos, 4)
    fresh = Search(0, 150, fakes.Hashed(), noise=False).run(pos, 4)
    assert not np.array_equal(with_reuse, fresh)            # the case distinguishes the two
    mcts.clear_tree()
    assert np.array_equal(mcts.run(game, 4), fresh)
    assert mcts.engine.stats()["live_nodes"] <= 150 + 150 // 32 + 2
    mcts.engine.close()


# ------------------------------------------------------------------------------------------------ f2
class ArenaFake:
    def __init__(self, name):
        self.inner = fakes.BY_NAME[name]()
        self.board_size = 15

    def predict(self, X):
        return self.inner.predict(X)


@pytest.mark.parametrize("name", ["hashed_vs_spiky", "spiky_vs_hashed"])
def test_batched_arena_equals_reference_transcript(name):
    """evaluate_models (all games in lock step on two batched engines) against the transcript of the
    REFERENCE's evaluate_models (train.py:418-486) between the same two injected-prior models with the same
    ``random`` seed: identical opening stones, identical move lists, identical (wins, win rate, draws)."""
    from alphazero_gomoku_b200 import train as tr
    z = load_golden("arena_transcripts.npz")
    n_games, n_sims = (int(v) for v in z[f"{name}/cfg"])
    a, b = (str(x) for x in z[f"{name}/models"])
    random.seed(int(z["seed"][0]))
    log = []
    wins, rate, draws = tr.evaluate_models(ArenaFake(a), ArenaFake(b), "gomoku", n_games=n_games, n_simulations=n_sims,
                                           cpuct=float(z[f"{name}/cpuct"][0]), transcript=log)
    assert [wins, draws] == z[f"{name}/result"].tolist() and rate == float(z[f"{name}/win_rate"][0])
    want = [[int(x) for x in row if x >= 0] for row in z[f"{name}/moves"]]
    assert log == want


# ------------------------------------------------------------------------------------------------ f3
def test_reference_checkpoints_load_and_predict():
    """Checkpoints written by the reference's PyTorchModel.save (network.py:240-248): architecture inferred,
    weights + optimiser state loaded, CUDA predict within the stated bf16 tolerance of the reference's own
    fp32 outputs for those weights."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from oracle import net as onet
    z = load_golden("net_outputs.npz")
    m = PyTorchModel.from_checkpoint(os.path.join(GOLDEN, "ref_checkpoint_3x64_seed0.pt"), device="cuda:0")
    assert len(m.net.res_blocks) == 3 and m.net.channels == 64
    probs, values = m.predict(z["X"])
    kl = onet.policy_kl(z["3x64/probs"], probs)
    assert kl.mean() < 2e-4 and kl.max() < 1.2e-3 and np.abs(values - z["3x64/values"]).max() < 0.025
    # the model the reference trained for three steps: BatchNorm running statistics and Adam moments included
    t = load_golden("train_steps.npz")
    m2 = PyTorchModel(n_res_blocks=2, channels=64, device="cuda:0")
    path = os.path.join(GOLDEN, "ref_train_2x64_after3.pt")
    m2.load(path)
    ref = torch.load(path, map_location="cpu")
    st = m2.optimizer.state_dict()["state"]
    assert len(st) == len(ref["opt"]["state"]) > 0
    for i, s in ref["opt"]["state"].items():
        assert torch.equal(st[i]["exp_avg"].cpu(), s["exp_avg"]) and int(st[i]["step"]) == 3
    X = unpack_planes(t["batch/planes_bits"])[:8]
    probs, values = m2.predict(X)
    kl = onet.policy_kl(t["2x64/probs_after"], probs)
    assert kl.max() < 1.2e-3 and np.abs(values - t["2x64/values_after"]).max() < 0.025, (kl.max(), np.abs(values - t["2x64/values_after"]).max())
