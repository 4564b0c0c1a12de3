"""GPU: BASELINE.json's full sizes (2 048 concurrent games, 800 simulations per move) checked through
size-independent properties - the oracle cannot run these sizes in seconds."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_rules_one_million_positions_properties():
    """pack -> play -> unpack on 2^20 positions: round trips are the identity, a legal move adds exactly
    one stone of the mover (Gomoku), legal mask == empties, planes partition the stones, and the
    winner flag implies game over."""
    import alphazero_gomoku_b200 as m
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(3)
    R = m.Rules(0, "cuda:0")
    r = torch.rand((n, 225), device="cuda", generator=g)
    boards = torch.zeros((n, 225), dtype=torch.int8, device="cuda")
    boards[r < 0.18] = 1
    boards[(r >= 0.18) & (r < 0.36)] = 2
    players = torch.randint(1, 3, (n,), device="cuda", generator=g, dtype=torch.int32)
    pos = R.pack(boards, players)
    b2, p2, lasts, caps, plies = R.unpack(pos)
    assert torch.equal(b2, boards) and torch.equal(p2, players) and bool((lasts == -1).all())
    legal = R.legal(pos)
    assert torch.equal(legal, (boards == 0).float())
    enc = R.encode(pos)
    me = players.view(-1, 1).to(torch.int8)
    assert torch.equal(enc[:, 0].reshape(n, 225), (boards == me).float())
    assert torch.equal(enc[:, 1].reshape(n, 225), ((boards != 0) & (boards != me)).float())
    assert bool((enc[:, 2] == 1).all())
    acts = torch.multinomial(legal, 1, generator=g).squeeze(1).to(torch.int32)
    st = R.play(pos, acts)
    b3, p3, l3, _, pl3 = R.unpack(pos)
    assert bool(((st & 8) == 0).all()) and torch.equal(l3, acts) and torch.equal(p3, 3 - players) and bool((pl3 == 1).all())
    diff = (b3 != boards)
    assert bool((diff.sum(1) == 1).all())
    assert torch.equal(b3[torch.arange(n, device="cuda"), acts.long()].to(torch.int32), players)
    assert bool((((st & 3) == 0) | ((st & 4) != 0)).all())
    # a second move on the same cell is rejected and changes nothing
    st2 = R.play(pos, acts)
    b4, p4, _, _, _ = R.unpack(pos)
    assert bool(((st2 & 8) != 0).all()) and torch.equal(b4, b3) and torch.equal(p4, p3)


@pytest.mark.parametrize("rule", [0, 1])
def test_selfplay_full_size_properties(rule):
    """2 048 games x 800 simulations, two plies, 6x128 network (configs[1] / configs[2])."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(0)
    model = PyTorchModel(n_res_blocks=6, channels=128, device="cuda:0")
    G, S = 2048, 800
    sp = SelfPlay(model, rule=rule, n_games=G, n_sims=S, node_capacity=8192, example_capacity=1 << 16, seed=7)
    for step in range(2):
        before = sp.engine.roots().clone()
        evals0 = sp.total_evals
        sp.step()
        pi = sp.last_pi
        # pi is a distribution over legal moves of the position it was computed for
        assert torch.allclose(pi.sum(1), torch.ones(G, device="cuda"), atol=1e-4)
        legal = sp.engine.rules.legal(before)
        assert bool((pi[legal == 0] == 0).all())
        # the chosen move was legal and had a visit (or pi was the uniform fallback)
        a = sp.actions.long()
        assert bool((legal[torch.arange(G, device="cuda"), a] == 1).all())
        # reference accounting: every simulation queues at most one leaf plus one per mid-run flush
        evals = sp.total_evals - evals0
        assert G * S * 0.5 < evals <= G * (S + S // 32 + 1)
        boards, players, lasts, caps, plies = sp.engine.rules.unpack(sp.engine.roots())
        alive = (sp.done == 0)
        assert bool(((boards != 0).sum(1) == plies - 2 * caps.sum(1))[alive].all())
        assert bool((plies[alive] == step + 1).all())
    st = sp.engine.stats()
    assert st["games_in_error"] == 0 and st["dropped_trees"] == 0 and st["sims"] == 2 * G * S
    assert 1.0 <= st["visits"] / st["sims"] <= 226
    sp.close()


def test_selfplay_is_deterministic():
    """Same seeds -> identical visit distributions and moves, run to run (integer atomics only)."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(4)
    model = PyTorchModel(n_res_blocks=2, channels=64, device="cuda:0")
    outs = []
    for _ in range(2):
        sp = SelfPlay(model, n_games=256, n_sims=200, node_capacity=4096, example_capacity=1 << 14, seed=21)
        acts, pis = [], []
        for _ in range(4):
            sp.step()
            acts.append(sp.actions.clone())
            pis.append(sp.last_pi.clone())
        outs.append((torch.stack(acts), torch.stack(pis)))
        sp.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("rule", [0, 1])
def test_full_size_step_sampled_games_replayed_through_oracle(rule):
    """A REAL configs[1] / configs[2] step - 2 048 concurrent games x 800 simulations, 6x128 network, root noise,
    games scattered over plies 0..39 as in bench.py - with a sample of its games replayed one by one through
    the oracle search (oracle.search.Search, the restatement of mcts/new_mcts_alpha.py).  The oracle is fed
    the same network outputs (CUDA evaluator through its numpy API; a position's output does not depend on the
    batch it travels in) and the same Dirichlet draws (read back from the device), so the visit counts of
    every sampled game must be IDENTICAL, not just plausible."""
    import bench
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    from oracle import rules as orules
    from oracle.search import Search
    torch.manual_seed(0)
    model = PyTorchModel(n_res_blocks=6, channels=128, device="cuda:0")
    G, S = 2048, 800
    sp = SelfPlay(model, rule=rule, n_games=G, n_sims=S, node_capacity=8192, noise=True, alpha=0.05, eps=0.15, noise_plies=10,
                  example_capacity=1 << 16, seed=12345)
    pos = bench.scatter_start(sp, 777)
    pi, visits = sp.search()
    pi, visits, noise = pi.cpu().numpy(), visits.cpu().numpy(), sp.noise.cpu().numpy()
    boards, players, lasts, caps, plies = (t.cpu().numpy() for t in sp.engine.rules.unpack(pos))
    assert sp.engine.stats()["games_in_error"] == 0

    class DeviceNet:
        def predict(self, X):
            return model.predict(X)

    sample = [0, 43, 126, 369, 1010, 1297, 1953, 2047]          # root plies 0, 3, 6, 9, 10, 17, 33, 7: noised and plain roots
    for g in sample:
        p = orules.Position(rule)
        p.cells = boards[g].astype(np.int8).copy()
        p.player, p.last, p.caps, p.plies = int(players[g]), int(lasts[g]), [int(caps[g][0]), int(caps[g][1])], int(plies[g])
        s = Search(rule, S, DeviceNet(), cpuct=1.0, queue_len=32, alpha=0.05, eps=0.15, noise_plies=10, noise=True,
                   noise_fn=lambda n, g=g: noise[g].copy())
        want = s.run(p, p.plies)
        assert np.array_equal(visits[g], s.Nv[p.key()].astype(np.int32)), (g, p.plies)
        assert np.array_equal(pi[g], want), (g, p.plies)
    sp.close()
