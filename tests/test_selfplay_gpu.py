"""GPU: self-play driver kernels - Philox Dirichlet noise and temperature sampling (statistical),
argmax at T = 0, outcome labels and the 8 symmetries in the reference's order (exact)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rules as orules
from oracle.search import dihedral8

pytestmark = pytest.mark.gpu


def test_dirichlet_moments():
    """Marginals of Dirichlet(alpha * 1_225): mean 1/225, var = (1/225)(1-1/225)/(225*alpha+1)."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200._lib import check, lib, ptr
    G, alpha = 4096, 0.05
    eng = m.SearchEngine(0, G, noise=True, alpha=alpha, eps=0.25, node_capacity=64)
    noise = torch.empty((G, 225), dtype=torch.float64, device="cuda")
    draws = []
    for d in range(4):
        eng._sync_stream()
        check(lib.azg_selfplay_noise(eng._h, d, ptr(noise)))
        draws.append(noise.cpu().numpy().copy())
    x = np.concatenate(draws)                         # 16384 samples of a 225-vector
    assert np.allclose(x.sum(1), 1.0, atol=1e-12) and (x >= 0).all()
    mean, var = 1 / 225, (1 / 225) * (1 - 1 / 225) / (225 * alpha + 1)
    assert abs(x.mean() - mean) < 1e-9
    assert np.abs(x.mean(0) - mean).max() < 6 * np.sqrt(var / len(x))
    assert abs(x.var(0).mean() / var - 1) < 0.03
    assert not np.array_equal(draws[0], draws[1])
    # same (seed, draw) -> same numbers
    eng._sync_stream()
    check(lib.azg_selfplay_noise(eng._h, 0, ptr(noise)))
    assert np.array_equal(noise.cpu().numpy(), draws[0])
    eng.close()


def test_choose_argmax_and_sampling():
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200._lib import check, lib, ptr
    G = 8192
    eng = m.SearchEngine(0, G, node_capacity=64)
    check(lib.azg_selfplay_enable(eng._h, 225))
    pi = torch.zeros((G, 225), dtype=torch.float32, device="cuda")
    pi[:, 10] = 0.5; pi[:, 20] = 0.3; pi[:, 200] = 0.2
    acts = torch.empty(G, dtype=torch.int32, device="cuda")
    # ply 0, threshold 10 -> temperature 1: frequencies follow pi
    eng._sync_stream()
    check(lib.azg_selfplay_choose(eng._h, ptr(pi), C.c_float(10.0), 1, ptr(acts)))
    a = acts.cpu().numpy()
    f = np.array([(a == 10).mean(), (a == 20).mean(), (a == 200).mean()])
    assert f.sum() == 1.0 and np.abs(f - [0.5, 0.3, 0.2]).max() < 0.02
    # ply 1 with threshold 1 -> temperature 0 -> first argmax even with ties
    pi2 = torch.zeros((G, 225), dtype=torch.float32, device="cuda")
    pi2[:, 37] = 0.4; pi2[:, 99] = 0.4; pi2[:, 5] = 0.2
    check(lib.azg_selfplay_choose(eng._h, ptr(pi2), C.c_float(1.0), 2, ptr(acts)))
    assert (acts.cpu().numpy() == 37).all()
    eng.close()


def test_examples_labels_and_symmetries():
    """Play two short scripted games to the end; exported rows must equal the oracle's
    play_game_and_collect output format: 8 symmetries per ply in reference order, z per player."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200._lib import check, lib, ptr
    G = 2
    eng = m.SearchEngine(0, G, node_capacity=64)
    check(lib.azg_selfplay_enable(eng._h, 225))
    # game 0: player 1 wins on the top row; game 1: same moves shifted down (still player 1 wins)
    moves0 = [0, 30, 1, 31, 2, 32, 3, 33, 4]
    moves1 = [m_ + 45 for m_ in moves0]
    rng = np.random.default_rng(0)
    pis, pos = [], [orules.Position(0), orules.Position(0)]
    out = torch.zeros((4096, 901), dtype=torch.float32, device="cuda")
    cursor = torch.zeros(1, dtype=torch.int64, device="cuda")
    done = torch.zeros(G, dtype=torch.int32, device="cuda")
    winners = torch.zeros(G, dtype=torch.int32, device="cuda")
    expected = [[], []]
    for t in range(len(moves0)):
        pi = np.zeros((G, 225), np.float32)
        for g, mv in enumerate((moves0[t], moves1[t])):
            pi[g] = rng.random(225).astype(np.float32) * 0.001
            pi[g, mv] = 5.0                                   # argmax = the scripted move
            pi[g] /= pi[g].sum()
            expected[g].append((orules.encode(pos[g]), pi[g].copy(), pos[g].player))
            orules.play(pos[g], mv)
        acts = torch.empty(G, dtype=torch.int32, device="cuda")
        eng._sync_stream()
        check(lib.azg_selfplay_choose(eng._h, ptr(torch.from_numpy(pi).cuda()), C.c_float(1e-9), 7, ptr(acts)))   # T = 0 after ply 0
        if t > 0:
            assert acts.cpu().tolist() == [moves0[t], moves1[t]]
        else:
            acts = torch.tensor([moves0[0], moves1[0]], dtype=torch.int32, device="cuda")
        status = eng.advance(acts, gc=True)
        check(lib.azg_selfplay_finish(eng._h, ptr(status), 225, 1, ptr(out), 4096, ptr(cursor), ptr(done), ptr(winners)))
    assert done.cpu().tolist() == [1, 1] and winners.cpu().tolist() == [1, 1]
    n = int(cursor.item())
    assert n == 2 * len(moves0) * 8
    rows = out[:n].cpu().numpy()
    # rows of a game are contiguous (one atomic reservation per game, in either order); the first
    # ply's pi tells the two games apart (their first positions are both the empty board)
    blocks = [rows[:n // 2], rows[n // 2:]]
    if not np.array_equal(blocks[0][0, 675:900], expected[0][0][1]):
        blocks = blocks[::-1]
    for g in range(G):
        k = 0
        for planes, pi, player in expected[g]:
            z = 1.0 if player == 1 else -1.0
            for s, p in dihedral8(planes, pi):
                assert np.array_equal(blocks[g][k, :675].reshape(3, 15, 15), s), (g, k)
                assert np.array_equal(blocks[g][k, 675:900], p)
                assert blocks[g][k, 900] == z
                k += 1
    eng.close()


def test_selfplay_runs_games_to_completion():
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(0)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    sp = SelfPlay(model, n_games=32, n_sims=48, node_capacity=2048, example_capacity=1 << 16, max_moves=60)
    finished = 0
    for _ in range(70):
        sp.step()
        finished += int(sp.done.sum().item())
    st = sp.engine.stats()
    assert st["games_in_error"] == 0 and finished >= 32
    rows = sp.drain_examples()
    states, pis, zs = SelfPlay.split(rows)
    assert rows.shape[0] > 0 and rows.shape[0] % 8 == 0
    assert torch.all((zs == 0) | (zs == 1) | (zs == -1))
    assert torch.allclose(pis.sum(1), torch.ones_like(pis[:, 0]), atol=1e-4)
    assert torch.all(states[:, 2] == 1.0) and torch.all((states[:, 0] * states[:, 1]) == 0)
    sp.close()


def test_pente_selfplay_invariants():
    """BASELINE config 3 (Pente batched self-play) at test size: size-independent properties that
    must hold at every ply of every game - stones on the board = plies - 2 * captured pairs,
    capture counters < 5 while the game runs, the side to move alternates."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(1)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    sp = SelfPlay(model, rule=1, n_games=128, n_sims=40, node_capacity=4096, example_capacity=1 << 17, temp_threshold=30.0)
    finished = 0
    for step in range(90):
        sp.step()
        boards, players, lasts, caps, plies = [x.cpu().numpy() for x in sp.engine.rules.unpack(sp.engine.roots())]
        stones = (boards != 0).sum(1)
        assert np.array_equal(stones, plies - 2 * caps.sum(1)), step
        assert (caps < 5).all() and ((players == 1) | (players == 2)).all()
        assert np.array_equal(players, 1 + (plies % 2))                # player 1 moves on even plies
        finished += int(sp.done.sum().item())
    st = sp.engine.stats()
    assert st["games_in_error"] == 0 and st["dropped_trees"] == 0
    rows = sp.drain_examples()
    assert rows.shape[0] % 8 == 0
    sp.close()


def test_results_do_not_depend_on_sharding():
    """games are independent and the on-device RNG is keyed by the global game id, so
    16 games in one engine == the same games split over two engines of 8 (as two ranks would)."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(2)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    kw = dict(n_sims=64, node_capacity=2048, example_capacity=1 << 14, seed=99, noise=True, alpha=0.3, eps=0.25)
    whole = SelfPlay(model, n_games=16, game_base=0, **kw)
    lo = SelfPlay(model, n_games=8, game_base=0, **kw)
    hi = SelfPlay(model, n_games=8, game_base=8, **kw)
    for step in range(12):
        whole.step(); lo.step(); hi.step()
        a = whole.actions.cpu().numpy()
        assert np.array_equal(a[:8], lo.actions.cpu().numpy()), step
        assert np.array_equal(a[8:], hi.actions.cpu().numpy()), step
        assert torch.equal(whole.last_pi[:8], lo.last_pi) and torch.equal(whole.last_pi[8:], hi.last_pi)
    assert len(set(whole.actions.cpu().tolist())) > 1, "games must have diverged (noise + sampling)"
    for s in (whole, lo, hi):
        s.close()


def test_pipelined_driver_matches_plain_driver():
    """Two game groups on two streams (fill of one group overlapping the other's leaf evaluation) must
    produce exactly the moves and visit distributions of the single-group driver."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import PipelinedSelfPlay, SelfPlay
    torch.manual_seed(6)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    kw = dict(n_sims=96, node_capacity=2048, example_capacity=1 << 14, seed=5, noise=True, alpha=0.3, eps=0.25)
    plain = SelfPlay(model, n_games=32, **kw)
    piped = PipelinedSelfPlay(model, n_games=32, **kw)
    for step in range(10):
        plain.step()
        piped.step()
        torch.cuda.synchronize()
        assert torch.equal(plain.actions, piped.actions), step
        assert torch.equal(plain.last_pi, piped.last_pi), step
    c = piped.counters()
    assert c["total_sims"] == plain.total_sims and c["total_evals"] == plain.total_evals
    assert piped.stats()["games_in_error"] == 0
    plain.close()
    piped.close()


def test_graph_captured_ply_matches_host_driven_loop():
    """One CUDA graph per ply (fixed number of rounds, no host sync) == the host-driven loop."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(8)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    kw = dict(n_sims=100, node_capacity=2048, example_capacity=1 << 14, seed=3, noise=True, alpha=0.3, eps=0.25, max_moves=40)
    plain = SelfPlay(model, n_games=24, **kw)
    graph = SelfPlay(model, n_games=24, **kw)
    graph.enable_graph()
    for step in range(45):                      # long enough for games to end and restart (max_moves 40)
        plain.step()
        graph.step()
        torch.cuda.synchronize()
        assert torch.equal(plain.actions, graph.actions), step
        assert torch.equal(plain.last_pi, graph.last_pi), step
        assert torch.equal(plain.done, graph.done), step
    assert plain.n_examples() == graph.n_examples() > 0
    assert graph.engine.stats()["evals"] == plain.engine.stats()["evals"]
    plain.close()
    graph.close()


def test_max_games_plays_exactly_that_many_complete_games():
    """SelfPlay(max_games=n): exactly n games are started, every one is played to its end and counted once
    (train.py:671-694 plays games_per_iteration full games) - finished slots restart only while games remain
    to be started, then retire; no game is cut off, none is counted twice."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(2)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    for G, n, use_graph in ((4, 7, False), (8, 5, False), (4, 6, True)):
        sp = SelfPlay(model, n_games=G, n_sims=32, noise=True, max_moves=225, node_capacity=2048, example_capacity=1 << 16,
                      seed=3, max_games=n)
        if use_graph:
            sp.enable_graph()
        finished, steps, rows_by_len = 0, 0, 0
        lengths = torch.zeros(G, dtype=torch.int64, device="cuda")
        while sp.games_running() > 0:
            active_before = sp.active.clone()
            sp.step()
            steps += 1
            lengths += active_before
            done = sp.done.bool()
            assert bool((active_before[done] == 1).all())              # only running slots finish
            rows_by_len += int(lengths[done].sum().item()) * 8
            lengths[done] = 0
            finished += int(done.sum().item())
            assert steps < 2000
        assert finished == n and int(sp.started.item()) == n
        assert sp.n_examples() == rows_by_len                          # every finished game contributed all its plies
        before = sp.n_examples()
        sp.step()                                                      # all retired: nothing moves, nothing is emitted
        assert sp.n_examples() == before and int(sp.done.sum().item()) == 0
        sp.close()


def test_predict_rejects_inputs_that_are_not_encoded_boards():
    from alphazero_gomoku_b200.network import PyTorchModel
    torch.manual_seed(1)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    X = np.zeros((2, 3, 15, 15), np.float32)
    X[:, 2] = 1
    X[0, 0, 7, 7] = 1
    model.predict(X)
    for bad in ("frac", "overlap", "turn"):
        Y = X.copy()
        if bad == "frac":
            Y[1, 1, 3, 3] = 0.5
        elif bad == "overlap":
            Y[0, 1, 7, 7] = 1
        else:
            Y[1, 2, 0, 0] = 0
        with pytest.raises(ValueError):
            model.predict(Y)
    with pytest.raises(ValueError):
        PyTorchModel(action_size=226, device="cuda:0")
    # a write through .data is invisible to the version counters: invalidate() makes it count
    p0, _ = model.predict(X)
    model.net.policy_fc.bias.data[5] += 3.0
    model.invalidate()
    p1, _ = model.predict(X)
    assert not np.array_equal(p0, p1)


def test_graph_mode_round_bound_small_queue():
    """The captured ply must hold enough rounds for n*q/(q-1) leaves (every flush parks one simulation that
    queues a second leaf): graph == host-driven loop also for a short queue, and queue_len 1 is refused."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    torch.manual_seed(6)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    outs = []
    for use_graph in (False, True):
        sp = SelfPlay(model, n_games=16, n_sims=50, queue_len=3, node_capacity=2048, example_capacity=1 << 14, seed=9)
        if use_graph:
            sp.enable_graph()
        pis = []
        for _ in range(3):
            sp.step()
            pis.append(sp.last_pi.clone())
        outs.append(torch.stack(pis))
        assert sp.engine.stats()["games_in_error"] == 0
        sp.close()
    assert torch.equal(outs[0], outs[1])
    sp = SelfPlay(model, n_games=2, n_sims=8, queue_len=1, node_capacity=256, example_capacity=1 << 10)
    with pytest.raises(ValueError):
        sp.enable_graph()
    sp.close()


def test_packed_examples_expand_to_the_same_rows():
    """packed_examples mode (one 976-byte record per ply, the format ranks exchange) + azg_examples_expand ==
    the rows the direct mode writes (8 symmetries per ply in the reference's order, labels included)."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import PACKED_WORDS, SelfPlay, expand_examples
    torch.manual_seed(2)
    model = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    rows = []
    for packed in (False, True):
        sp = SelfPlay(model, n_games=16, n_sims=32, noise=True, max_moves=60, node_capacity=2048,
                      example_capacity=1 << 15, seed=11, max_games=16, packed_examples=packed)
        while sp.games_running() > 0:
            sp.step()
        if packed:
            p = sp.drain_packed()
            assert p.dtype == torch.int32 and p.shape[1] == PACKED_WORDS and p.shape[0] * 8 == rows[0].shape[0]
            assert p.shape[0] * PACKED_WORDS * 4 * 29 < rows[0].size * 4                 # > 29x fewer bytes
            rows.append(expand_examples(p, True).cpu().numpy())
            one = expand_examples(p[:3], False).cpu().numpy()                            # without symmetries: the identity image
            assert one.shape == (3, 901) and np.array_equal(one, rows[1][0:24:8])
        else:
            rows.append(sp.drain_examples().cpu().numpy())
        sp.close()
    a, b = (np.unique(r, axis=0) for r in rows)            # game blocks land in reservation order: compare as sorted sets
    assert rows[0].shape == rows[1].shape and np.array_equal(a, b)
