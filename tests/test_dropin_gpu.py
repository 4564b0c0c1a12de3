"""GPU: the remaining drop-in surfaces - game objects, Player, checkpoint format, train loop smoke."""
import os

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rules as orules

pytestmark = pytest.mark.gpu


def test_game_objects_follow_golden_traces():
    """Gomoku / Pente shims (do_move through the CUDA rule kernels) on the reference's traces."""
    from alphazero_gomoku_b200.games import Gomoku, Pente
    z = load_golden("rules_traces.npz")
    for name in ("gomoku_overline", "gomoku_illegal", "pente_capture_kat", "pente_double", "pente_capture_win", "pente_random1"):
        rule = int(z["rules"][list(z["names"]).index(name)])
        g = (Pente if rule else Gomoku)(15)
        for i, (r, c) in enumerate(z[f"{name}/moves"]):
            ok = g.do_move((int(r), int(c)))
            assert ok == bool(z[f"{name}/oks"][i]), (name, i)
            assert np.array_equal(g.board.reshape(-1), z[f"{name}/boards"][i]), (name, i)
            assert g.current_player == z[f"{name}/players"][i]
            assert g.check_winner() == z[f"{name}/winners"][i] and g.is_game_over() == bool(z[f"{name}/overs"][i])
            if rule:
                assert [g.captures[1], g.captures[2]] == z[f"{name}/caps"][i].tolist()
        c = g.clone()
        assert np.array_equal(c.board, g.board) and c.current_player == g.current_player
        assert np.array_equal(g.get_valid_moves(), (g.board.reshape(-1) == 0).astype(np.float32))
        enc = g.get_encoded_state()
        assert enc.shape == (3, 15, 15) and np.array_equal(enc[0], (g.board == g.current_player).astype(np.float32))
        assert np.array_equal(enc[1], (g.board == 3 - g.current_player).astype(np.float32)) and (enc[2] == 1).all()
    # capture history of the Pente KAT
    p = Pente(15)
    for mv in [(7, 7), (7, 8), (0, 0), (7, 9), (7, 10)]:
        p.do_move(mv)
    assert sorted(p.capture_history[-1]) == [(7, 8), (7, 9)] and p.captures == {1: 1, 2: 0}


def test_undo_move_follows_reference_traces():
    """do_move / undo_move sequences recorded from the reference's classes (tests/golden/undo_traces.npz, oracle/make_golden.py
    build_undo): board, side to move, last move and capture counts after every operation - including pente.py:100-103,
    which puts captured stones back in the CAPTURER's colour (20 such undos in the Pente traces)."""
    from alphazero_gomoku_b200.games import Gomoku, Pente
    z = load_golden("undo_traces.npz")
    undone_captures = 0
    for tag, cls in (("g", Gomoku), ("p", Pente)):
        for k in range(3):
            g = cls(15)
            ops, oks = z[f"{tag}{k}/ops"], z[f"{tag}{k}/ok"]
            prev_caps = [0, 0]
            for i, op in enumerate(ops):
                if op < 0:
                    g.undo_move()
                else:
                    assert g.do_move((int(op) // 16, int(op) % 16)) == bool(oks[i]), (tag, k, i)
                assert np.array_equal(np.asarray(g.board), z[f"{tag}{k}/boards"][i]), (tag, k, i)
                assert g.current_player == int(z[f"{tag}{k}/players"][i]), (tag, k, i)
                last = -1 if g.last_move is None else g.last_move[0] * 15 + g.last_move[1]
                assert last == int(z[f"{tag}{k}/last"][i]), (tag, k, i)
                caps = [g.captures[1], g.captures[2]] if tag == "p" else [0, 0]
                assert caps == z[f"{tag}{k}/caps"][i].tolist(), (tag, k, i)
                if op < 0 and caps != prev_caps:
                    undone_captures += 1
                prev_caps = caps
    assert undone_captures >= 15


def test_checkpoint_roundtrip_and_player(tmp_path):
    """network.py:240-258 format: {"net","opt","board_size","action_size"}; Player loads it and moves."""
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.players import Player
    torch.manual_seed(3)
    m = PyTorchModel(board_size=15)                                   # 3x64 default like the reference
    path = os.path.join(tmp_path, "models", "snap.pt")
    m.save(path)
    state = torch.load(path, map_location="cpu")
    assert set(state) == {"net", "opt", "board_size", "action_size"} and state["action_size"] == 225
    assert "res_blocks.2.conv2.weight" in state["net"] and "policy_fc.bias" in state["net"]
    pl = Player(rules="gomoku", board_size=15, n_simulations=96, model_path=path)
    board = [[0] * 15 for _ in range(15)]
    board[7][7] = 1
    r, c = pl.play(board, 1, (7, 7))
    assert 0 <= r < 15 and 0 <= c < 15 and board[r][c] == 0
    with pytest.raises(ValueError):
        Player(rules="pente", model_path=None)
    X = np.zeros((2, 3, 15, 15), np.float32); X[:, 2] = 1
    p1, v1 = m.predict(X)
    m2 = PyTorchModel(board_size=15)
    m2.load(path)
    p2, v2 = m2.predict(X)
    assert np.array_equal(p1, p2) and np.array_equal(v1, v2)


def test_train_loop_smoke(tmp_path):
    """train_alphazero with the reference's kwargs on a tiny budget: self-play on the device driver,
    replay buffer in the reference's pickle format, Adam steps, evaluation, snapshot."""
    from alphazero_gomoku_b200 import train as tr
    best = tr.train_alphazero(game_name="gomoku", board_size=15, num_iterations=1, games_per_iteration=6, n_simulations=24,
                              buffer_size=4000, batch_size=64, epochs_per_iter=1, temp_threshold=8, eval_games=2,
                              eval_mcts_simulations=16, win_rate_threshold=0.55, cpuct=1.2, model_dir=str(tmp_path),
                              dirichlet_alpha=0.3, dirichlet_epsilon=0.25, dirichlet_n_moves=30, n_res_blocks=1, channels=64)
    files = os.listdir(tmp_path)
    assert any(f.startswith("snapshot_iter1_") for f in files) and "replay_buffer_latest.pkl" in files
    buf = tr.load_replay_buffer(os.path.join(tmp_path, "replay_buffer_latest.pkl"), 4000)
    assert len(buf) > 0 and len(buf) % 8 == 0
    s, p, z = buf.sample(16)
    assert s.shape == (16, 3, 15, 15) and p.shape == (16, 225) and z.shape == (16, 1)
    assert best.net.channels == 64


def test_batched_arena_equals_game_by_game():
    """evaluate_models on the two batched engines == the reference-shaped loop (one game at a time
    through the drop-in MCTS) for the same random openings: argmax play is deterministic."""
    import random
    from alphazero_gomoku_b200 import train as tr
    from alphazero_gomoku_b200.network import PyTorchModel
    torch.manual_seed(11)
    a = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    torch.manual_seed(12)
    b = PyTorchModel(n_res_blocks=1, channels=64, device="cuda:0")
    random.seed(5)
    batched = tr.evaluate_models(a, b, "gomoku", n_games=6, n_simulations=40, cpuct=1.0)
    random.seed(5)
    serial = tr.evaluate_models_serial(a, b, "gomoku", n_games=6, n_simulations=40, cpuct=1.0)
    assert batched == serial, (batched, serial)
    assert 0 <= batched[0] + batched[2] <= 6
    # the data-parallel arena plays contiguous slices of the same match on different ranks: slices must add up
    random.seed(5)
    stones = tr.draw_first_stones(6)
    parts = [tr.evaluate_models(a, b, "gomoku", hi - lo, 40, 1.0, first_stones=stones[lo:hi], first_game=lo)
             for lo, hi in (tr.shard_bounds(6, 4, r) for r in range(4))]
    assert sum(p[0] for p in parts) == batched[0] and sum(p[2] for p in parts) == batched[2]


@pytest.mark.parametrize("noise", [False, True])
def test_mcts_device_path_equals_host_model_path(noise):
    """MCTS.run with this package's PyTorchModel keeps the leaf batch on the GPU and syncs every few rounds;
    with the same network hidden behind a bare ``predict`` it takes the reference's host round trip per
    queue flush.  Both must give the same visit counts, move after move (tree reuse + GC included), and
    advance numpy's generator identically."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.games import Gomoku
    from alphazero_gomoku_b200.network import PyTorchModel

    torch.manual_seed(5)
    model = PyTorchModel(n_res_blocks=2, channels=64, device="cuda:0")

    class HostOnly:
        def predict(self, X):
            return model.predict(X)

    runs = []
    for evaluator in (model, HostOnly()):
        np.random.seed(7)
        mcts = m.MCTS(Gomoku, 300, evaluator, add_dirichlet_noise=noise, dirichlet_alpha=0.3, epsilon=0.25)
        game = Gomoku(15)
        seq = []
        for ply in range(5):
            pi = mcts.run(game, ply)
            seq.append((pi.copy(), mcts.last_visits.copy(), mcts.n_evals))
            a = int(np.argmax(pi))
            game.do_move((a // 15, a % 15))
        seq.append(np.random.random())
        runs.append(seq)
        mcts.engine.close()
    for (pa, va, ea), (pb, vb, eb) in zip(runs[0][:-1], runs[1][:-1]):
        assert np.array_equal(va, vb) and np.array_equal(pa, pb) and ea == eb
    assert runs[0][-1] == runs[1][-1]
