"""CPU: host-side logic that needs no GPU - argument mapping of the drop-in classes, replay buffers,
sampling helpers, checkpoint architecture inference."""
import numpy as np
import pytest
import torch

from conftest import load_golden


def test_rule_mapping_and_game_fields():
    from alphazero_gomoku_b200.engine import rule_of, GOMOKU, PENTE
    from alphazero_gomoku_b200.mcts import game_fields

    class Gomoku:
        pass

    class Pente:
        pass

    assert rule_of(Gomoku) == GOMOKU and rule_of(Pente) == PENTE and rule_of("pente") == PENTE and rule_of(Gomoku()) == GOMOKU
    with pytest.raises(ValueError):
        rule_of("chess")

    class G:
        board = np.zeros((15, 15), dtype=int)          # int64 board as players/player_alpha.py:60 builds it
        current_player = 2
        last_move = (7, 8)
        move_history = [(7, 8)]
        captures = {1: 2, 2: 1}
    G.board[7, 8] = 1
    board, player, last, caps, plies = game_fields(G())
    assert board.dtype == np.int8 and board[7 * 15 + 8] == 1 and player == 2 and last == 113 and caps == (2, 1) and plies == 1
    G.last_move = None
    assert game_fields(G())[2] == -1


def test_sampling_helpers_match_reference_fixture():
    from alphazero_gomoku_b200 import train as tr
    z = load_golden("symmetry_sampling.npz")
    for t, want in zip(z["temps"], z["tempered"]):
        assert np.array_equal(np.asarray(tr.softmax_temperature(z["pi"], float(t)), dtype=np.float64), want)
    assert tr.sample_action_from_pi(z["pi"], 0) == int(np.argmax(z["pi"]))
    np.random.seed(0)
    a = [tr.sample_action_from_pi(z["pi"], 1.0) for _ in range(50)]
    assert all(0 <= x < 225 for x in a) and len(set(a)) > 5


def test_replay_buffers(tmp_path):
    from alphazero_gomoku_b200 import train as tr
    rows = torch.arange(12 * 901, dtype=torch.float32).reshape(12, 901)
    host = tr.ReplayBuffer(capacity=10)
    host.add_rows(rows)
    assert len(host) == 10 and host.buffer[0][0].shape == (3, 15, 15) and host.buffer[0][1].shape == (225,)
    s, p, zz = host.sample(4)
    assert s.shape == (4, 3, 15, 15) and p.shape == (4, 225) and zz.shape == (4, 1)
    path = str(tmp_path / "buf.pkl")
    assert tr.save_replay_buffer(host, path)
    back = tr.load_replay_buffer(path, 10)
    assert len(back) == 10 and np.array_equal(back.buffer[3][1], host.buffer[3][1])
    import pickle
    raw = pickle.load(open(path, "rb"))
    assert set(raw) == {"buffer", "capacity"} and raw["capacity"] == 10          # the reference's dictionary (train.py:309-312)
    dev = tr.DeviceReplayBuffer(10, "cpu")
    dev.add_rows(rows[:7])
    dev.add_rows(rows[7:])
    assert len(dev) == 10
    h2 = dev.to_host()
    assert [float(x[2]) for x in h2.buffer] == [float(x[2]) for x in host.buffer]      # oldest-first, same eviction as the deque
    d2 = tr.DeviceReplayBuffer.from_host(host, "cpu")
    assert len(d2) == 10 and torch.equal(d2.rows[:10, 900], torch.tensor([float(x[2]) for x in host.buffer]))
    s, p, zz = dev.sample(5)
    assert s.shape == (5, 3, 15, 15) and zz.shape == (5, 1)
    assert tr.load_replay_buffer(str(tmp_path / "missing.pkl"), 10) is None


def test_architecture_inference():
    import alphazero_gomoku_b200.network as mynet
    for blocks, ch in ((3, 64), (6, 128), (2, 256)):
        net = mynet.AlphaZeroNet(n_res_blocks=blocks, channels=ch)
        assert mynet.infer_architecture(net.state_dict()) == (blocks, ch)
    assert sum(p.numel() for p in mynet.AlphaZeroNet(n_res_blocks=3, channels=64).parameters()) == 340010
    assert sum(p.numel() for p in mynet.AlphaZeroNet().parameters()) == 1892650       # SURVEY: 6x128


def test_mcts_symmetries_reference_order():
    """The host helper of the drop-in MCTS needs no engine: call it unbound."""
    from alphazero_gomoku_b200.mcts import MCTS
    z = load_golden("symmetry_sampling.npz")
    out = MCTS.symmetries(None, z["planes"], z["pi"])
    for i, (s, g) in enumerate(out):
        assert np.array_equal(s, z["sym_planes"][i]) and np.array_equal(g, z["sym_pi"][i])


def test_public_surface_matches_reference_fixture():
    """Every public member of the reference classes behind the drop-in boundary exists here with the same
    parameter names, order and defaults (tests/golden/api_surface.json, written by oracle/make_golden.py
    from the reference itself).  Extra engine-specific keyword arguments may follow the reference's."""
    import inspect
    import json
    import os
    from conftest import ROOT
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200 import games, network, players, train

    api = json.load(open(os.path.join(ROOT, "tests", "golden", "api_surface.json")))
    ours = {"MCTS": m.MCTS, "PyTorchModel": network.PyTorchModel, "AlphaZeroNet": network.AlphaZeroNet,
            "Gomoku": games.Gomoku, "Pente": games.Pente, "Player": players.Player}

    def check_params(where, fn, ref_params):
        mine = list(inspect.signature(fn).parameters.items())
        assert len(mine) >= len(ref_params), where
        for (name, prm), (rname, rdefault, has_default) in zip(mine, ref_params):
            assert name == rname, (where, name, rname)
            if has_default:
                d = prm.default
                assert d is not inspect._empty, (where, name)
                got = d.__name__ if inspect.isclass(d) else repr(d)
                assert got == rdefault, (where, name, got, rdefault)

    for cls_name, cls in ours.items():
        for member, ref_params in api[cls_name].items():
            assert hasattr(cls, member), f"{cls_name}.{member} missing"
            if isinstance(ref_params, list):
                check_params(f"{cls_name}.{member}", getattr(cls, member), ref_params)
    for fn_name, ref_params in api["train"].items():
        assert hasattr(train, fn_name), f"train.{fn_name} missing"
        check_params(f"train.{fn_name}", getattr(train, fn_name), ref_params)
    for cls_name, members in api["train_classes"].items():
        for member, ref_params in members.items():
            assert hasattr(getattr(train, cls_name), member), f"train.{cls_name}.{member} missing"
            if isinstance(ref_params, list):
                check_params(f"train.{cls_name}.{member}", getattr(getattr(train, cls_name), member), ref_params)


def test_shard_bounds_partition():
    from alphazero_gomoku_b200.train import shard_bounds
    for n in (0, 1, 5, 12, 13, 40):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_reference_written_files_load():
    """Files written by the REFERENCE's own save functions (oracle/make_golden.py build_train): the replay
    pickle (train.py:302-320) through load_replay_buffer, the checkpoint dictionary (network.py:240-248)
    into this package's parameter container, with the architecture read off the state_dict."""
    import os
    from conftest import GOLDEN
    from alphazero_gomoku_b200 import train as tr
    import alphazero_gomoku_b200.network as mynet
    from oracle import net as onet
    rows = load_golden("ref_replay_rows.npz")
    buf = tr.load_replay_buffer(os.path.join(GOLDEN, "ref_replay_buffer.pkl"), 50)
    assert buf is not None and buf.capacity == 50 and len(buf) == 40
    planes = np.unpackbits(rows["planes_bits"], axis=1)[:, :675].astype(np.float32).reshape(-1, 3, 15, 15)
    for i, (s, p, zz) in enumerate(buf.buffer):
        assert s.dtype == np.float32 and np.array_equal(s, planes[i]) and np.array_equal(p, rows["pi"][i]) and zz == rows["z"][i]
    dev = tr.DeviceReplayBuffer.from_host(buf, "cpu")
    assert torch.equal(dev.rows[:40, :675], torch.from_numpy(planes.reshape(40, -1)))
    assert torch.equal(dev.rows[:40, 675:900], torch.from_numpy(rows["pi"])) and torch.equal(dev.rows[:40, 900], torch.from_numpy(rows["z"]))
    smaller = tr.load_replay_buffer(os.path.join(GOLDEN, "ref_replay_buffer.pkl"), 16)       # deque(maxlen) keeps the newest
    assert len(smaller) == 16 and np.array_equal(smaller.buffer[0][1], rows["pi"][24])
    state = torch.load(os.path.join(GOLDEN, "ref_checkpoint_3x64_seed0.pt"), map_location="cpu")
    assert set(state) == {"net", "opt", "board_size", "action_size"} and state["board_size"] == 15 and state["action_size"] == 225
    assert mynet.infer_architecture(state["net"]) == (3, 64)
    net = mynet.AlphaZeroNet(n_res_blocks=3, channels=64)
    net.load_state_dict(state["net"])                       # strict: same keys and shapes
    z = load_golden("net_outputs.npz")
    logits, _ = onet.forward(net.state_dict(), torch.from_numpy(z["X"]))
    assert np.allclose(logits.numpy(), z["3x64/logits"], rtol=1e-4, atol=1e-4)
    assert mynet.infer_architecture(torch.load(os.path.join(GOLDEN, "ref_train_2x64_after3.pt"), map_location="cpu")["net"]) == (2, 64)
