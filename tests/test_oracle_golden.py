"""CPU: the oracle restatement against the fixtures generated from the reference
(oracle/make_golden.py).  No GPU, no /root/reference needed."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden
from oracle import fakes, rules, selfplay
from oracle.search import Search, dihedral8


def traces():
    z = load_golden("rules_traces.npz")
    for name, rule in zip(z["names"], z["rules"]):
        yield str(name), int(rule), {k: z[f"{name}/{k}"] for k in ("moves", "boards", "players", "caps", "winners", "overs", "oks", "lasts")}


@pytest.mark.parametrize("name,rule,t", list(traces()), ids=[n for n, _, _ in traces()])
def test_rules_trace(name, rule, t):
    pos = rules.Position(rule)
    for i, (r, c) in enumerate(t["moves"]):
        ok = rules.play_rc(pos, int(r), int(c))
        assert ok == bool(t["oks"][i]), (name, i)
        assert np.array_equal(pos.cells, t["boards"][i]), (name, i)
        assert pos.player == t["players"][i] and pos.caps == t["caps"][i].tolist() and pos.last == t["lasts"][i]
        assert rules.winner(pos) == t["winners"][i] and rules.game_over(pos) == bool(t["overs"][i])
        assert np.array_equal(rules.legal_mask(pos), (t["boards"][i] == 0).astype(np.float32))


def test_survey_known_facts():
    z = load_golden("rules_traces.npz")
    assert z["pente_capture_kat/caps"][-1].tolist() == [1, 0]
    assert z["pente_capture_kat/boards"][-1].reshape(15, 15)[7, 6:12].tolist() == [0, 1, 0, 0, 1, 0]
    assert z["pente_capture_win/winners"][-1] == 1 and z["pente_capture_win/caps"][-1][0] == 5
    assert z["gomoku_full_draw/winners"][-1] == 0 and z["gomoku_full_draw/overs"][-1]
    assert z["gomoku_overline/winners"][-1] == 1


SMALL = ["kat_uniform_100", "kat_uniform_400", "kat_uniform_400_pente", "g_hashed_1", "g_hashed_31", "g_hashed_32",
         "g_hashed_33", "g_hashed_100_q8", "g_fixedspike_200", "p_hashed_100_q8", "g_endgame_400", "p_endgame_300"]


@pytest.mark.parametrize("name", SMALL)
def test_search_visits(name):
    z = load_golden("search_visits.npz")
    rule, n_sims, q, _ = (int(x) for x in z[f"{name}/cfg"])
    model = fakes.BY_NAME[str(z[f"{name}/model"][0])]()
    s = Search(rule, n_sims, model, cpuct=float(z[f"{name}/cpuct"][0]), queue_len=q, noise=False)
    pos = rules.Position(rule)
    for a in z[f"{name}/opening"]:
        assert rules.play(pos, int(a))
    n_runs = min(len(z[f"{name}/moves"]), 4)
    for i in range(n_runs):
        pi = s.run(pos, pos.plies)
        assert np.array_equal(s.Nv[pos.key()].astype(np.int32), z[f"{name}/N"][i]), (name, i)
        assert np.array_equal(pi, z[f"{name}/pi"][i])
        assert (model.rows, model.calls) == tuple(z[f"{name}/evals"][i])
        a = int(np.argmax(pi))
        assert a == z[f"{name}/moves"][i]
        rules.play(pos, a)


def test_survey_visit_hashes():
    z = load_golden("search_visits.npz")
    for n, sha in {100: "7bbbd774cfe75043", 400: "7c21360da2a20891", 800: "06859f65f08731e8"}.items():
        assert hashlib.sha256(z[f"kat_uniform_{n}/N"][0].astype(np.int32).tobytes()).hexdigest()[:16] == sha


def test_noise_root_f64():
    z = load_golden("search_noise.npz")
    for tag, rule in (("g", rules.GOMOKU), ("p", rules.PENTE)):
        draws = list(z[f"{tag}/draws"])
        s = Search(rule, 300, fakes.Hashed(), cpuct=1.0, queue_len=32, alpha=0.05, eps=0.25, noise_plies=10, noise=True,
                   noise_fn=lambda n: draws.pop(0))
        pos = rules.Position(rule)
        for i in range(2):
            pi = s.run(pos, pos.plies)
            assert np.array_equal(pi, z[f"{tag}/pi"][i])
            rules.play(pos, int(np.argmax(pi)))


def test_symmetries_and_temperature():
    z = load_golden("symmetry_sampling.npz")
    out = dihedral8(z["planes"], z["pi"])
    for i, (s, g) in enumerate(out):
        assert np.array_equal(s, z["sym_planes"][i]) and np.array_equal(g, z["sym_pi"][i])
    for t, want in zip(z["temps"], z["tempered"]):
        assert np.array_equal(np.asarray(selfplay.temper(z["pi"], float(t)), dtype=np.float64), want)


def test_net_forward():
    import torch
    from oracle import net as onet
    import alphazero_gomoku_b200.network as mynet
    z = load_golden("net_outputs.npz")
    torch.set_num_threads(1)
    torch.manual_seed(0)
    net = mynet.AlphaZeroNet(n_res_blocks=3, channels=64)
    sd = net.state_dict()
    names = [str(n) for n in z["3x64/param_names"]]
    assert sorted(k for k, v in sd.items() if v.dtype.is_floating_point) == names
    sums = np.array([float(sd[k].double().sum()) for k in names])
    assert np.array_equal(sums, z["3x64/param_sums"]), "same seed must give the reference's initial weights"
    logits, value = onet.forward(sd, torch.from_numpy(z["X"]))
    assert np.allclose(logits.numpy(), z["3x64/logits"], rtol=1e-4, atol=1e-4)
    p, v = onet.CpuModel(sd).predict(z["X"])
    assert np.allclose(p, z["3x64/probs"], atol=1e-5) and np.allclose(v, z["3x64/values"], atol=1e-5)


def unpack_planes(bits):
    return np.unpackbits(bits, axis=1)[:, :675].astype(np.float32).reshape(-1, 3, 15, 15)


@pytest.mark.parametrize("name", ["g_t0", "p_t0", "g_cut", "g_temp_noise"])
def test_whole_selfplay_game(name):
    """oracle.selfplay.play_one == the reference's play_game_and_collect (train.py:360-412), row for row."""
    z = load_golden("selfplay_games.npz")
    rule, n_sims, max_moves, thr, syms, noise, winner = (int(v) for v in z[f"{name}/cfg"])
    temp_fn = (lambda mn: 0.0) if thr < 0 else (lambda mn: max(0.0, 1.0 - mn / thr))
    s = Search(rule, n_sims, fakes.BY_NAME[str(z[f"{name}/model"][0])](), cpuct=1.0, queue_len=32, alpha=0.3, eps=0.25,
               noise_plies=6, noise=bool(noise))
    np.random.seed(int(z["seed"][0]))
    rows, won = selfplay.play_one(s, rules.Position(rule), temp_fn, max_plies=max_moves, expand=bool(syms))
    assert won == winner and len(rows) == len(z[f"{name}/z"])
    assert np.array_equal(np.stack([r[0] for r in rows]), unpack_planes(z[f"{name}/planes_bits"]))
    assert np.array_equal(np.stack([r[1] for r in rows]), z[f"{name}/pi"])
    assert np.array_equal(np.array([r[2] for r in rows], dtype=np.float32), z[f"{name}/z"])


def test_train_step_oracle_against_reference_steps():
    """oracle.train.train_step == three reference train_batch steps (network.py:199-235): losses, every
    updated tensor, BatchNorm running statistics and the Adam moments the reference saved afterwards."""
    import os
    import torch
    from conftest import GOLDEN
    from oracle import train as otrain
    import alphazero_gomoku_b200.network as mynet
    z = load_golden("train_steps.npz")
    X, P, Z = unpack_planes(z["batch/planes_bits"]), z["batch/pi"], z["batch/z"]
    torch.set_num_threads(4)
    torch.manual_seed(0)
    sd = {k: v.clone() for k, v in mynet.AlphaZeroNet(n_res_blocks=2, channels=64).state_dict().items()}
    opt = otrain.Adam(sd)
    got = [otrain.train_step(sd, opt, X, P, Z) for _ in range(3)]
    want = z["2x64/losses"]
    for g, w in zip(got, want):
        assert np.allclose([g["policy_loss"], g["value_loss"], g["total_loss"]], w, rtol=2e-4, atol=2e-5), (g, w)
    ref = torch.load(os.path.join(GOLDEN, "ref_train_2x64_after3.pt"), map_location="cpu")
    assert set(ref) == {"net", "opt", "board_size", "action_size"}
    for k, v in ref["net"].items():
        if v.dtype.is_floating_point:
            # Adam normalises the step: a gradient element near zero can land on either side, one lr apart
            assert torch.allclose(sd[k], v, atol=2.5e-4), (k, float((sd[k] - v).abs().max()))
            assert float((sd[k] - v).abs().mean()) < 2e-5, k
        else:
            assert int(sd[k]) == int(v) == 3, k
    names = otrain.param_names(sd)
    for i, k in enumerate(names):
        st = ref["opt"]["state"][i]
        assert int(st["step"]) == 3
        # fp32 summation order of the convolutions differs with the thread count: compare at the tensor's own scale
        assert torch.allclose(opt.m[k], st["exp_avg"], rtol=1e-3, atol=1e-3 * float(st["exp_avg"].abs().max())), k
        assert torch.allclose(opt.v[k], st["exp_avg_sq"], rtol=1e-3, atol=1e-3 * float(st["exp_avg_sq"].abs().max())), k


def test_byte_compiled_reference_agrees_with_oracle():
    """oracle/_ref (the unmodified reference, byte-compiled by oracle/build_ref.py; present wherever the build ran) is what
    bench.py's CPU arm executes: run it side by side with the oracle on a small search with tree reuse."""
    from oracle import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built here")
    assert build_ref.activate()
    from games.gomoku import Gomoku                      # the reference's own modules, from oracle/_ref
    from mcts.new_mcts_alpha import MCTS
    ref = MCTS(Gomoku, 70, fakes.Hashed(), cpuct=1.0, batch_size=32, add_dirichlet_noise=False)
    orc = Search(rules.GOMOKU, 70, fakes.Hashed(), cpuct=1.0, queue_len=32, noise=False)
    g, pos = Gomoku(15), rules.Position(rules.GOMOKU)
    for _ in range(3):
        pi_ref, pi_orc = ref.run(g, len(g.move_history)), orc.run(pos, pos.plies)
        assert np.array_equal(pi_ref, pi_orc)
        a = int(np.argmax(pi_ref))
        g.do_move(divmod(a, 15))
        rules.play(pos, a)
