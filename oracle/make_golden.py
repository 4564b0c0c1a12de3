"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle to it.

Run in the build container only (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

For every fixture the reference's own classes (games/gomoku.py, games/pente.py,
mcts/new_mcts_alpha.py, network.py, train.py helpers) produce the expected
values; the oracle restatement is run on the same inputs and must agree bit for
bit (asserted here) before anything is written.  The GPU box has no
/root/reference: tests there compare against these committed files and against
the oracle.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

REF = os.environ.get("AZG_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

from games.gomoku import Gomoku          # noqa: E402  (reference)
from games.pente import Pente            # noqa: E402  (reference)
from mcts.new_mcts_alpha import MCTS     # noqa: E402  (reference)

from oracle import fakes, rules          # noqa: E402
from oracle.search import Search, dihedral8   # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
REF_GAME = {rules.GOMOKU: Gomoku, rules.PENTE: Pente}


def ref_caps(g):
    return [g.captures[1], g.captures[2]] if hasattr(g, "captures") else [0, 0]


def check_same(g, pos, where):
    assert np.array_equal(g.board.reshape(-1), pos.cells), where
    assert g.current_player == pos.player, where
    assert ref_caps(g) == pos.caps, where
    assert g.check_winner() == rules.winner(pos), where
    assert bool(g.is_game_over()) == rules.game_over(pos), where
    assert np.array_equal(g.get_valid_moves(), rules.legal_mask(pos)), where
    assert np.array_equal(g.get_encoded_state(), rules.encode(pos)), where


# --------------------------------------------------------------------------- rules
def scripted_traces():
    """Hand-built move lists (r, c) that hit the edge cases SURVEY section 4 lists."""
    T = []
    # horizontal five on the top edge, player 1
    T.append(("gomoku_edge_row", rules.GOMOKU, [(0, 0), (5, 5), (0, 1), (5, 6), (0, 2), (5, 7), (0, 3), (9, 9), (0, 4)]))
    # vertical five in the last column, player 2
    T.append(("gomoku_edge_col", rules.GOMOKU, [(7, 7), (14, 14), (7, 8), (13, 14), (1, 1), (12, 14), (2, 2), (11, 14), (9, 3), (10, 14)]))
    # overline (six) completed in the middle wins
    T.append(("gomoku_overline", rules.GOMOKU, [(3, 3), (9, 0), (3, 4), (9, 1), (3, 5), (9, 2), (3, 7), (12, 0), (3, 8), (12, 1), (3, 6)]))
    # both diagonals
    T.append(("gomoku_diag", rules.GOMOKU, [(4, 4), (0, 9), (5, 5), (1, 9), (6, 6), (2, 9), (7, 7), (5, 9), (8, 8)]))
    T.append(("gomoku_antidiag", rules.GOMOKU, [(0, 0), (4, 10), (1, 1), (5, 9), (2, 2), (6, 8), (9, 9), (7, 7), (10, 10), (8, 6)]))
    # illegal attempts: occupied, off board (do_move must return False and change nothing)
    T.append(("gomoku_illegal", rules.GOMOKU, [(7, 7), (7, 7), (-1, 0), (15, 0), (0, 15), (0, -1), (7, 8)]))
    # SURVEY 8c Pente fact: capture along (0,-1) from the new stone
    T.append(("pente_capture_kat", rules.PENTE, [(7, 7), (7, 8), (0, 0), (7, 9), (7, 10)]))
    # double capture with one stone (two rays)
    T.append(("pente_double", rules.PENTE, [(7, 4), (7, 5), (4, 7), (7, 6), (0, 0), (5, 7), (0, 1), (6, 7), (7, 7)]))
    # capture on the edge rays / no capture when the far cell is off board
    T.append(("pente_edge", rules.PENTE, [(0, 3), (0, 1), (9, 9), (0, 2), (5, 5), (14, 14), (0, 0)]))
    # five pairs captured by player 1 -> capture win
    seq = []
    for i in range(5):
        r = 2 * i + 1
        seq += [(r, 3), (r, 1), (r + 1, 9 + (i % 2)), (r, 2), (r, 0)]
        if i < 4:
            seq += [(14, i)]
    T.append(("pente_capture_win", rules.PENTE, seq))
    T.append(("pente_illegal", rules.PENTE, [(7, 7), (7, 7), (20, 3), (3, -2), (7, 8)]))
    return T


def run_trace(rule, moves):
    """Play (r,c) moves on reference and oracle; return per-ply arrays."""
    g = REF_GAME[rule](15)
    pos = rules.Position(rule)
    boards, players, caps, wins, overs, oks, lasts = [], [], [], [], [], [], []
    for i, (r, c) in enumerate(moves):
        ok_ref = g.do_move((r, c))
        ok_or = rules.play_rc(pos, r, c)
        assert ok_ref == ok_or, (rule, i, r, c)
        check_same(g, pos, (rule, i))
        assert (g.last_move is None and pos.last < 0) or (g.last_move[0] * 15 + g.last_move[1] == pos.last)
        boards.append(g.board.reshape(-1).astype(np.int8).copy())
        players.append(g.current_player)
        caps.append(ref_caps(g))
        wins.append(g.check_winner())
        overs.append(bool(g.is_game_over()))
        oks.append(bool(ok_ref))
        lasts.append(pos.last)
    return dict(moves=np.array(moves, dtype=np.int16).reshape(-1, 2), boards=np.array(boards, dtype=np.int8),
                players=np.array(players, dtype=np.int8), caps=np.array(caps, dtype=np.int16),
                winners=np.array(wins, dtype=np.int8), overs=np.array(overs, dtype=np.bool_),
                oks=np.array(oks, dtype=np.bool_), lasts=np.array(lasts, dtype=np.int16))


def random_trace(rule, rng, p_illegal=0.03, dense=False):
    """Random playout to the end.  ``dense`` biases moves towards existing stones
    so that captures and lines actually occur."""
    g = REF_GAME[rule](15)
    moves = []
    while not g.is_game_over():
        if rng.random() < p_illegal:
            moves.append((int(rng.integers(-2, 17)), int(rng.integers(-2, 17))))
            if g.clone().do_move(moves[-1]):
                g.do_move(moves[-1])
            continue
        empties = np.flatnonzero(g.board.reshape(-1) == 0)
        if dense and len(g.move_history) > 0 and rng.random() < 0.85:
            pr, pc = g.move_history[int(rng.integers(0, len(g.move_history)))]
            cand = [(pr + dr) * 15 + (pc + dc) for dr in range(-3, 4) for dc in range(-3, 4)
                    if 0 <= pr + dr < 15 and 0 <= pc + dc < 15 and g.board[pr + dr, pc + dc] == 0]
            a = int(cand[int(rng.integers(0, len(cand)))]) if cand else int(empties[int(rng.integers(0, len(empties)))])
        else:
            a = int(empties[int(rng.integers(0, len(empties)))])
        moves.append(divmod(a, 15))
        g.do_move(moves[-1])
    return moves


def build_rules():
    out = {}
    names = []
    for name, rule, moves in scripted_traces():
        t = run_trace(rule, moves)
        names.append((name, rule))
        for k, v in t.items():
            out[f"{name}/{k}"] = v
    rng = np.random.default_rng(20261018)
    for rule, tag in ((rules.GOMOKU, "gomoku"), (rules.PENTE, "pente")):
        for i in range(12):
            moves = random_trace(rule, rng, dense=(i % 2 == 1))
            t = run_trace(rule, moves)
            name = f"{tag}_random{i}"
            names.append((name, rule))
            for k, v in t.items():
                out[f"{name}/{k}"] = v
    # board-full draw: fill the board in an order that never makes five in a row
    g = Gomoku(15)
    moves = draw_fill_order()
    t = run_trace(rules.GOMOKU, moves)
    assert t["winners"][-1] == 0 and t["overs"][-1] and not t["overs"][-2], "draw fill must end in a draw"
    names.append(("gomoku_full_draw", rules.GOMOKU))
    for k, v in t.items():
        out[f"gomoku_full_draw/{k}"] = v
    out["names"] = np.array([n for n, _ in names])
    out["rules"] = np.array([r for _, r in names], dtype=np.int8)
    # the SURVEY 8c facts, asserted on the reference itself
    t = {k.split("/", 1)[1]: v for k, v in out.items() if k.startswith("pente_capture_kat/")}
    assert t["caps"][-1].tolist() == [1, 0] and t["boards"][-1].reshape(15, 15)[7, 6:12].tolist() == [0, 1, 0, 0, 1, 0]
    np.savez_compressed(os.path.join(OUT, "rules_traces.npz"), **out)
    plies = sum(len(v) for k, v in out.items() if k.endswith("/players"))
    print(f"rules_traces.npz: {len(names)} traces, {plies} plies, caps max {max(int(v.max()) for k, v in out.items() if k.endswith('/caps'))}")


def draw_fill_order():
    """A full 225-move game with no five in a row: colour(r,c) = ((c + 2r) mod 4) // 2
    gives runs of at most two in every direction, 113 cells for player 1 and 112 for
    player 2; the moves are interleaved to respect turn order."""
    want = np.array([[((c + 2 * r) % 4) // 2 + 1 for c in range(15)] for r in range(15)], dtype=np.int8)
    ones = [tuple(x) for x in np.argwhere(want == 1)]
    twos = [tuple(x) for x in np.argwhere(want == 2)]
    assert len(ones) == 113 and len(twos) == 112
    order = []
    for i in range(112):
        order += [ones[i], twos[i]]
    order.append(ones[112])
    return [(int(r), int(c)) for r, c in order]


# --------------------------------------------------------------------------- search
class RefGameFactory:
    """Build a reference game object from an oracle Position (for MCTS.run)."""

    @staticmethod
    def make(pos):
        g = REF_GAME[pos.rule](15)
        g.board = pos.cells.reshape(15, 15).copy()
        g.current_player = pos.player
        g.last_move = None if pos.last < 0 else divmod(pos.last, 15)
        g.move_history = [(0, 0)] * pos.plies          # only its length is read (train.py:371)
        if pos.rule == rules.PENTE:
            g.captures = {1: pos.caps[0], 2: pos.caps[1]}
        return g


def search_case(rule, model_name, n_sims, queue_len, n_moves, cpuct=1.0, opening=()):
    """Play ``n_moves`` argmax moves with tree reuse on reference and oracle; record
    N[root] (int32), pi (f32), eval counts and the chosen move for every run."""
    ref_model = fakes.BY_NAME[model_name]()
    orc_model = fakes.BY_NAME[model_name]()
    ref = MCTS(REF_GAME[rule], n_sims, ref_model, cpuct=cpuct, batch_size=queue_len, add_dirichlet_noise=False)
    orc = Search(rule, n_sims, orc_model, cpuct=cpuct, queue_len=queue_len, noise=False)
    g = REF_GAME[rule](15)
    pos = rules.Position(rule)
    for mv in opening:
        assert g.do_move(divmod(mv, 15)) and rules.play(pos, mv)
    Ns, pis, evals, moves, nodes = [], [], [], [], []
    for _ in range(n_moves):
        if g.is_game_over():
            break
        pi_ref = ref.run(g, len(g.move_history))
        pi_orc = orc.run(pos, pos.plies)
        key = ref._state_key(g)
        assert key == pos.key()
        assert pi_ref.dtype == np.float32 and np.array_equal(pi_ref, pi_orc), (rule, model_name, n_sims, queue_len)
        assert np.array_equal(ref.N[key], orc.Nv[key])
        assert len(ref.P) == len(orc.P) and ref_model.rows == orc_model.rows and ref_model.calls == orc_model.calls
        Ns.append(ref.N[key].astype(np.int32))
        pis.append(pi_ref.copy())
        evals.append((ref_model.rows, ref_model.calls))
        nodes.append(len(ref.P))
        a = int(np.argmax(pi_ref))
        moves.append(a)
        g.do_move(divmod(a, 15))
        rules.play(pos, a)
    return dict(N=np.array(Ns, dtype=np.int32), pi=np.array(pis, dtype=np.float32),
                evals=np.array(evals, dtype=np.int64), moves=np.array(moves, dtype=np.int16),
                nodes=np.array(nodes, dtype=np.int64), opening=np.array(opening, dtype=np.int16))


SEARCH_CASES = [
    # (name, rule, model, n_sims, queue_len, n_moves, cpuct, opening)
    ("kat_uniform_100", rules.GOMOKU, "uniform", 100, 32, 1, 1.0, ()),
    ("kat_uniform_400", rules.GOMOKU, "uniform", 400, 32, 1, 1.0, ()),
    ("kat_uniform_800", rules.GOMOKU, "uniform", 800, 32, 1, 1.0, ()),
    ("kat_uniform_400_pente", rules.PENTE, "uniform", 400, 32, 1, 1.0, ()),
    ("g_hashed_1", rules.GOMOKU, "hashed", 1, 32, 3, 1.0, ()),
    ("g_hashed_31", rules.GOMOKU, "hashed", 31, 32, 4, 1.0, ()),
    ("g_hashed_32", rules.GOMOKU, "hashed", 32, 32, 4, 1.0, ()),
    ("g_hashed_33", rules.GOMOKU, "hashed", 33, 32, 4, 1.0, ()),
    ("g_hashed_100_q8", rules.GOMOKU, "hashed", 100, 8, 6, 1.0, ()),
    ("g_hashed_60_q1", rules.GOMOKU, "hashed", 60, 1, 3, 1.0, ()),
    ("g_hashed_400", rules.GOMOKU, "hashed", 400, 32, 30, 1.0, ()),
    ("g_hashed_800", rules.GOMOKU, "hashed", 800, 32, 6, 1.0, ()),
    ("g_hashed_400_c25", rules.GOMOKU, "hashed", 400, 32, 8, 2.5, ()),
    ("g_spiky_400", rules.GOMOKU, "spiky", 400, 32, 40, 1.0, ()),
    ("g_fixedspike_200", rules.GOMOKU, "fixedspike", 200, 32, 12, 1.0, ()),
    ("p_hashed_400", rules.PENTE, "hashed", 400, 32, 30, 1.0, ()),
    ("p_spiky_300", rules.PENTE, "spiky", 300, 32, 40, 1.0, ()),
    ("p_hashed_100_q8", rules.PENTE, "hashed", 100, 8, 10, 1.0, ()),
    # terminal-heavy: an opening where both sides hold open fours / capture threats
    ("g_endgame_400", rules.GOMOKU, "hashed", 400, 32, 12, 1.0,
     (112, 97, 113, 98, 114, 99, 115, 100, 7 * 15 + 2, 6 * 15 + 2)),
    ("p_endgame_300", rules.PENTE, "hashed", 300, 32, 12, 1.0,
     (112, 113, 0, 114, 115, 128, 1, 143, 158, 127, 2, 142)),
]


def build_search():
    out = {"names": np.array([c[0] for c in SEARCH_CASES])}
    for name, rule, model, n_sims, q, n_moves, cpuct, opening in SEARCH_CASES:
        r = search_case(rule, model, n_sims, q, n_moves, cpuct, opening)
        for k, v in r.items():
            out[f"{name}/{k}"] = v
        out[f"{name}/cfg"] = np.array([rule, n_sims, q, n_moves], dtype=np.int64)
        out[f"{name}/cpuct"] = np.array([cpuct], dtype=np.float64)
        out[f"{name}/model"] = np.array([model])
        print(f"  {name}: {len(r['moves'])} runs, evals {r['evals'][-1].tolist()}, nodes {int(r['nodes'][-1])}, sumN0 {int(r['N'][0].sum())}")
    # SURVEY 8c known answers, re-derived from the reference here
    for n, (tot, nnz, mx, sha) in {100: (69, 69, 1, "7bbbd774cfe75043"), 400: (369, 225, 2, "7c21360da2a20891"),
                                   800: (769, 225, 4, "06859f65f08731e8")}.items():
        N0 = out[f"kat_uniform_{n}/N"][0]
        got = hashlib.sha256(N0.astype(np.int32).tobytes()).hexdigest()[:16]
        assert (int(N0.sum()), int((N0 > 0).sum()), int(N0.max()), got) == (tot, nnz, mx, sha), (n, got)
    np.savez_compressed(os.path.join(OUT, "search_visits.npz"), **out)
    print("search_visits.npz written")


# --------------------------------------------------------------------------- noise (f64 root)
def build_noise():
    """Root Dirichlet mixing with an injected noise vector: the reference draws from
    numpy's global generator (:172), so seed it, capture the draw, and feed the same
    vector to the oracle.  Pins the float64 PUCT path at a noised root."""
    out = {}
    for tag, rule in (("g", rules.GOMOKU), ("p", rules.PENTE)):
        ref_model, orc_model = fakes.Hashed(), fakes.Hashed()
        ref = MCTS(REF_GAME[rule], 300, ref_model, cpuct=1.0, batch_size=32, dirichlet_alpha=0.05, epsilon=0.25,
                   apply_dirichlet_n_first_moves=10, add_dirichlet_noise=True)
        draws = []
        rs = np.random.RandomState(99)

        def noise_fn(n, rs=rs, draws=draws):
            d = rs.dirichlet([0.05] * n)
            draws.append(d)
            return d
        orc = Search(rule, 300, orc_model, cpuct=1.0, queue_len=32, alpha=0.05, eps=0.25, noise_plies=10, noise=True,
                     noise_fn=noise_fn)
        np.random.seed(99)
        g = REF_GAME[rule](15)
        pos = rules.Position(rule)
        Ns, pis, moves = [], [], []
        for _ in range(4):
            state = np.random.get_state()
            pi_ref = ref.run(g, len(g.move_history))
            after = np.random.get_state()
            np.random.set_state(state)
            rs.set_state(state)
            pi_orc = orc.run(pos, pos.plies)
            np.random.set_state(after)
            assert np.array_equal(pi_ref, pi_orc)
            key = pos.key()
            Ns.append(ref.N[key].astype(np.int32))
            pis.append(pi_ref.astype(np.float32))
            a = int(np.argmax(pi_ref))
            moves.append(a)
            g.do_move(divmod(a, 15))
            rules.play(pos, a)
        out[f"{tag}/N"] = np.array(Ns)
        out[f"{tag}/pi"] = np.array(pis)
        out[f"{tag}/moves"] = np.array(moves, dtype=np.int16)
        out[f"{tag}/draws"] = np.array(draws, dtype=np.float64)
        print(f"  noise {tag}: {len(draws)} draws, sumN {[int(n.sum()) for n in Ns]}")
    np.savez_compressed(os.path.join(OUT, "search_noise.npz"), **out)


# --------------------------------------------------------------------------- symmetries / sampling
def build_misc():
    import train as ref_train        # reference train.py (imports torch)
    rng = np.random.default_rng(5)
    planes = (rng.random((3, 15, 15)) < 0.3).astype(np.float32)
    pi = rng.random(225).astype(np.float32)
    pi /= pi.sum()
    ref = MCTS(Gomoku, 1, None).symmetries(planes, pi)
    mine = dihedral8(planes, pi)
    for (a, b), (c, d) in zip(ref, mine):
        assert np.array_equal(a, c) and np.array_equal(b, d)
    from oracle import selfplay
    temps = [0.0, 1.0, 0.7, 0.1]
    tp = []
    for t in temps:
        r = ref_train.softmax_temperature(pi, t)
        m = selfplay.temper(pi, t)
        assert np.array_equal(r, m)
        tp.append(np.asarray(r, dtype=np.float64))
    assert ref_train.sample_action_from_pi(pi, 0) == selfplay.pick(pi, 0)
    np.savez_compressed(os.path.join(OUT, "symmetry_sampling.npz"), planes=planes, pi=pi,
                        sym_planes=np.array([np.ascontiguousarray(a) for a, _ in ref]),
                        sym_pi=np.array([b for _, b in ref]), temps=np.array(temps), tempered=np.array(tp))
    print("symmetry_sampling.npz written")


# --------------------------------------------------------------------------- network
def build_net():
    import torch
    import network as ref_net          # reference network.py
    from oracle import net as onet
    out = {}
    rng = np.random.default_rng(11)
    # 24 positions from random legal playouts of 0..120 plies (SURVEY 8d recipe)
    X = []
    for i in range(24):
        pos = rules.Position(rules.GOMOKU)
        for _ in range(int(rng.integers(0, 121))):
            e = np.flatnonzero(pos.cells == 0)
            rules.play(pos, int(e[int(rng.integers(0, len(e)))]))
        X.append(rules.encode(pos))
    X = np.stack(X).astype(np.float32)
    out["X"] = X
    torch.set_num_threads(1)
    for tag, blocks, ch in (("3x64", 3, 64), ("6x128", 6, 128)):
        torch.manual_seed(0)
        m = ref_net.PyTorchModel(board_size=15, n_res_blocks=blocks, channels=ch, device="cpu")
        sd = m.net.state_dict()
        probs, values = m.predict(X)
        with torch.no_grad():
            m.net.eval()
            logits, _ = m.net(torch.from_numpy(X))
        lo, va = onet.forward(sd, torch.from_numpy(X))
        assert torch.equal(lo, logits), "oracle forward differs from reference forward"
        cm = onet.CpuModel(sd)
        p2, v2 = cm.predict(X)
        assert np.array_equal(p2, probs) and np.array_equal(v2, values)
        out[f"{tag}/logits"] = logits.numpy()
        out[f"{tag}/probs"] = probs
        out[f"{tag}/values"] = values
        # per-tensor checksums so tests can prove "same seed -> same weights" without shipping them
        names = sorted(k for k, v in sd.items() if v.dtype.is_floating_point)
        out[f"{tag}/param_names"] = np.array(names)
        out[f"{tag}/param_sums"] = np.array([float(sd[k].double().sum()) for k in names])
        out[f"{tag}/param_abs"] = np.array([float(sd[k].double().abs().sum()) for k in names])
        out[f"{tag}/n_params"] = np.array([sum(p.numel() for p in m.net.parameters())])
        print(f"  net {tag}: params {int(out[f'{tag}/n_params'][0])}, logit std {float(logits.std()):.2f}")
    np.savez_compressed(os.path.join(OUT, "net_outputs.npz"), **out)


def build_api():
    """Public surface of the reference classes/functions behind the drop-in boundary: member names
    and, for callables, parameter names with their defaults -> tests/golden/api_surface.json."""
    import importlib
    import inspect
    import json

    def params(fn):
        out = []
        try:
            sig = inspect.signature(fn)
        except (TypeError, ValueError):
            return None
        for name, prm in sig.parameters.items():
            d = prm.default
            out.append([name, None if d is inspect._empty else repr(d) if not inspect.isclass(d) else d.__name__,
                        d is not inspect._empty])
        return out

    def surface(cls):
        members = {}
        for name, obj in inspect.getmembers(cls):
            if name.startswith("_") and name != "__init__":
                continue
            members[name] = params(obj) if callable(obj) else "attribute"
        return members

    train = importlib.import_module("train")
    api = {
        "MCTS": surface(MCTS),
        "PyTorchModel": surface(importlib.import_module("network").PyTorchModel),
        "AlphaZeroNet": surface(importlib.import_module("network").AlphaZeroNet),
        "Gomoku": surface(Gomoku),
        "Pente": surface(Pente),
        "Player": surface(importlib.import_module("players.player_alpha").Player),
        "train": {n: params(o) for n, o in inspect.getmembers(train)
                  if inspect.isfunction(o) and o.__module__ == "train" and not n.startswith("_")},
        "train_classes": {n: surface(o) for n, o in inspect.getmembers(train)
                          if inspect.isclass(o) and o.__module__ == "train"},
    }
    # torch.nn.Module's own members are not the reference's surface
    import torch
    for k in set(dir(torch.nn.Module)):
        api["AlphaZeroNet"].pop(k, None) if k not in ("forward", "__init__") else None
    with open(os.path.join(OUT, "api_surface.json"), "w") as f:
        json.dump(api, f, indent=1, sort_keys=True)
    print("api surface:", {k: len(v) for k, v in api.items()})


# --------------------------------------------------------------------------- whole self-play game (a17)
def _pack_rows(examples):
    """(planes, pi, z) tuples -> (planes as packed bits uint8[n, 85], pi f32[n,225], z f32[n])."""
    st = np.stack([e[0] for e in examples]).astype(np.float32)
    assert ((st == 0) | (st == 1)).all()
    bits = np.packbits(st.reshape(len(examples), -1).astype(np.uint8), axis=1)
    return bits, np.stack([e[1] for e in examples]).astype(np.float32), np.array([e[2] for e in examples], dtype=np.float32)


def build_game():
    """train.py:360-412 ``play_game_and_collect`` run by the reference with injected priors: whole games,
    every example row (planes, pi, z) in order.  Cases: T = 0 without noise (Gomoku, Pente by duck typing,
    a max_moves cut-off) and the reference's temperature schedule with root noise, numpy seeded."""
    import train as ref_train
    from oracle import selfplay
    out = {}
    cases = [("g_t0", rules.GOMOKU, "hashed", 64, 225, None, True), ("p_t0", rules.PENTE, "spiky", 48, 225, None, True),
             ("g_cut", rules.GOMOKU, "hashed", 40, 7, None, False), ("g_temp_noise", rules.GOMOKU, "hashed", 96, 225, 8, True)]
    for name, rule, model, n_sims, max_moves, thr, syms in cases:
        noise = thr is not None
        temp_fn = (lambda mn: 0.0) if thr is None else (lambda mn, thr=thr: max(0.0, 1.0 - mn / thr))
        kw = dict(cpuct=1.0, dirichlet_alpha=0.3, epsilon=0.25, apply_dirichlet_n_first_moves=6, add_dirichlet_noise=noise)
        ref = MCTS(REF_GAME[rule], n_sims, fakes.BY_NAME[model](), **kw)
        g = REF_GAME[rule](15)
        np.random.seed(4242)
        ex_ref, win_ref = ref_train.play_game_and_collect(ref, g, temp_fn, max_moves=max_moves, use_symmetries=syms)
        orc = Search(rule, n_sims, fakes.BY_NAME[model](), cpuct=1.0, queue_len=32, alpha=0.3, eps=0.25, noise_plies=6, noise=noise)
        np.random.seed(4242)
        ex_orc, win_orc = selfplay.play_one(orc, rules.Position(rule), temp_fn, max_plies=max_moves, expand=syms)
        assert win_ref == win_orc and len(ex_ref) == len(ex_orc)
        for (a, b, c), (d, e, f) in zip(ex_ref, ex_orc):
            assert a.dtype == np.float32 and b.dtype == np.float32
            assert np.array_equal(a, d) and np.array_equal(b, e) and c == f
        bits, pis, zs = _pack_rows(ex_ref)
        out[f"{name}/planes_bits"], out[f"{name}/pi"], out[f"{name}/z"] = bits, pis, zs
        out[f"{name}/moves"] = np.array([r * 15 + c for r, c in g.move_history], dtype=np.int16)
        out[f"{name}/cfg"] = np.array([rule, n_sims, max_moves, -1 if thr is None else thr, int(syms), int(noise), win_ref], dtype=np.int64)
        out[f"{name}/model"] = np.array([model])
        print(f"  game {name}: {len(g.move_history)} plies, winner {win_ref}, {len(ex_ref)} rows")
    out["names"] = np.array([c[0] for c in cases])
    out["seed"] = np.array([4242])
    np.savez_compressed(os.path.join(OUT, "selfplay_games.npz"), **out)


# --------------------------------------------------------------------------- evaluation arena (f2)
class _ArenaFake:
    """A fake evaluator with the one attribute evaluate_models reads from a model (train.py:430)."""

    def __init__(self, name):
        self.inner = fakes.BY_NAME[name]()
        self.board_size = 15

    def predict(self, X):
        return self.inner.predict(X)


def build_arena():
    """train.py:418-486 ``evaluate_models`` run by the reference between two injected-prior models:
    result tuple plus the transcript (every move of every game), captured from outside by giving the
    reference's module a Gomoku subclass that logs ``do_move``."""
    import random
    import train as ref_train
    games = []

    class Logged(Gomoku):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            self._top = True
            games.append([])
            self._log = games[-1]

        def clone(self):
            c = super().clone()
            c._top = False
            return c

        def do_move(self, mv):
            ok = super().do_move(mv)
            if getattr(self, "_top", False) and ok:
                self._log.append(mv[0] * 15 + mv[1])
            return ok

    out = {}
    saved = ref_train.GameClass
    ref_train.GameClass = Logged
    try:
        for name, a, b, n_games, n_sims, cpuct in (("hashed_vs_spiky", "hashed", "spiky", 6, 48, 1.0),
                                                    ("spiky_vs_hashed", "spiky", "hashed", 5, 64, 1.5)):
            games.clear()
            random.seed(77)
            res = ref_train.evaluate_models(_ArenaFake(a), _ArenaFake(b), "gomoku", n_games=n_games, n_simulations=n_sims, cpuct=cpuct)
            # evaluate_models builds its MCTS with game_class=GameClass: instances made inside the search count too
            tops = [g for g in games if len(g) > 0][:n_games]
            assert len(tops) == n_games
            L = max(len(g) for g in tops)
            moves = np.full((n_games, L), -1, dtype=np.int16)
            for i, g in enumerate(tops):
                moves[i, :len(g)] = g
            out[f"{name}/moves"] = moves
            out[f"{name}/result"] = np.array([res[0], res[2]], dtype=np.int64)
            out[f"{name}/win_rate"] = np.array([res[1]])
            out[f"{name}/cfg"] = np.array([n_games, n_sims], dtype=np.int64)
            out[f"{name}/cpuct"] = np.array([cpuct])
            out[f"{name}/models"] = np.array([a, b])
            print(f"  arena {name}: new wins {res[0]}, draws {res[2]}, game lengths {[len(g) for g in tops]}")
    finally:
        ref_train.GameClass = saved
    out["names"] = np.array(["hashed_vs_spiky", "spiky_vs_hashed"])
    out["seed"] = np.array([77])
    np.savez_compressed(os.path.join(OUT, "arena_transcripts.npz"), **out)


# --------------------------------------------------------------------------- train_batch (f1) + formats (f3)
def _train_data(n, seed):
    """n training rows shaped like self-play output: positions from random legal playouts, a smooth
    target distribution over the empties, z in {-1, 0, 1}."""
    rng = np.random.default_rng(seed)
    X, P, Z = [], [], []
    for _ in range(n):
        pos = rules.Position(rules.GOMOKU)
        for _ in range(int(rng.integers(0, 60))):
            e = np.flatnonzero(pos.cells == 0)
            rules.play(pos, int(e[int(rng.integers(0, len(e)))]))
        X.append(rules.encode(pos))
        w = rng.random(225) ** 4 * (pos.cells == 0)
        P.append((w / w.sum()).astype(np.float32))
        Z.append(float(rng.integers(-1, 2)))
    return np.stack(X).astype(np.float32), np.stack(P), np.array(Z, dtype=np.float32).reshape(-1, 1)


def build_train():
    """network.py:199-235 ``train_batch`` run by the reference (fp32, CPU): three Adam steps on one seeded
    batch for a 2x64 and a 6x128 net (losses; the 2x64 model is then written with the reference's own
    ``save`` so weights, BatchNorm statistics and Adam moments after the steps are the fixture), and a
    100-step loss curve over a fixed 256-row data set.  Also the format fixtures: a seed-0 3x64 checkpoint
    (``PyTorchModel.save``, network.py:240-248) and a replay-buffer pickle (train.py:302-320)."""
    import torch
    import network as ref_net
    import train as ref_train
    torch.set_num_threads(4)
    out = {}
    X, P, Z = _train_data(32, 101)
    out["batch/planes_bits"] = np.packbits(X.reshape(32, -1).astype(np.uint8), axis=1)
    out["batch/pi"], out["batch/z"] = P, Z
    for tag, blocks, ch in (("2x64", 2, 64), ("6x128", 6, 128)):
        torch.manual_seed(0)
        m = ref_net.PyTorchModel(board_size=15, n_res_blocks=blocks, channels=ch, device="cpu")
        before = {k: v.clone() for k, v in m.net.state_dict().items()}
        losses = [m.train_batch(X, P, Z, epochs=1) for _ in range(3)]
        out[f"{tag}/losses"] = np.array([[l["policy_loss"], l["value_loss"], l["total_loss"]] for l in losses])
        after = m.net.state_dict()
        names = sorted(k for k, v in after.items() if v.dtype.is_floating_point)
        out[f"{tag}/names"] = np.array(names)
        out[f"{tag}/delta_l2"] = np.array([float((after[k].double() - before[k].double()).norm()) for k in names])
        out[f"{tag}/after_sum"] = np.array([float(after[k].double().sum()) for k in names])
        probs, values = m.predict(X[:8])          # eval-mode forward with the updated running statistics
        out[f"{tag}/probs_after"], out[f"{tag}/values_after"] = probs, values
        if tag == "2x64":
            m.save(os.path.join(OUT, "ref_train_2x64_after3.pt"))
        print(f"  train {tag}: losses {[round(l['total_loss'], 4) for l in losses]}")
    # 100-step curve, 2x64, fixed data set of 256 rows, batches of 32 by a seeded permutation stream
    DX, DP, DZ = _train_data(256, 202)
    out["curve/planes_bits"] = np.packbits(DX.reshape(256, -1).astype(np.uint8), axis=1)
    out["curve/pi"], out["curve/z"] = DP, DZ
    rng = np.random.default_rng(303)
    idx = np.stack([rng.choice(256, 32, replace=False) for _ in range(100)])
    out["curve/idx"] = idx.astype(np.int16)
    torch.manual_seed(0)
    m = ref_net.PyTorchModel(board_size=15, n_res_blocks=2, channels=64, device="cpu")
    curve = []
    for i in range(100):
        l = m.train_batch(DX[idx[i]], DP[idx[i]], DZ[idx[i]], epochs=1)
        curve.append([l["policy_loss"], l["value_loss"], l["total_loss"]])
    out["curve/losses"] = np.array(curve)
    print(f"  curve: total loss {curve[0][2]:.4f} -> {curve[-1][2]:.4f}")
    np.savez_compressed(os.path.join(OUT, "train_steps.npz"), **out)
    # format fixtures written by the reference's own functions
    torch.manual_seed(0)
    ref_net.PyTorchModel(board_size=15).save(os.path.join(OUT, "ref_checkpoint_3x64_seed0.pt"))
    buf = ref_train.ReplayBuffer(capacity=50)
    ref = MCTS(Gomoku, 24, fakes.Hashed(), add_dirichlet_noise=False)
    ex, _ = ref_train.play_game_and_collect(ref, Gomoku(15), lambda mn: 0.0, max_moves=5, use_symmetries=True)
    buf.add(ex)
    assert len(buf) == 40
    assert ref_train.save_replay_buffer(buf, os.path.join(OUT, "ref_replay_buffer.pkl"))
    bits, pis, zs = _pack_rows(ex)
    np.savez_compressed(os.path.join(OUT, "ref_replay_rows.npz"), planes_bits=bits, pi=pis, z=zs)


def _capture_moves(board, player):
    """Empty cells from which `player` flanks exactly two opposing stones in some direction (pente.py:114-152)."""
    opp, out = 3 - player, []
    for r in range(15):
        for c in range(15):
            if board[r, c] != 0:
                continue
            for dr in (-1, 0, 1):
                for dc in (-1, 0, 1):
                    if dr == 0 and dc == 0:
                        continue
                    r3, c3 = r + 3 * dr, c + 3 * dc
                    if 0 <= r3 < 15 and 0 <= c3 < 15 and board[r + dr, c + dc] == opp and board[r + 2 * dr, c + 2 * dc] == opp \
                            and board[r3, c3] == player:
                        out.append((r, c))
    return out


def build_undo():
    """do_move / undo_move sequences on the reference's game objects (gomoku.py:57-98, pente.py:57-109): after every
    operation the board, the side to move, the last move and (Pente) the capture counts.  undo_move is off the search
    path (the search clones), but it is part of the drop-in classes - including pente.py:100-103, which restores captured
    stones in the capturer's colour."""
    rng = np.random.default_rng(20260202)
    out = {}
    for rule, cls, tag in ((rules.GOMOKU, Gomoku, "g"), (rules.PENTE, Pente, "p")):
        for k in range(3):
            g = cls(size=15)
            ops, boards, players, lasts, caps, oks = [], [], [], [], [], []
            for _ in range(220 if rule == rules.PENTE else 140):
                undo = len(g.move_history) > 0 and rng.random() < 0.3
                if undo:
                    g.undo_move()
                    ops.append(-1); oks.append(1)
                else:
                    empties = np.argwhere(np.asarray(g.board) == 0)
                    if rng.random() < 0.04 or len(empties) == 0:        # an occupied or off-board cell: refused, nothing changes
                        r, c = (int(rng.integers(0, 15)), int(rng.integers(0, 15))) if rng.random() < 0.5 else (15, 3)
                        if 0 <= r < 15 and g.board[r][c] == 0:
                            r, c = 15, 3
                    elif rule == rules.PENTE and rng.random() < 0.7 and _capture_moves(np.asarray(g.board), g.current_player):
                        cm = _capture_moves(np.asarray(g.board), g.current_player)       # a move that captures: its undo is the case to pin
                        r, c = cm[int(rng.integers(0, len(cm)))]
                    elif rule == rules.PENTE and rng.random() < 0.5:
                        # next to an existing stone: makes custodial captures (and their undo) likely
                        stones = np.argwhere(np.asarray(g.board) != 0)
                        r, c = empties[int(rng.integers(0, len(empties)))]
                        if len(stones):
                            sr, sc = stones[int(rng.integers(0, len(stones)))]
                            cand = [(sr + dr, sc + dc) for dr in (-3, -1, 0, 1, 3) for dc in (-3, -1, 0, 1, 3)
                                    if 0 <= sr + dr < 15 and 0 <= sc + dc < 15 and g.board[sr + dr][sc + dc] == 0]
                            if cand:
                                r, c = cand[int(rng.integers(0, len(cand)))]
                    else:
                        r, c = empties[int(rng.integers(0, len(empties)))]
                    ok = g.do_move((int(r), int(c)))
                    ops.append(int(r) * 16 + int(c)); oks.append(int(bool(ok)))
                boards.append(np.asarray(g.board, dtype=np.int8).copy())
                players.append(int(g.current_player))
                lasts.append(-1 if g.last_move is None else int(g.last_move[0]) * 15 + int(g.last_move[1]))
                caps.append(ref_caps(g))
            n_undo_caps = sum(1 for i, o in enumerate(ops) if o == -1 and i > 0 and caps[i] != caps[i - 1])
            print(f"undo trace {tag}{k}: {len(ops)} ops, {ops.count(-1)} undos, {n_undo_caps} of them undo a capture")
            out[f"{tag}{k}/ops"] = np.array(ops, np.int32)          # -1 = undo_move, else row * 16 + col (row 15 = off the board)
            out[f"{tag}{k}/ok"] = np.array(oks, np.int8)
            out[f"{tag}{k}/boards"] = np.stack(boards)
            out[f"{tag}{k}/players"] = np.array(players, np.int8)
            out[f"{tag}{k}/last"] = np.array(lasts, np.int16)
            out[f"{tag}{k}/caps"] = np.array(caps, np.int16)
    np.savez_compressed(os.path.join(OUT, "undo_traces.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["rules", "search", "noise", "misc", "net", "api", "game", "arena", "train", "undo"]
    for w in which:
        print(f"== {w}")
        {"rules": build_rules, "search": build_search, "noise": build_noise, "misc": build_misc, "net": build_net, "api": build_api,
         "game": build_game, "arena": build_arena, "train": build_train, "undo": build_undo}[w]()


if __name__ == "__main__":
    main()
