"""GPU: bitboard rule kernels through the C ABI, bit-exact against the golden traces of the
reference and against the oracle on seeded random play (both games)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import rules as orules

pytestmark = pytest.mark.gpu


def _engine_rules(rule):
    import alphazero_gomoku_b200 as m
    return m.Rules(rule, "cuda:0")


def test_golden_traces_lockstep():
    """All traces of one rule advance in lock step as one batch; shorter traces repeat an
    occupied-cell move (which must be rejected and change nothing)."""
    z = load_golden("rules_traces.npz")
    for rule in (0, 1):
        names = [str(n) for n, r in zip(z["names"], z["rules"]) if int(r) == rule]
        R = _engine_rules(rule)
        n = len(names)
        pos = R.pack(np.zeros((n, 225), np.int8), np.ones(n, np.int32), -np.ones(n, np.int32), np.zeros((n, 2), np.int32), np.zeros(n, np.int32))
        T = max(len(z[f"{k}/moves"]) for k in names)
        for t in range(T):
            acts = np.full(n, -7, np.int32)
            for i, k in enumerate(names):
                mv = z[f"{k}/moves"]
                if t < len(mv):
                    r, c = int(mv[t][0]), int(mv[t][1])
                    acts[i] = r * 15 + c if (0 <= r < 15 and 0 <= c < 15) else -1 - abs(r) - abs(c)
            st = R.play(pos, torch.from_numpy(acts).cuda()).cpu().numpy()
            boards, players, lasts, caps, plies = [x.cpu().numpy() for x in R.unpack(pos)]
            legal = R.legal(pos).cpu().numpy()
            enc = R.encode(pos).cpu().numpy()
            for i, k in enumerate(names):
                if t >= len(z[f"{k}/moves"]):
                    continue
                assert np.array_equal(boards[i], z[f"{k}/boards"][t]), (k, t)
                assert players[i] == z[f"{k}/players"][t] and lasts[i] == z[f"{k}/lasts"][t], (k, t)
                assert caps[i].tolist() == z[f"{k}/caps"][t].tolist(), (k, t)
                assert (st[i] & 3) == z[f"{k}/winners"][t], (k, t)
                assert bool(st[i] & 4) == bool(z[f"{k}/overs"][t]), (k, t)
                assert bool(st[i] & 8) == (not bool(z[f"{k}/oks"][t])), (k, t)
                assert np.array_equal(legal[i], (z[f"{k}/boards"][t] == 0).astype(np.float32))
                b = z[f"{k}/boards"][t].reshape(15, 15)
                me = players[i]
                assert np.array_equal(enc[i, 0], (b == me).astype(np.float32))
                assert np.array_equal(enc[i, 1], (b == 3 - me).astype(np.float32))
                assert np.all(enc[i, 2] == 1.0)


@pytest.mark.parametrize("rule", [0, 1])
def test_random_play_vs_oracle(rule):
    """4096 seeded games played to the end, every ply compared with the oracle
    (board, side, captures, winner, game-over, rejection of illegal moves)."""
    n = 4096
    rng = np.random.default_rng(100 + rule)
    R = _engine_rules(rule)
    pos = R.pack(np.zeros((n, 225), np.int8), np.ones(n, np.int32))
    ref = [orules.Position(rule) for _ in range(n)]
    done = np.zeros(n, bool)
    plies_checked = 0
    for t in range(260):
        acts = np.full(n, -1, np.int32)
        for i in range(n):
            if done[i]:
                continue
            if rng.random() < 0.02:
                acts[i] = int(rng.integers(-5, 240))          # maybe occupied / off board
            else:
                e = np.flatnonzero(ref[i].cells == 0)
                # cluster moves near the last stone so lines and captures happen
                if ref[i].last >= 0 and rng.random() < 0.8:
                    r, c = divmod(ref[i].last, 15)
                    near = [a for a in e if abs(a // 15 - r) <= 2 and abs(a % 15 - c) <= 2]
                    e = np.array(near) if near else e
                acts[i] = int(e[int(rng.integers(0, len(e)))])
        st = R.play(pos, torch.from_numpy(acts).cuda()).cpu().numpy()
        boards, players, lasts, caps, plies = [x.cpu().numpy() for x in R.unpack(pos)]
        for i in range(n):
            if done[i]:
                continue
            ok = orules.play(ref[i], int(acts[i]))
            assert bool(st[i] & 8) == (not ok), (i, t)
            assert np.array_equal(boards[i], ref[i].cells), (i, t)
            assert players[i] == ref[i].player and lasts[i] == ref[i].last and caps[i].tolist() == ref[i].caps
            assert plies[i] == ref[i].plies
            assert (st[i] & 3) == orules.winner(ref[i]) and bool(st[i] & 4) == orules.game_over(ref[i])
            plies_checked += 1
            if st[i] & 4:
                done[i] = True
        if done.all():
            break
    assert done.all() and plies_checked > 100000


def test_host_entry_point():
    """azg_rules_play_host: host buffers in, host buffers out (the Python game shims use it)."""
    import ctypes as C
    import alphazero_gomoku_b200 as m
    n = 3
    boards = np.zeros((n, 225), np.int8)
    boards[1, 112] = 1
    players = np.array([1, 2, 1], np.int32)
    lasts = np.array([-1, 112, -1], np.int32)
    caps = np.zeros((n, 2), np.int32)
    plies = np.array([0, 1, 0], np.int32)
    acts = np.array([0, 112, 224], np.int32)
    status = np.zeros(n, np.int32)
    m._lib.check(m.lib.azg_rules_play_host(0, 0, m._lib.ptr(boards), m._lib.ptr(players), m._lib.ptr(lasts), m._lib.ptr(caps),
                                           m._lib.ptr(plies), m._lib.ptr(acts), m._lib.ptr(status), n))
    assert boards[0, 0] == 1 and players[0] == 2 and status[0] == 0
    assert status[1] & 8 and players[1] == 2 and boards[1, 112] == 1
    assert boards[2, 224] == 1 and lasts[2] == 224 and plies[2] == 1
