"""Randomised pinning of the oracle against the UNMODIFIED reference (build container only: needs /root/reference).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.fuzz_vs_reference [--cases 200] [--seed 0]

Complements the fixed fixtures of make_golden.py: every case draws a rule, a simulation count, a queue length,
cpuct, an injected-prior model, a random legal opening and whether root noise is on, then plays a few consecutive
argmax moves with tree reuse on the reference's MCTS + game classes and on oracle.search / oracle.rules, and demands
bit-equal pi, N[root], evaluation batches and tree sizes after every run.  Nothing is written; the summary line of
the last run is quoted in DESIGN.md.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import argparse
import time

import numpy as np

from oracle.make_golden import MCTS, REF_GAME      # the reference's classes (imported there from /root/reference)
from oracle import fakes, rules
from oracle.search import Search


def one_case(rng: np.random.Generator, idx: int):
    rule = int(rng.integers(0, 2))
    n_sims = int(rng.choice([1, 2, 7, 31, 32, 33, 50, 64, 65, 100, 150, 257]))
    queue = int(rng.choice([1, 2, 5, 8, 16, 32]))
    cpuct = float(rng.choice([0.5, 1.0, 1.2, 2.5, 4.0]))
    model_name = str(rng.choice(sorted(fakes.BY_NAME)))
    noise = bool(rng.integers(0, 2))
    alpha, eps = float(rng.choice([0.03, 0.3, 1.0])), float(rng.choice([0.03, 0.25]))
    n_first = int(rng.choice([0, 2, 10]))
    g = REF_GAME[rule](15)
    pos = rules.Position(rule)
    for _ in range(int(rng.integers(0, 30))):               # random legal opening (stops before a finished game)
        empties = np.flatnonzero(pos.cells == 0)
        mv = int(empties[int(rng.integers(0, len(empties)))])
        trial_g, trial_p = g.clone(), pos.copy()
        trial_g.do_move(divmod(mv, 15)); rules.play(trial_p, mv)
        if trial_g.is_game_over():
            break
        g, pos = trial_g, trial_p
    ref_model, orc_model = fakes.BY_NAME[model_name](), fakes.BY_NAME[model_name]()
    ref = MCTS(REF_GAME[rule], n_sims, ref_model, cpuct=cpuct, batch_size=queue, dirichlet_alpha=alpha, epsilon=eps,
               apply_dirichlet_n_first_moves=n_first, add_dirichlet_noise=noise)
    orc = Search(rule, n_sims, orc_model, cpuct=cpuct, queue_len=queue, alpha=alpha, eps=eps, noise_plies=n_first, noise=noise)
    sims = 0
    for mv_no in range(int(rng.integers(1, 5))):
        if g.is_game_over():
            break
        seed = int(rng.integers(0, 2**31))
        np.random.seed(seed); pi_ref = ref.run(g, len(g.move_history))
        np.random.seed(seed); pi_orc = orc.run(pos, pos.plies)
        key = ref._state_key(g)
        tag = (idx, rule, n_sims, queue, cpuct, model_name, noise, mv_no)
        assert key == pos.key(), tag
        assert pi_ref.dtype == pi_orc.dtype and np.array_equal(pi_ref, pi_orc), tag
        assert np.array_equal(np.asarray(ref.N[key]), np.asarray(orc.Nv[key])), tag
        assert len(ref.P) == len(orc.P) and (ref_model.rows, ref_model.calls) == (orc_model.rows, orc_model.calls), tag
        sims += n_sims
        a = int(np.argmax(pi_ref))
        g.do_move(divmod(a, 15)); rules.play(pos, a)
        assert g.get_winner() == rules.winner(pos) and g.is_game_over() == rules.game_over(pos), tag
    return sims


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    sims = sum(one_case(rng, i) for i in range(args.cases))
    print(f"fuzz_vs_reference: {args.cases} random cases, {sims} simulations, oracle == reference bit for bit "
          f"(seed {args.seed}, {time.time() - t0:.0f} s)")


if __name__ == "__main__":
    main()
