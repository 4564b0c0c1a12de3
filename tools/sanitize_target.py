"""Small end-to-end workload that launches every kernel of the library at least once (rules, search with tree reuse
and GC, both rules, noise, network at 64 and 128 channels, batched self-play) at sizes a checker tool finishes in
minutes, plus the training step and the packed example format.  Written as a `compute-sanitizer` target (memcheck /
racecheck); profiles/README.md records what the pool allowed.

    python tools/sanitize_target.py [--no-net]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-net", action="store_true", help="skip the tcgen05 network kernels")
    args = ap.parse_args()
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.engine import Rules
    from alphazero_gomoku_b200.games import Gomoku, Pente
    from oracle import fakes

    # rules kernels: random playouts of 64 games per rule
    rng = np.random.default_rng(0)
    for rule in (0, 1):
        r = Rules(rule, "cuda:0")
        pos = r.pack(np.zeros((64, 225), np.int8), [1] * 64, [-1] * 64, [[0, 0]] * 64, [0] * 64)
        for _ in range(40):
            legal = r.legal(pos).cpu().numpy()
            acts = np.array([rng.choice(np.flatnonzero(l)) if l.any() else 0 for l in legal], np.int32)
            r.play(pos, torch.from_numpy(acts).cuda())
            r.status(pos)
        r.encode(pos)
        r.unpack(pos)
    # search kernels with injected priors, tree reuse + GC between moves, both rules, noise on
    for cls in (Gomoku, Pente):
        np.random.seed(1)
        mcts = m.MCTS(cls, 150, fakes.Spiky(), add_dirichlet_noise=True, dirichlet_alpha=0.3, epsilon=0.25,
                      node_capacity=512)
        game = cls(15)
        for ply in range(6):
            pi = mcts.run(game, ply)
            a = int(np.argmax(pi))
            game.do_move((a // 15, a % 15))
        mcts.engine.close()
    if args.no_net:
        print("sanitize target done (no net)")
        return
    # network kernels + batched self-play (Philox noise, sampling, example capture, respawn)
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay
    for blocks, ch in ((1, 64), (1, 128)):
        torch.manual_seed(0)
        net = PyTorchModel(n_res_blocks=blocks, channels=ch, device="cuda:0")
        sp = SelfPlay(net, n_games=4, n_sims=40, node_capacity=512, example_capacity=4096, max_moves=3)
        for _ in range(4):
            sp.step()
        torch.cuda.synchronize()
        sp.close()
    # training step (forward, backward, clip + Adam) at both widths, without CUDA graphs
    for blocks, ch in ((1, 64), (1, 128)):
        torch.manual_seed(0)
        net = PyTorchModel(n_res_blocks=blocks, channels=ch, device="cuda:0")
        net.train_graphs = False
        x = torch.zeros((6, 3, 15, 15), device="cuda")
        x[:, 2] = 1.0
        x[1, 0, 7, 7] = 1.0
        pi = torch.full((6, 225), 1.0 / 225, device="cuda")
        z = torch.tensor([[1.0], [-1.0], [0.0], [1.0], [0.0], [-1.0]], device="cuda")
        for _ in range(2):
            net.train_batch(x, pi, z)
        net._trainer.check()
    # packed example exchange format
    from alphazero_gomoku_b200.selfplay import expand_examples
    sp = SelfPlay(net, n_games=4, n_sims=24, node_capacity=512, example_capacity=256, max_moves=3, packed_examples=True, max_games=5)
    while sp.games_running() > 0:
        sp.step()
    expand_examples(sp.drain_packed(), True)
    torch.cuda.synchronize()
    sp.close()
    print("sanitize target done")


if __name__ == "__main__":
    main()
