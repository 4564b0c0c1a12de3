"""Single-game latency of the drop-in ``Player.play`` / ``MCTS.run``: one game, n_sims
simulations per move, deferred-evaluation queue of 32 leaves, so a move is ~n_sims/32 dependent
FILL -> network -> COMMIT rounds.  Prints one JSON line per configuration.

    python tools/player_latency.py [--sims 5000] [--moves 6] [--out profiles/player_latency_r01.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sims", type=int, default=5000)
    ap.add_argument("--moves", type=int, default=6)
    ap.add_argument("--out", default=None)
    ap.add_argument("--fast-warps", type=int, default=0, help="non-parity fast mode: warps per game walking the tree concurrently")
    ap.add_argument("--batch-size", type=int, default=32, help="the reference's MCTS batch_size (deferred-evaluation queue length)")
    args = ap.parse_args()
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.games import Gomoku
    from alphazero_gomoku_b200.network import PyTorchModel
    lines = []
    for blocks, ch in ((3, 64), (6, 128)):
        torch.manual_seed(0)
        net = PyTorchModel(n_res_blocks=blocks, channels=ch, device="cuda:0")
        mcts = m.MCTS(Gomoku, args.sims, net, cpuct=1.0, add_dirichlet_noise=False, fast_warps=args.fast_warps,
                      batch_size=args.batch_size)
        game = Gomoku(15)
        times, evals = [], []
        for ply in range(args.moves + 1):
            e0 = mcts.n_evals
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pi = mcts.run(game, ply)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if ply > 0:                         # ply 0 pays the one-time packing of the weights
                times.append(dt)
                evals.append(mcts.n_evals - e0)
            a = int(np.argmax(pi))
            game.do_move((a // 15, a % 15))
            if game.is_game_over():
                break
        ms = 1e3 * float(np.median(times))
        line = {"path": "MCTS.run single game", "mode": "exact" if args.fast_warps == 0 else f"fast, {args.fast_warps} warps",
                "batch_size": args.batch_size, "net": f"{blocks}x{ch}", "sims_per_move": args.sims,
                "ms_per_move_median": round(ms, 2), "ms_per_move_all": [round(1e3 * t, 2) for t in times],
                "sims_per_s": round(args.sims / (ms * 1e-3), 1), "evals_per_move": int(np.median(evals)),
                "rounds_per_move": int(np.ceil(np.median(evals) / args.batch_size)),
                "ms_per_round": round(ms / max(1.0, np.ceil(np.median(evals) / args.batch_size)), 4)}
        print(json.dumps(line), flush=True)
        lines.append(line)
        mcts.engine.close()
    if args.out:
        with open(args.out, "a") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
