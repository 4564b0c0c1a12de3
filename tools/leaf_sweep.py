"""Leaf-evaluation throughput sweep (BASELINE.json configs[4]): batch 1..16384 positions, 6x128 and
10x256 ResNet, against the measured bf16 tensor roofline.  Positions: k ~ U[0,120] random legal plies
from the empty board (numpy default_rng(0)), random-init weights (torch.manual_seed(0)).

    python tools/leaf_sweep.py [--out profiles/leaf_sweep_r01.jsonl]

Each batch size: 3 warm-up passes, then `reps` timed passes of PyTorchModel-equivalent inference
(planes on the device -> probs/values on the device) with CUDA events; a 256 MB buffer is rewritten
between passes so that small batches do not run out of a warm L2."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def positions(n, seed=0):
    rng = np.random.default_rng(seed)
    X = np.zeros((n, 3, 15, 15), np.float32)
    X[:, 2] = 1.0
    base = min(n, 2048)                     # distinct positions; larger batches tile them
    for i in range(base):
        k = int(rng.integers(0, 121))
        cells = rng.permutation(225)[:k]
        for j, c in enumerate(cells):
            X[i, (j + k) % 2, c // 15, c % 15] = 1.0          # mover / opponent stones alternate back from the last ply
    for i in range(base, n):
        X[i] = X[i % base]
    return X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "leaf_sweep_r01.jsonl"))
    ap.add_argument("--max-batch", type=int, default=16384)
    args = ap.parse_args()
    import alphazero_gomoku_b200.network as mynet
    from alphazero_gomoku_b200.nn_engine import NetEngine
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))           # burst figure: kernels timed alone
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    lines = []
    for blocks, ch, flops in ((6, 128, 798221828), (10, 256, 5312103428)):
        torch.manual_seed(0)
        net = mynet.AlphaZeroNet(n_res_blocks=blocks, channels=ch)
        eng = NetEngine(blocks, ch, "cuda:0", max_batch=args.max_batch)
        eng.load_state_dict(net.state_dict())
        X = torch.from_numpy(positions(args.max_batch)).cuda()
        B = 1
        while B <= args.max_batch:
            x = X[:B].contiguous()
            for _ in range(3):
                eng.forward(x)
            reps = 20 if B <= 1024 else 6
            total = 0.0
            for _ in range(reps):
                flush.fill_(1.0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                eng.forward(x)
                e1.record()
                torch.cuda.synchronize()
                total += e0.elapsed_time(e1)
            ms = total / reps
            tf = B * flops / (ms * 1e-3) / 1e12
            line = {"net": f"{blocks}x{ch}", "batch": B, "ms": round(ms, 4), "evals_per_s": round(B / (ms * 1e-3), 1),
                    "tflops": round(tf, 2), "frac_of_measured_bf16_burst": round(tf / peak, 4), "peak_tflops": peak}
            print(json.dumps(line), flush=True)
            lines.append(line)
            B *= 2
        eng.close()
    with open(args.out, "w") as f:
        for l in lines:
            f.write(json.dumps(l) + "\n")


if __name__ == "__main__":
    main()
