"""Throughput of the rule kernels (HBM roofline): n positions, CUDA-event timing, algorithmic bytes.

    python tools/rules_bench.py [--n 4194304]

Per position: play reads + writes the 96-byte packed record and a 4-byte action, writes 4 status bytes
(200 B); legal reads 96 B and writes 900 B; encode reads 96 B and writes 2 700 B."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 22)
    args = ap.parse_args()
    import alphazero_gomoku_b200 as m
    n = args.n
    peak = 6436.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device="cuda").manual_seed(1)
    for rule, name in ((0, "gomoku"), (1, "pente")):
        R = m.Rules(rule, "cuda:0")
        r = torch.rand((n, 225), device="cuda", generator=g)
        boards = torch.zeros((n, 225), dtype=torch.int8, device="cuda")
        boards[r < 0.15] = 1
        boards[(r >= 0.15) & (r < 0.30)] = 2
        del r
        players = torch.randint(1, 3, (n,), device="cuda", generator=g, dtype=torch.int32)
        pos0 = R.pack(boards, players)
        legal = R.legal(pos0)
        acts = torch.multinomial(legal[: 1 << 20], 1, generator=g).squeeze(1).to(torch.int32).repeat(n >> 20)
        del boards
        pos = pos0.clone()

        def play():
            pos.copy_(pos0)            # restore (the copy is timed separately and subtracted)
            R.play(pos, acts)
        t_copy = timed(lambda: pos.copy_(pos0))
        t_play = timed(play) - t_copy
        t_legal = timed(lambda: R.legal(pos0))
        t_enc = timed(lambda: R.encode(pos0))
        for kern, ms, bytes_per in (("azg_rules_play", t_play, 200), ("azg_rules_legal", t_legal, 996), ("azg_rules_encode", t_enc, 2796)):
            gbs = n * bytes_per / (ms * 1e-3) / 1e9
            print(json.dumps({"kernel": kern, "rule": name, "positions": n, "ms": round(ms, 4), "positions_per_s": round(n / (ms * 1e-3)),
                              "algorithmic_bytes_per_position": bytes_per, "achieved_GBs": round(gbs, 1), "peak_GBs": peak,
                              "frac": round(gbs / peak, 3)}), flush=True)
        del pos, pos0, legal, acts
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
