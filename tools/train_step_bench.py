"""Throughput of the tensor-core training step (PyTorchModel.train_batch, network.py:199-235).

    python tools/train_step_bench.py [--blocks 6 --channels 128] [--batches 128,256,512,1024] [--steps 30]

For every batch size: positions/s and ms/step of (a) the CUDA step replayed from CUDA graphs, (b) the CUDA step
launched kernel by kernel, (c) the torch autograd formulation of the same step (cuDNN / cuBLAS library kernels, fp32 -
what round 1 shipped), each timed with CUDA events over --steps steps after 5 warm-up steps, inputs resident in
HBM.  Useful FLOPs per position = 3 x the forward 3x3-trunk FLOPs (forward, input gradient, weight gradient) on the
225 real pixels; the fraction is against MEASURED_PEAKS.json's sustained bf16 rate.  One JSON line per batch size.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, steps):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--batches", default="128,256,512,1024")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--skip-autograd", action="store_true")
    args = ap.parse_args()
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import trunk_flops
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    flops_pos = 3 * trunk_flops(args.channels) * 2 * args.blocks
    g = torch.Generator(device="cuda").manual_seed(1)
    for B in [int(b) for b in args.batches.split(",")]:
        x = (torch.rand((B, 3, 15, 15), device="cuda", generator=g) < 0.15).float()
        x[:, 1] *= (1 - x[:, 0])
        x[:, 2] = 1.0
        pi = torch.softmax(torch.randn((B, 225), device="cuda", generator=g), dim=1)
        z = torch.randint(-1, 2, (B, 1), device="cuda", generator=g).float()
        out = {"net": f"{args.blocks}x{args.channels}", "batch": B, "steps": args.steps}
        for tag, graphs in (("cuda_graph", True), ("cuda_launches", False)):
            torch.manual_seed(0)
            m = PyTorchModel(n_res_blocks=args.blocks, channels=args.channels, device="cuda:0")
            m.train_graphs = graphs
            ms = timed(lambda: m.train_batch_async(x, pi, z), args.steps)
            st = m._trainer.check()
            out[tag] = {"ms_per_step": round(ms, 4), "positions_per_s": round(B / ms * 1e3, 1),
                        "useful_tflops": round(B * flops_pos / (ms * 1e-3) / 1e12, 1),
                        "frac_of_sustained_bf16_peak": round(B * flops_pos / (ms * 1e-3) / 1e12 / peak, 4)}
            out["trainer_gb"] = round(m._trainer.memory_bytes / 1e9, 2)
            out["grad_norm_last"] = round(st["grad_norm"], 4)
            m._trainer.close()
            del m
        if not args.skip_autograd:
            torch.manual_seed(0)
            m = PyTorchModel(n_res_blocks=args.blocks, channels=args.channels, device="cuda:0")

            def lib_step():
                m.net.train()
                m.optimizer.zero_grad()
                pl, vl = m.losses(x, pi, z)
                (pl + vl).backward()
                torch.nn.utils.clip_grad_norm_(m.net.parameters(), 3.0)
                m.optimizer.step()
            ms = timed(lib_step, args.steps)
            out["torch_autograd_fp32"] = {"ms_per_step": round(ms, 4), "positions_per_s": round(B / ms * 1e3, 1)}
            out["speedup_vs_autograd"] = round(out["torch_autograd_fp32"]["ms_per_step"] / out["cuda_graph"]["ms_per_step"], 2)
            del m
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
