"""BASELINE.json configs[3]: one iteration of self-play -> replay buffer -> Adam training across N B200s,
phase by phase, with the SURVEY 8(d) settings (6x128 net, batch 128 per GPU, Adam lr 1e-3 wd 1e-4, clip 3.0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/train_loop_bench.py [--games 512] [--sims 200] [--plies 64] [--train-steps 40]

Phases (every time is the MAX over ranks of a CUDA-event interval bracketed by barriers):
  self-play   each rank plays `--games` concurrent games for `--plies` plies on its own GPU (no collective);
  gather      the finished games of all ranks are all-gathered as packed plies (976 B each) and expanded to the 8
              symmetries on every rank into its HBM-resident replay buffer (NCCL);
  train       `--train-steps` data-parallel Adam steps: same global batch drawn on every rank, each rank takes
              its slice, runs the tensor-core forward/backward (CUDA graph), flat fp32 gradient all-reduce
              (1.89 M elements), clip + Adam (second graph): identical update on every rank.
Prints one JSON line on rank 0 and checks that all ranks end with bit-identical weights."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, dev):
    """Run fn between barriers; return (result, max-over-ranks milliseconds)."""
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--games", type=int, default=512)
    ap.add_argument("--sims", type=int, default=200)
    ap.add_argument("--plies", type=int, default=64)
    ap.add_argument("--train-steps", type=int, default=40)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--batch-per-gpu", type=int, default=128)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    world, rank = dist.get_world_size(), dist.get_rank()
    from alphazero_gomoku_b200 import train as tr
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import SelfPlay

    torch.manual_seed(0)
    model = PyTorchModel(n_res_blocks=args.blocks, channels=args.channels, device=str(dev))
    tr.broadcast_model(model)
    sp = SelfPlay(model, n_games=args.games, n_sims=args.sims, noise=True, alpha=0.05, eps=0.15, noise_plies=10,
                  temp_threshold=10.0, example_capacity=args.games * (args.plies + 2), seed=12345,
                  game_base=rank * args.games, node_capacity=4096, device=str(dev), packed_examples=True)
    sp.step()                                                   # warm-up ply (weight packing, allocator)
    sims0 = sp.total_sims

    def selfplay():
        for _ in range(args.plies):
            sp.step()
    _, t_sp = timed(selfplay, dev)
    sims = (sp.total_sims - sims0) * world
    from alphazero_gomoku_b200.selfplay import expand_examples
    packed_local = sp.drain_packed()                           # plies of the games that ended, 976 bytes each
    synthetic = 0
    if packed_local.shape[0] * 8 < args.batch_per_gpu:         # too short a run for games to end: pad, and say so
        synthetic = args.games
        extra = torch.zeros((synthetic, 244), dtype=torch.int32, device=dev)
        extra[:, 16] = 1
        extra[:, 18:243] = torch.full((225,), 1.0 / 225).view(torch.int32).to(dev)
        packed_local = torch.cat([packed_local, extra])

    def exchange():
        return expand_examples(tr.gather_rows(packed_local), True)     # all-gather packed plies, expand 8 symmetries locally
    del_me = exchange()                                                # NCCL connection set-up and the allocator's first
    del del_me                                                         # cudaMalloc of the row buffer are not the exchange
    rows, t_gather = timed(exchange, dev)
    del rows
    gathered, t_allgather = timed(lambda: tr.gather_rows(packed_local), dev)     # the two halves on their own
    rows, t_expand = timed(lambda: expand_examples(gathered, True), dev)
    buf = tr.DeviceReplayBuffer(max(rows.shape[0], 1), dev)
    buf.add_rows(rows)
    gen = torch.Generator(device=dev)
    gen.manual_seed(17)
    B = args.batch_per_gpu * world

    reduce = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if world > 1 else None

    def step():
        states, pis, zs = buf.sample(B, generator=gen)
        sl = slice(rank * args.batch_per_gpu, (rank + 1) * args.batch_per_gpu)
        # train_batch_dp without its per-step host read of the losses: forward/backward graph, all-reduce, clip + Adam graph
        return model.train_batch_async(states[sl], pis[sl], zs[sl], world=world, reduce_grads=reduce)
    for _ in range(3):
        step()

    def train():
        for _ in range(args.train_steps):
            last = step()
        return last
    losses, t_train = timed(train, dev)
    losses = dict(zip(("policy_loss", "value_loss"), losses.tolist()))
    tr.sync_batchnorm_buffers(model)
    # gradient all-reduce alone, same element count
    n_param = sum(p.numel() for p in model.net.parameters())
    flat = torch.zeros(n_param, device=dev)

    def allreduce():
        for _ in range(20):
            dist.all_reduce(flat)
    _, t_ar = timed(allreduce, dev)
    w = torch.cat([p.detach().reshape(-1) for p in model.net.parameters()])
    parts = [torch.empty_like(w) for _ in range(world)]
    dist.all_gather(parts, w)
    same = all(torch.equal(parts[0], q) for q in parts[1:])
    if rank == 0:
        print(json.dumps({
            "workload": "self-play -> replay buffer -> Adam train loop (BASELINE configs[3])", "n_gpus": world,
            "net": f"{args.blocks}x{args.channels}", "games_per_gpu": args.games, "sims_per_move": args.sims, "plies": args.plies,
            "selfplay_ms": round(t_sp, 1), "selfplay_sims_per_s": round(sims / (t_sp * 1e-3), 1),
            "examples_gathered": int(rows.shape[0]), "synthetic_plies_per_gpu": synthetic, "gather_ms": round(t_gather, 2),
            "gather_allgather_ms": round(t_allgather, 2), "gather_expand_ms": round(t_expand, 2),
            "exchange": "all-gather of packed plies (976 B each) + local 8-symmetry expansion",
            "exchanged_bytes": int(rows.shape[0] // 8 * 976), "expanded_bytes": int(rows.numel() * 4),
            "train_steps": args.train_steps, "global_batch": B, "train_ms_per_step": round(t_train / args.train_steps, 3),
            "train_positions_per_s": round(B * args.train_steps / (t_train * 1e-3), 1),
            "grad_allreduce_ms": round(t_ar / 20, 4), "grad_elements": n_param, "last_losses": losses,
            "weights_identical_on_all_ranks": same}), flush=True)
    dist.barrier()
    sp.close()
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
