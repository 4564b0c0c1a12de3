"""Multi-GPU smoke run of the self-play -> replay -> train loop (BASELINE.json configs[3]):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_dp_smoke.py

Every rank plays its share of the games on its own GPU, examples are all-gathered, the Adam step
uses NCCL-averaged gradients (AZG_SMOKE_BATCH_PER_RANK=64, the default) or - below train.DP_MIN_POSITIONS_PER_RANK
positions per rank, e.g. AZG_SMOKE_BATCH_PER_RANK=16 - every rank trains the whole batch and rank 0's result is
broadcast; at the end all ranks must hold a bit-identical state_dict (weights and BatchNorm buffers)."""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from alphazero_gomoku_b200 import train as tr
    out = tempfile.mkdtemp(prefix=f"azg_dp_{dist.get_rank()}_")
    best = tr.train_alphazero(num_iterations=2, games_per_iteration=8 * dist.get_world_size(), n_simulations=32, buffer_size=20000,
                              batch_size=int(os.environ.get("AZG_SMOKE_BATCH_PER_RANK", "64")) * dist.get_world_size(), epochs_per_iter=1, temp_threshold=8, eval_games=4,
                              eval_mcts_simulations=16, win_rate_threshold=0.0, cpuct=1.2, model_dir=out, dirichlet_alpha=0.3,
                              dirichlet_epsilon=0.25, dirichlet_n_moves=30, n_res_blocks=1, channels=64)
    # parameters AND buffers (BatchNorm running statistics are averaged over the ranks after every training phase)
    flat = torch.cat([p.detach().reshape(-1).double() for p in list(best.net.parameters()) + list(best.net.buffers())])
    parts = [torch.empty_like(flat) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, flat)
    same = all(torch.equal(parts[0], q) for q in parts[1:])
    if dist.get_rank() == 0:
        print("DP_TRAIN_SMOKE", "world", dist.get_world_size(), "weights_identical", same, "param_sum", float(flat.double().sum()))
    dist.barrier()
    dist.destroy_process_group()
    if not same:
        sys.exit(1)


if __name__ == "__main__":
    main()
