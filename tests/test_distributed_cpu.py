"""CPU, world_size 2 over gloo: the host-side multi-rank logic of the train loop - ragged example
all-gather and gradient averaging (every rank must apply the identical update, equal to the
average of the per-rank gradients, clipped at 3.0, then Adam)."""
import copy
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from conftest import ROOT


class _CpuModel:
    """The attributes train_batch_dp uses, on the CPU (PyTorchModel itself refuses to run without CUDA)."""

    def __init__(self, net):
        self.net = net
        self.device = "cpu"
        self.optimizer = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
        self.value_loss_fn = torch.nn.MSELoss()
        self.policy_loss_fn = torch.nn.KLDivLoss(reduction="batchmean")
        self.invalidated = 0

    def invalidate(self):
        self.invalidated += 1


def _batch(seed, n):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand((n, 3, 15, 15), generator=g) < 0.2).float()
    x[:, 2] = 1.0
    pi = torch.softmax(torch.randn((n, 225), generator=g), dim=1)
    z = torch.randint(-1, 2, (n, 1), generator=g).float()
    return x, pi, z


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    import alphazero_gomoku_b200.network as mynet
    import alphazero_gomoku_b200.train as tr

    # ragged all-gather
    rows = torch.full((3 + 2 * rank, 901), float(rank + 1))
    allrows = tr.gather_rows(rows)
    assert allrows.shape == (3 + 5, 901)
    assert torch.all(allrows[:3] == 1.0) and torch.all(allrows[3:] == 2.0)
    packed = torch.full((2 + rank, 244), rank + 7, dtype=torch.int32)             # packed plies travel as int32 words
    allpacked = tr.gather_rows(packed)
    assert allpacked.dtype == torch.int32 and allpacked.shape == (5, 244)
    assert torch.all(allpacked[:2] == 7) and torch.all(allpacked[2:] == 8)

    torch.manual_seed(0)
    net = mynet.AlphaZeroNet(n_res_blocks=1, channels=16)
    model = _CpuModel(net)
    ref = _CpuModel(copy.deepcopy(net))
    x, pi, z = _batch(5, 16)
    sl = slice(rank * 8, rank * 8 + 8)
    losses = tr.train_batch_dp(model, x[sl], pi[sl], z[sl])
    assert np.isfinite(losses["total_loss"])

    # reference: both slices on one process, gradients averaged, same clip and Adam step
    ref.net.train()
    grads = []
    for r in range(world):
        ref.optimizer.zero_grad()
        s = slice(r * 8, r * 8 + 8)
        snap = copy.deepcopy(ref.net.state_dict())          # BatchNorm running stats must not accumulate across slices
        lo, v = ref.net(x[s])
        (ref.policy_loss_fn(F.log_softmax(lo, dim=1), pi[s]) + ref.value_loss_fn(v, z[s])).backward()
        grads.append([p.grad.clone() for p in ref.net.parameters()])
        if r != rank:
            ref.net.load_state_dict(snap)
    ref.optimizer.zero_grad()
    for p, g0, g1 in zip(ref.net.parameters(), *grads):
        p.grad = (g0 + g1) / 2
    torch.nn.utils.clip_grad_norm_(ref.net.parameters(), 3.0)
    ref.optimizer.step()
    worst = max(float((a - b).abs().max()) for a, b in zip(model.net.parameters(), ref.net.parameters()))
    # identical parameters on every rank
    flat = torch.cat([p.detach().reshape(-1) for p in model.net.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    same = bool(torch.equal(parts[0], parts[1]))
    # BatchNorm running statistics: every rank normalised its own slice, so they differ until they are synchronised;
    # afterwards the WHOLE state_dict (buffers included) is bit-identical on all ranks and the engine cache is dropped
    def buffers():
        return torch.cat([b.detach().reshape(-1).double() for b in model.net.buffers()])
    parts = [torch.empty_like(buffers()) for _ in range(world)]
    dist.all_gather(parts, buffers())
    assert not torch.equal(parts[0], parts[1])
    tr.sync_batchnorm_buffers(model)
    assert model.invalidated == 1
    parts = [torch.empty_like(buffers()) for _ in range(world)]
    dist.all_gather(parts, buffers())
    same = same and bool(torch.equal(parts[0], parts[1]))
    assert int(model.net.bn.num_batches_tracked) == 1
    # broadcast_model goes through the tensors (not .data) and drops the packed-weight cache
    with torch.no_grad():
        for p in model.net.parameters():
            p.add_(float(rank))
    tr.broadcast_model(model, 0)
    assert model.invalidated == 2
    flat = torch.cat([p.detach().reshape(-1) for p in model.net.parameters()])
    parts = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(parts, flat)
    same = same and bool(torch.equal(parts[0], parts[1]))
    # replicated phase (small batches: every rank trains the WHOLE batch, train.DP_MIN_POSITIONS_PER_RANK): the ranks
    # step on their own - here even on different data - and rank 0's parameters, buffers and Adam state are broadcast after
    xr, pir, zr = _batch(100 + rank, 8)
    lo = tr.train_step_device(model, xr, pir, zr)                      # world = 1: no collective inside
    assert lo.shape == (2,) and bool(torch.isfinite(lo).all())
    def moments():
        return torch.cat([model.optimizer.state[p][k].detach().reshape(-1).double() for p in model.net.parameters()
                          for k in ("exp_avg", "exp_avg_sq")] +
                         [torch.as_tensor(model.optimizer.state[p]["step"]).double().reshape(-1) for p in model.net.parameters()])
    parts = [torch.empty_like(moments()) for _ in range(world)]
    dist.all_gather(parts, moments())
    assert not torch.equal(parts[0], parts[1])
    tr.broadcast_model(model, 0)
    tr.broadcast_optimizer(model, 0)
    parts = [torch.empty_like(moments()) for _ in range(world)]
    dist.all_gather(parts, moments())
    same = same and bool(torch.equal(parts[0], parts[1]))
    full = torch.cat([t.detach().reshape(-1).double() for t in list(model.net.parameters()) + list(model.net.buffers())])
    parts = [torch.empty_like(full) for _ in range(world)]
    dist.all_gather(parts, full)
    same = same and bool(torch.equal(parts[0], parts[1]))
    out.put((rank, worst, same))
    dist.destroy_process_group()


def test_gather_and_gradient_average_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, worst, same in res:
        assert same, "ranks diverged"
        assert worst < 1e-5, (rank, worst)
