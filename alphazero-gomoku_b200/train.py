"""Self-play -> replay buffer -> train loop on the B200 engine (SURVEY 8f "next" rows).

Keeps the reference's call shapes (train.py of the reference): ``softmax_temperature``,
``sample_action_from_pi`` (:252-266), ``ReplayBuffer`` (:272-296), ``save_replay_buffer`` /
``load_replay_buffer`` (:302-354, same pickle dictionary), ``play_game_and_collect`` (:360-412),
``evaluate_models`` (:418-486) and ``train_alphazero`` with the same keyword arguments (:575-609).
Self-play inside ``train_alphazero`` runs on the batched device driver (``selfplay.SelfPlay``);
with ``torch.distributed`` initialised every rank plays its own games, the examples are
all-gathered and the gradients of ``train_batch`` are all-reduced over NCCL.
"""
from __future__ import annotations

import os
import pickle
import random
import time
from collections import deque
from datetime import datetime
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F

from .games import Gomoku
from .mcts import MCTS
from .network import PyTorchModel
from .selfplay import SelfPlay

GameClass = Gomoku


# ------------------------------------------------------------------ sampling helpers (train.py:252-266)
def softmax_temperature(pi: np.ndarray, temp: float) -> np.ndarray:
    if temp <= 0:
        return pi
    z = np.log(pi + 1e-15) / temp
    e = np.exp(z - np.max(z))
    return e / np.sum(e)


def sample_action_from_pi(pi: np.ndarray, temp: float) -> int:
    if temp == 0:
        return int(np.argmax(pi))
    p = softmax_temperature(pi, temp)
    return int(np.random.choice(len(p), p=p))


# ------------------------------------------------------------------ replay buffer (train.py:272-354)
class ReplayBuffer:
    def __init__(self, capacity: int = 20000):
        self.capacity = capacity
        self.buffer = deque(maxlen=capacity)

    def add(self, examples: List[Tuple[np.ndarray, np.ndarray, float]]):
        for ex in examples:
            self.buffer.append(ex)

    def add_rows(self, rows: torch.Tensor):
        """Rows of ``SelfPlay.drain_examples`` (float32[n, 901]) -> (state, pi, z) tuples."""
        r = rows.detach().cpu().numpy()
        for x in r:
            self.buffer.append((x[:675].reshape(3, 15, 15).copy(), x[675:900].copy(), float(x[900])))

    def sample(self, batch_size: int):
        batch = random.sample(self.buffer, k=batch_size)
        states, pis, zs = zip(*batch)
        return (np.stack(states, axis=0).astype(np.float32), np.stack(pis, axis=0).astype(np.float32),
                np.array(zs, dtype=np.float32).reshape(-1, 1))

    def __len__(self):
        return len(self.buffer)


class DeviceReplayBuffer:
    """The same interface with the rows kept in HBM (float32[capacity, 901] ring: planes 675, pi 225, z):
    ``add_rows`` takes ``SelfPlay.drain_examples()`` without a host copy, ``sample`` returns device
    tensors that ``train_batch`` consumes directly.  ``to_host`` / ``from_host`` convert to and from the
    reference's buffer (and therefore its pickle format)."""

    def __init__(self, capacity: int = 20000, device="cuda"):
        self.capacity = capacity
        self.rows = torch.empty((capacity, 901), dtype=torch.float32, device=device)
        self.size = 0
        self.head = 0                      # next slot to overwrite (oldest first, like deque(maxlen))

    def add_rows(self, rows: torch.Tensor):
        rows = rows.to(self.rows.device, torch.float32)
        if rows.shape[0] >= self.capacity:
            rows = rows[-self.capacity:]
        n = rows.shape[0]
        first = min(n, self.capacity - self.head)
        self.rows[self.head:self.head + first] = rows[:first]
        if n > first:
            self.rows[: n - first] = rows[first:]
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def add(self, examples):
        if examples:
            flat = np.stack([np.concatenate([np.asarray(s, np.float32).reshape(-1), np.asarray(p, np.float32).reshape(-1),
                                             np.array([z], np.float32)]) for s, p, z in examples])
            self.add_rows(torch.from_numpy(flat))

    def sample(self, batch_size: int, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.rows.device, generator=generator)
        r = self.rows[idx]
        return r[:, :675].reshape(-1, 3, 15, 15), r[:, 675:900], r[:, 900:901]

    def __len__(self):
        return self.size

    def to_host(self) -> ReplayBuffer:
        buf = ReplayBuffer(self.capacity)
        order = torch.arange(self.size, device=self.rows.device)
        if self.size == self.capacity:
            order = (order + self.head) % self.capacity          # oldest first
        buf.add_rows(self.rows[order])
        return buf

    @classmethod
    def from_host(cls, buf: ReplayBuffer, device="cuda") -> "DeviceReplayBuffer":
        out = cls(buf.capacity, device)
        out.add(list(buf.buffer))
        return out


def save_replay_buffer(buffer: ReplayBuffer, filepath: str):
    try:
        with open(filepath, "wb") as f:
            pickle.dump({"buffer": list(buffer.buffer), "capacity": buffer.capacity}, f, protocol=pickle.HIGHEST_PROTOCOL)
        print(f"[Buffer] saved: {filepath} ({len(buffer)} samples)")
        return True
    except Exception as e:          # the reference prints and carries on
        print(f"[Buffer] save failed: {e}")
        return False


def load_replay_buffer(filepath: str, capacity: int) -> Optional[ReplayBuffer]:
    if not os.path.exists(filepath):
        print(f"[Buffer] no saved buffer: {filepath}")
        return None
    try:
        with open(filepath, "rb") as f:
            data = pickle.load(f)
        buf = ReplayBuffer(capacity=capacity)
        if data.get("capacity", capacity) != capacity:
            print(f"[Buffer] warning: saved capacity {data.get('capacity')} != configured {capacity}")
        for item in data["buffer"]:
            buf.buffer.append(item)
        print(f"[Buffer] loaded: {filepath} ({len(buf)} samples)")
        return buf
    except Exception as e:
        print(f"[Buffer] load failed: {e}")
        return None


# ------------------------------------------------------------------ one game on the host API (train.py:360-412)
def play_game_and_collect(mcts: MCTS, game, temp_fn, max_moves=225, use_symmetries=True):
    examples = []
    move_number = 0
    while True:
        state_enc = game.get_encoded_state()
        pi = mcts.run(game, len(game.move_history))
        keep = pi.copy()
        action = sample_action_from_pi(pi, temp_fn(move_number))
        if game.get_valid_moves()[action] != 1.0:
            action = int(np.argmax(pi))
        examples.append((state_enc, keep, int(game.current_player)))
        game.do_move(divmod(action, game.size))
        move_number += 1
        if game.is_game_over() or move_number >= max_moves:
            break
    winner = game.get_winner()
    out = []
    for state_enc, pi_vec, player in examples:
        z = 0.0 if winner == 0 else (1.0 if winner == player else -1.0)
        if use_symmetries:
            for s_aug, pi_aug in mcts.symmetries(state_enc, pi_vec):
                out.append((s_aug.astype(np.float32), pi_aug.astype(np.float32), z))
        else:
            out.append((state_enc.astype(np.float32), pi_vec.astype(np.float32), z))
    return out, winner


# ------------------------------------------------------------------ evaluation arena (train.py:418-486)
def draw_first_stones(n_games: int, board_size: int = 15) -> np.ndarray:
    """The arena's opening stones, one per game: uniform over the central 9x9, drawn from ``random`` in the
    reference's order (row, then column; train.py:431-434)."""
    center, radius = board_size // 2, 4
    cells = np.zeros(n_games, np.int32)
    for i in range(n_games):
        r, c = random.randint(center - radius, center + radius), random.randint(center - radius, center + radius)
        cells[i] = r * board_size + c
    return cells


def evaluate_models(model_new: PyTorchModel, model_best: PyTorchModel, game_name: str, n_games: int = 20,
                    n_simulations: int = 100, cpuct: float = 1.0, *, first_stones: Optional[np.ndarray] = None,
                    first_game: int = 0, transcript: Optional[list] = None) -> Tuple[int, float, int]:
    """Same protocol as the reference - random first stone in the central 9x9, the new model starts the
    even games, argmax play without noise, one search tree per (model, game) kept for the whole game -
    with all ``n_games`` games advancing in lock step on two batched engines (one per model).
    ``first_stones`` / ``first_game`` let a caller play a slice of a larger match (the data-parallel arena):
    the opening cells of this slice and the match index of its first game (which decides who starts).
    A model without this package's CUDA evaluator (anything with the reference's ``predict`` protocol, e.g. an
    injected-prior fake) is evaluated through one host round trip per leaf batch.  ``transcript``, if a list,
    receives the move list of every game (opening stone first)."""
    from .engine import SearchEngine
    dev_name = str(getattr(model_new, "device", "cuda"))
    dev = torch.device(dev_name if dev_name.startswith("cuda:") else f"cuda:{torch.cuda.current_device()}")
    G = n_games
    if first_stones is None:
        first_stones = draw_first_stones(G, model_new.board_size)
    if G == 0:
        return 0, 0.0, 0
    nets = [m._ensure_engine() if hasattr(m, "_ensure_engine") else None for m in (model_new, model_best)]      # 0: new model, 1: best model
    limit = min([n.max_batch for n in nets if n is not None] or [1 << 30]) // 32 // 2 * 2
    if G > limit:                # more games than one leaf batch of the models' evaluators holds: play them in even-sized groups
        wins = draws = 0
        for first in range(0, G, limit):
            m = min(limit, G - first)
            w, _, d = evaluate_models(model_new, model_best, game_name, m, n_simulations, cpuct,
                                      first_stones=first_stones[first:first + m], first_game=first_game + first, transcript=transcript)
            wins, draws = wins + w, draws + d
        return wins, wins / float(G), draws
    boards = np.zeros((G, 225), np.int8)
    lasts = np.asarray(first_stones, np.int32).copy()
    boards[np.arange(G), lasts] = 1
    engines = [SearchEngine(0, G, cpuct=cpuct, queue_len=32, node_capacity=max(4096, 8 * n_simulations), noise=False, device=dev)
               for _ in range(2)]
    models = [model_new, model_best]
    played = [np.asarray(first_stones, np.int32).copy()]
    rules = engines[0].rules
    pos = rules.pack(boards, np.full(G, 2, np.int32), lasts, np.zeros((G, 2), np.int32), np.ones(G, np.int32))
    new_starts = ((torch.arange(G, device=dev) + first_game) % 2 == 0)
    alive = torch.ones(G, dtype=torch.bool, device=dev)
    status = torch.zeros(G, dtype=torch.int32, device=dev)
    probs = torch.empty((G * 32, 225), dtype=torch.float32, device=dev)
    reserve = n_simulations + n_simulations // 32 + 8
    for move_number in range(1, 226):
        if not bool(alive.any()):
            break
        players = pos[:, 16]
        new_to_move = ((players == 1) & new_starts) | ((players == 2) & ~new_starts)
        action = torch.full((G,), -1, dtype=torch.int32, device=dev)
        for side, (eng, net, model) in enumerate(zip(engines, nets, models)):
            mask = alive & (new_to_move if side == 0 else ~new_to_move)
            if not bool(mask.any()):
                continue
            eng.set_roots(pos, clear_tree=False)
            eng.advance(torch.full((G,), -1, dtype=torch.int32, device=dev), gc=True, reserve=reserve)
            eng.begin(n_simulations, mask=mask)
            while True:
                n_leaves, n_more, _ = eng.fill()
                if n_leaves > 0 and net is not None:
                    net.forward_leaves(eng, probs)
                    eng.commit(probs, None)
                elif n_leaves > 0:           # the reference's predict protocol on the host (new_mcts_alpha.py:160-161)
                    p, _ = model.predict(eng.leaf_planes(n_leaves).cpu().numpy())
                    probs[:n_leaves] = torch.from_numpy(np.ascontiguousarray(np.asarray(p, np.float32).reshape(n_leaves, 225))).to(dev)
                    eng.commit(probs, None)
                if n_more == 0:
                    break
            pi, _ = eng.result()
            action = torch.where(mask, pi.argmax(dim=1).to(torch.int32), action)
        if transcript is not None:
            played.append(action.cpu().numpy())
        st = rules.play(pos, action)
        status = torch.where(alive, st, status)
        alive = alive & ((st & 4) == 0)
    for e in engines:
        e.close()
    winner = (status & 3).cpu().numpy()
    starts = new_starts.cpu().numpy()
    draws = int((winner == 0).sum())
    new_wins = int((((winner == 1) & starts) | ((winner == 2) & ~starts)).sum())
    if transcript is not None:
        moves = np.stack(played, axis=1)
        transcript.extend([int(a) for a in row if a >= 0] for row in moves)
    return new_wins, new_wins / float(n_games), draws


def evaluate_models_dp(model_new: PyTorchModel, model_best: PyTorchModel, game_name: str, n_games: int = 20,
                       n_simulations: int = 100, cpuct: float = 1.0) -> Tuple[int, float, int]:
    """The arena sharded over the ranks of an initialised process group (the reference's process-pool
    evaluation, train.py:492-569): rank 0 draws the opening stones, every rank plays a contiguous slice of the
    match on its own GPU with its own replica of both models, wins and draws are summed with one all-reduce.
    The result does not depend on the number of ranks.  Without a process group: ``evaluate_models``."""
    world, rank = _world()
    if world == 1:
        return evaluate_models(model_new, model_best, game_name, n_games, n_simulations, cpuct)
    dev = torch.device(f"cuda:{torch.cuda.current_device()}") if dist.get_backend() == "nccl" else torch.device("cpu")
    stones = torch.zeros(n_games, dtype=torch.int32, device=dev)
    if rank == 0:
        stones.copy_(torch.from_numpy(draw_first_stones(n_games, model_new.board_size)))
    dist.broadcast(stones, 0)
    lo, hi = shard_bounds(n_games, world, rank)
    wins, _, draws = evaluate_models(model_new, model_best, game_name, hi - lo, n_simulations, cpuct,
                                     first_stones=stones[lo:hi].cpu().numpy(), first_game=lo)
    tally = torch.tensor([wins, draws], dtype=torch.int64, device=dev)
    dist.all_reduce(tally)
    wins, draws = int(tally[0].item()), int(tally[1].item())
    return wins, wins / float(n_games), draws


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of n items for ``rank``."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def evaluate_models_mp(model_new: PyTorchModel, model_best: PyTorchModel, board_size: int, action_size: int, n_games: int,
                       n_simulations: int, cpuct: float, *, model_dir: str, num_workers: int, games_per_task: int = 1,
                       device: str = "cpu", base_seed: int = 54321, torch_threads: int = 1) -> Tuple[int, float, int]:
    """Signature of the reference's process-pool arena (train.py:492-569).  The pool, the checkpoint hand-off
    and the per-task seeds are replaced by device batching: all games advance in lock step on the GPU, so
    ``model_dir`` / ``num_workers`` / ``games_per_task`` / ``device`` / ``torch_threads`` are accepted and ignored;
    ``base_seed`` seeds the first-stone draws."""
    state = random.getstate()
    random.seed(base_seed)
    try:
        return evaluate_models(model_new, model_best, "gomoku", n_games=n_games, n_simulations=n_simulations, cpuct=cpuct)
    finally:
        random.setstate(state)


def evaluate_models_serial(model_new: PyTorchModel, model_best: PyTorchModel, game_name: str, n_games: int = 20,
                           n_simulations: int = 100, cpuct: float = 1.0) -> Tuple[int, float, int]:
    """The reference's loop verbatim in shape (one game at a time through the drop-in ``MCTS``); kept as
    the cross-check of the batched arena (same seeds -> same results)."""
    new_wins = draws = 0
    for i in range(n_games):
        game = GameClass(size=model_new.board_size)
        center, radius = model_new.board_size // 2, 4
        game.do_move((random.randint(center - radius, center + radius), random.randint(center - radius, center + radius)))
        new_starts = i % 2 == 0
        move_number = 1
        mcts_new = MCTS(GameClass, n_simulations, model_new, cpuct=cpuct, add_dirichlet_noise=False)
        mcts_best = MCTS(GameClass, n_simulations, model_best, cpuct=cpuct, add_dirichlet_noise=False)
        while not game.is_game_over():
            mine = (game.current_player == 1 and new_starts) or (game.current_player == 2 and not new_starts)
            pi = (mcts_new if mine else mcts_best).run(game, len(game.move_history))
            game.do_move(divmod(int(np.argmax(pi)), game.size))
            move_number += 1
            if move_number > game.size * game.size:
                break
        winner = game.get_winner()
        if winner == 0:
            draws += 1
        elif (winner == 1 and new_starts) or (winner == 2 and not new_starts):
            new_wins += 1
        mcts_new.engine.close()
        mcts_best.engine.close()
    return new_wins, new_wins / float(n_games), draws


# ------------------------------------------------------------------ data-parallel helpers
def _world():
    return (dist.get_world_size(), dist.get_rank()) if dist.is_available() and dist.is_initialized() else (1, 0)


def gather_examples(sp: SelfPlay) -> torch.Tensor:
    """The exchange step of the loop (train.py:737-742 pickles every worker's expanded rows through a pipe): every
    rank contributes the games it finished as PACKED plies (976 bytes each: stones as bits, side, z, pi), all ranks
    all-gather those, and each expands the 8 symmetries locally - 29.5x fewer bytes on NVLink than the float rows
    (symmetries are a pure function of the ply).  Returns the expanded rows float32[n, 901] of ALL ranks, rank-major."""
    from .selfplay import expand_examples
    packed = gather_rows(sp.drain_packed())
    return expand_examples(packed, sp.use_symmetries)


def gather_rows(rows: torch.Tensor) -> torch.Tensor:
    """All-gather variable-length rows [n_r, width] (example rows or packed plies) from every rank: one small
    all-gather of the counts (the only host synchronisation), one all-gather of the padded blocks into a single
    tensor, then the valid prefixes concatenated rank-major."""
    world, _ = _world()
    if world == 1:
        return rows
    into_tensor = dist.get_backend() == "nccl"            # gloo (CPU tests) only has the list form

    def all_gather(block):
        if into_tensor:
            out = torch.empty((world,) + tuple(block.shape), dtype=block.dtype, device=block.device)
            dist.all_gather_into_tensor(out, block)
            return out
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block)
        return torch.stack(parts)

    counts = all_gather(torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)).reshape(-1).tolist()
    m = max(counts)
    if m == 0:
        return rows[:0]
    if rows.shape[0] == m:
        pad = rows.contiguous()
    else:
        pad = torch.zeros((m, rows.shape[1]), dtype=rows.dtype, device=rows.device)
        pad[: rows.shape[0]] = rows
    parts = all_gather(pad)
    if all(c == m for c in counts):
        return parts.reshape(world * m, rows.shape[1])
    return torch.cat([parts[r, :c] for r, c in enumerate(counts)], dim=0)


def train_step_device(model: PyTorchModel, states, pis, zs, world: int = 1) -> torch.Tensor:
    """One ``train_batch`` step (network.py:199-235: KLDiv(batchmean) + MSE, clip 3.0, Adam) WITHOUT a host
    synchronisation: returns the device tensor [policy_loss, value_loss].  ``world`` > 1: this rank's micro-batch,
    gradients averaged over all ranks before the clip and the step, so every rank applies the identical update."""
    if getattr(model, "_ensure_trainer", None) is not None and model.net.channels in (64, 128, 256):
        # CUDA training step: one all-reduce over the flat gradient vector between backward and clip + Adam
        reduce = (lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM)) if world > 1 else None
        return model.train_batch_async(states, pis, zs, world=world, reduce_grads=reduce)
    net, dev = model.net, model.device
    to = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.asarray(a, dtype=np.float32))).to(dev, torch.float32)
    net.train()
    model.optimizer.zero_grad()
    logits, values = net(to(states))
    pl = model.policy_loss_fn(F.log_softmax(logits, dim=1), to(pis))
    vl = model.value_loss_fn(values, to(zs))
    (pl + vl).backward()
    if world > 1:
        grads = [p.grad for p in net.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= world
        o = 0
        for g in grads:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 3.0)
    model.optimizer.step()          # in-place parameter writes bump the tensor versions: the inference cache repacks on its own
    return torch.stack([pl.detach(), vl.detach()])


def train_batch_dp(model: PyTorchModel, states, pis, zs) -> dict:
    """``PyTorchModel.train_batch`` on this rank's micro-batch with the gradients averaged over all ranks
    (``train_step_device``), returning the reference's dictionary of floats (one host synchronisation)."""
    world, _ = _world()
    if world == 1:
        return model.train_batch(states, pis, zs, epochs=1)
    p, v = train_step_device(model, states, pis, zs, world).tolist()
    return {"policy_loss": p, "value_loss": v, "total_loss": p + v}


def broadcast_model(model: PyTorchModel, src: int = 0):
    world, _ = _world()
    if world == 1:
        return
    with torch.no_grad():
        for t in list(model.net.parameters()) + list(model.net.buffers()):
            dist.broadcast(t, src)
    model.invalidate()


def broadcast_optimizer(model: PyTorchModel, src: int = 0):
    """Adam's moments and step counters from ``src`` to every rank (after a replicated training phase)."""
    world, _ = _world()
    if world == 1:
        return
    with torch.no_grad():
        for p in model.net.parameters():
            st = model.optimizer.state.get(p)
            if not st:
                continue
            for k in ("exp_avg", "exp_avg_sq", "step"):
                v = st.get(k)
                if torch.is_tensor(v):
                    if v.device.type == "cpu" and dist.get_backend() == "nccl":       # torch keeps `step` on the host by default
                        t = v.to(model.device)
                        dist.broadcast(t, src)
                        v.copy_(t.cpu())
                    else:
                        dist.broadcast(v, src)


# Data-parallel training splits a batch over the ranks and BatchNorm then normalises every slice with its own
# statistics (as torch's DistributedDataParallel does without SyncBatchNorm).  The reference normalises the WHOLE batch
# (network.py:210-226, default batch_size 128): below this many positions per rank the loop keeps the reference's
# arithmetic instead - every rank trains the whole batch (same buffer, same draws) and rank 0's result is broadcast
# after the phase (the ranks differ only in the summation order of the weight gradients until then).
DP_MIN_POSITIONS_PER_RANK = 64


def sync_batchnorm_buffers(model: PyTorchModel):
    """Data-parallel training normalises every rank's micro-batch with its own statistics, so the BatchNorm
    running statistics drift apart while the parameters stay identical.  Average them over the ranks (each
    rank saw an equally sized slice of every batch) so that self-play, the arena and the saved checkpoint use
    one and the same network on every rank."""
    world, _ = _world()
    if world == 1:
        return
    with torch.no_grad():
        for name, b in model.net.named_buffers():
            if b.dtype.is_floating_point:
                dist.all_reduce(b, op=dist.ReduceOp.SUM)
                b /= world
            else:                                   # num_batches_tracked: identical on every rank already
                dist.broadcast(b, 0)
    model.invalidate()


# ------------------------------------------------------------------ the training loop (train.py:575-842)
def train_alphazero(game_name: str = "gomoku", board_size: int = 15, num_iterations: int = 5, games_per_iteration: int = 8,
                    n_simulations: int = 50, buffer_size: int = 10000, batch_size: int = 128, epochs_per_iter: int = 2,
                    temp_threshold: int = 8, eval_games: int = 12, eval_mcts_simulations: int = 200,
                    win_rate_threshold: float = 0.55, cpuct: float = 1.2, model_dir: str = "models", save_every: int = 1,
                    pretrained_model_path: Optional[str] = None, next_iteration_continuation: int = 1,
                    dirichlet_alpha: float = 0.03, dirichlet_epsilon: float = 0.25, dirichlet_n_moves: int = 30,
                    selfplay_num_workers: int = 0, selfplay_device: str = "cpu", selfplay_games_per_task: int = 1,
                    selfplay_base_seed: int = 12345, selfplay_torch_threads: int = 1, eval_num_workers: int = 0,
                    eval_device: str = "cpu", eval_games_per_task: int = 1, eval_base_seed: int = 54321,
                    eval_torch_threads: int = 1, n_res_blocks: int = 3, channels: int = 64, concurrent_games: int = 0):
    """Same keyword arguments as the reference (the multiprocessing ones are accepted and ignored:
    the process pool is replaced by device batching); ``n_res_blocks`` / ``channels`` /
    ``concurrent_games`` are engine extras.  Returns the best model."""
    world, rank = _world()
    dev = f"cuda:{torch.cuda.current_device()}"
    os.makedirs(model_dir, exist_ok=True)
    mk = lambda: PyTorchModel(board_size=board_size, device=dev, n_res_blocks=n_res_blocks, channels=channels)
    model_best, model_candidate = mk(), mk()
    if pretrained_model_path and os.path.exists(pretrained_model_path):
        model_best.load(pretrained_model_path)
    broadcast_model(model_best)
    model_candidate.net.load_state_dict(model_best.net.state_dict())
    model_candidate.optimizer.load_state_dict(model_best.optimizer.state_dict())
    buffer_path = os.path.join(model_dir, "replay_buffer_latest.pkl")
    host_buffer = load_replay_buffer(buffer_path, buffer_size)
    buffer = DeviceReplayBuffer.from_host(host_buffer, dev) if host_buffer else DeviceReplayBuffer(buffer_size, dev)
    sample_gen = torch.Generator(device=dev)
    my_games = (games_per_iteration + world - 1) // world
    G = concurrent_games or min(my_games, 2048)
    for it in range(next_iteration_continuation, next_iteration_continuation + num_iterations):
        t0 = time.time()
        if rank == 0:
            print(f"\n=== ITER {it}: self-play (games={games_per_iteration}, sims={n_simulations}) {datetime.now():%Y-%m-%d %H:%M:%S} ===")
        sp = SelfPlay(model_candidate, rule=0, n_games=G, n_sims=n_simulations, cpuct=cpuct, noise=True,
                      alpha=dirichlet_alpha, eps=dirichlet_epsilon, noise_plies=dirichlet_n_moves,
                      temp_threshold=float(temp_threshold), max_moves=board_size * board_size,
                      example_capacity=max(my_games, G) * 225, seed=selfplay_base_seed + it, game_base=rank * G,
                      node_capacity=max(4096, 8 * n_simulations), device=dev, max_games=my_games, packed_examples=True)    # soak: high-water 5.5 x sims
        # exactly my_games games are started and every one of them is played to the end (train.py:671-694):
        # slots restart only while games remain to be started, then retire
        finished = 0
        winners = {0: 0, 1: 0, 2: 0}
        while sp.games_running() > 0:
            sp.step()
            w = sp.winners[sp.done.bool()].cpu().tolist()
            for x in w:
                winners[int(x)] = winners.get(int(x), 0) + 1
            finished += len(w)
        assert finished == my_games
        st = sp.engine.stats()
        if st["dropped_trees"] or st["games_in_error"]:      # a dropped tree loses tree reuse for one move: say so, never silently
            print(f"[self-play] WARNING rank {rank}: {st['dropped_trees']} trees dropped (node slabs full), {st['games_in_error']} games in "
                  f"error (bits {st['error_bits']:#x}); raise node_capacity (high-water mark {st['max_nodes']} nodes per game)")
        rows = gather_examples(sp)
        sp.close()
        buffer.add_rows(rows)
        if rank == 0:
            print(f"[self-play] {finished * world} games, {rows.shape[0]} examples, winners {winners}, {(time.time() - t0) / 60:.2f} min")
        # ---- training (train.py:754-763): every rank draws the same batches (shared seed) and takes its slice
        n_batches = len(buffer) // batch_size
        sample_gen.manual_seed(1000003 * it + 17)      # identical buffers + identical draws on every rank
        replicated = world > 1 and batch_size // world < DP_MIN_POSITIONS_PER_RANK
        for ep in range(epochs_per_iter):
            tot = torch.zeros((), dtype=torch.float32, device=dev)      # losses stay on the device: one host read per epoch
            for _ in range(n_batches):
                states, pis, zs = buffer.sample(batch_size, generator=sample_gen)
                if world == 1 or replicated:
                    tot += train_step_device(model_candidate, states, pis, zs).sum()
                else:
                    sl = slice(rank * batch_size // world, (rank + 1) * batch_size // world)
                    tot += train_step_device(model_candidate, states[sl], pis[sl], zs[sl], world).sum()
            if rank == 0 and n_batches:
                print(f"[train] epoch {ep + 1}/{epochs_per_iter} mean loss {float(tot) / n_batches:.4f}")
        if replicated:
            broadcast_model(model_candidate)
            broadcast_optimizer(model_candidate)
        else:
            sync_batchnorm_buffers(model_candidate)
        # ---- evaluation and accept / reject (train.py:768-827), rank 0 decides
        accept = torch.zeros(1, dtype=torch.int32, device=dev)
        # every rank plays its slice of the match (collective inside: all ranks must call it)
        try:
            new_wins, win_rate, draws = evaluate_models_dp(model_candidate, model_best, game_name, n_games=eval_games,
                                                           n_simulations=eval_mcts_simulations, cpuct=cpuct)
        except Exception as e:          # single process: the reference prints and goes on with win_rate 0 (train.py:783-786)
            if world > 1:
                raise                   # a rank that skipped the collectives would hang the others
            print(f"[eval] failed: {e}")
            new_wins, win_rate, draws = 0, 0.0, 0
        if rank == 0:
            print(f"[eval] new model wins {new_wins}/{eval_games} (draws {draws}) win rate {win_rate:.3f}")
            accept[0] = int(win_rate >= win_rate_threshold)
        if world > 1:
            dist.broadcast(accept, 0)
        if int(accept.item()):
            model_best.net.load_state_dict(model_candidate.net.state_dict())
            model_best.optimizer.load_state_dict(model_candidate.optimizer.state_dict())
        else:
            model_candidate.net.load_state_dict(model_best.net.state_dict())
            model_candidate.optimizer.load_state_dict(model_best.optimizer.state_dict())
        if rank == 0 and it % save_every == 0:
            model_best.save(os.path.join(model_dir, f"snapshot_iter{it}_{datetime.now():%Y%m%d_%H%M%S}.pt"))
            save_replay_buffer(buffer.to_host(), buffer_path)
        if rank == 0:
            print(f"=== ITER {it} done in {(time.time() - t0) / 60:.2f} min ===")
    return model_best
