#!/usr/bin/env python
"""Headline benchmark: batched self-play MCTS simulations per second (BASELINE.json metric).

Workload (BASELINE.json configs[1]): Gomoku 15x15, 2048 concurrent self-play games per GPU,
800 simulations per move, random-init 6x128 ResNet (torch.manual_seed(0)), cpuct 1.0, leaf queue
32, Dirichlet alpha 0.05 / eps 0.15 on the first 10 plies, temperature threshold 10.  A "step" is
one ply of every game: 2048 x 800 simulations, i.e. 26+ rounds of FILL -> leaf evaluation ->
COMMIT.  Synthetic data: games start from seeded random legal positions of 0..39 plies so that
the batch is a steady-state mix of game phases; weights are random-init.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA engine)
    python bench.py --impl reference ...                      # the reference's own CPU path (oracle/_ref), all host cores
    python bench.py --soak 150                                # tree-memory soak: 150 plies with respawn, fails on a dropped tree

Prints ONE JSON line (see the task contract): value = device-resident throughput, e2e = the same
metric through host buffers (boards up, pi/actions/boards down every step), roofline = the
conv3x3 tcgen05 kernel against the measured bf16 peak, cpu_baseline = the reference's CPU path
(W worker processes, a bounded sample of the same workload) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "self-play MCTS sims/sec (Gomoku 15x15, 6x128 ResNet)"
UNIT = "sims/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--games", type=int, default=2048)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--blocks", type=int, default=6)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--rule", default="gomoku", choices=["gomoku", "pente"])
    ap.add_argument("--node-capacity", type=int, default=0,
                    help="nodes per game slab; default 16 384 for Gomoku, 24 576 for Pente (its garbage collector can only use a "
                         "capture-count bound, so about 16 plies of trees stay alive: soak high-water marks 4 432 and 19 216 nodes)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", action="store_true",
                    help="replay each ply as one CUDA graph (fixed round count, no host synchronisation inside the ply); "
                         "the per-launch trunk timing is then unavailable (events are not recorded inside the graph)")
    ap.add_argument("--soak", type=int, default=0, metavar="PLIES",
                    help="instead of the benchmark: play PLIES plies of the configured workload from the empty board with respawn, "
                         "then print {max_nodes_per_game, dropped_trees, games_finished, ...} and exit 1 if any tree was dropped")
    ap.add_argument("--no-train-probe", action="store_true", help="skip the training-step probe printed next to the metric (N = 1 only)")
    ap.add_argument("--pipeline", type=int, default=1, choices=[1, 2],
                    help="game groups per GPU: 2 overlaps one group's tree walk with the other group's leaf evaluation "
                         "(+1.6 %% sims/s when first measured, -0.6 %% with the final kernels; the per-launch CUDA-event timing of the trunk kernel, and with it the "
                         "roofline block, is only meaningful with 1 because the two groups' event brackets overlap)")
    return ap.parse_args()


def workload(args):
    return {"workload": f"{args.rule} 15x15 batched self-play, {args.games} concurrent games/GPU x {args.sims} sims/move, "
                        f"{args.blocks}x{args.channels} ResNet, leaf queue 32",
            "games_per_gpu": args.games, "sims_per_move": args.sims, "net": f"{args.blocks}x{args.channels}",
            "rule": args.rule, "cache": "working set (tree slabs + activations, >10 GB) exceeds the 126 MB L2",
            "parallelism": f"independent games per GPU x{args.gpus}", "game_groups_per_gpu": args.pipeline}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [x for x in sm if mx and x > 0.3 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU legs
# The reference's own CPU path, in the reference's own parallel scheme (train.py:695-742: spawned worker processes,
# one torch thread each, a private model per process).  Each worker owns ONE self-play game and advances it by one ply
# per step - reference ``MCTS.run`` of --sims simulations with tree reuse, temperature sampling, ``do_move``, symmetry
# expansion - through the reference's unmodified ``play_game_and_collect(max_moves=1)``; finished games restart from
# the empty board with a cleared tree.  That is exactly one "step" of this repository's own arm (one ply of every
# game), on W games instead of 2048 per GPU: a bounded sample of the same workload.
# The code executed is /root/reference byte-compiled into oracle/_ref (oracle/build_ref.py, kind "reference"); if that
# directory is missing the oracle port runs instead (kind "port").
def _ref_worker(conn, idx, blocks, channels, rule, sims, seed):
    import random
    import torch
    torch.set_num_threads(1)
    from oracle import build_ref
    kind = "reference" if build_ref.activate() else "port"
    np.random.seed(seed)
    random.seed(seed)
    rng = np.random.default_rng(seed)
    thr = 10.0

    class Counting:
        def __init__(self, inner):
            self.inner, self.rows = inner, 0

        def predict(self, X):
            self.rows += len(X)
            return self.inner.predict(X)

    if kind == "reference":
        from games.gomoku import Gomoku
        from games.pente import Pente
        from mcts.new_mcts_alpha import MCTS
        from network import PyTorchModel
        import train as ref_train
        torch.manual_seed(0)
        model = Counting(PyTorchModel(board_size=15, n_res_blocks=blocks, channels=channels, device="cpu"))
        cls = Pente if rule else Gomoku

        def new_game(plies):
            g = cls(15)
            for _ in range(plies):                      # seeded random legal opening, as bench.py's scatter_start
                e = np.flatnonzero(g.board.reshape(-1) == 0)
                g2 = g.clone()
                a = int(e[int(rng.integers(0, len(e)))])
                g2.do_move(divmod(a, 15))
                if g2.is_game_over():
                    break
                g = g2
            m = MCTS(game_class=cls, n_simulations=sims, nn_model=model, cpuct=1.0, dirichlet_alpha=0.05, epsilon=0.15,
                     apply_dirichlet_n_first_moves=10, add_dirichlet_noise=True)
            return g, m

        def one_ply(g, m):
            temp_fn = lambda _mn: max(0.0, 1.0 - len(g.move_history) / thr)
            ref_train.play_game_and_collect(m, g, temp_fn, max_moves=1, use_symmetries=True)
            return g.is_game_over()
    else:
        from oracle import net as onet, rules as orules, selfplay as oselfplay
        from oracle.search import Search
        import alphazero_gomoku_b200.network as mynet
        torch.manual_seed(0)
        model = Counting(onet.CpuModel(mynet.AlphaZeroNet(n_res_blocks=blocks, channels=channels).state_dict()))

        def new_game(plies):
            g = orules.Position(rule)
            for _ in range(plies):
                e = np.flatnonzero(g.cells == 0)
                g2 = g.copy()
                orules.play(g2, int(e[int(rng.integers(0, len(e)))]))
                if orules.game_over(g2):
                    break
                g = g2
            return g, Search(rule, sims, model, cpuct=1.0, queue_len=32, alpha=0.05, eps=0.15, noise_plies=10, noise=True)

        def one_ply(g, m):
            oselfplay.play_one(m, g, lambda _p: max(0.0, 1.0 - g.plies / thr), max_plies=1, expand=True)
            return orules.game_over(g)

    game, mcts = new_game(idx % 40)
    conn.send(("ready", kind))
    while True:
        cmd = conn.recv()
        if cmd == "quit":
            break
        rows0 = model.rows
        t0 = time.perf_counter()
        over = one_ply(game, mcts)
        dt = time.perf_counter() - t0
        if over:
            game, mcts = new_game(0)
        conn.send((sims, model.rows - rows0, dt))


class ReferencePool:
    def __init__(self, workers, args):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        rule = 1 if args.rule == "pente" else 0
        self.procs, self.conns = [], []
        for i in range(workers):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_ref_worker, args=(b, i, args.blocks, args.channels, rule, args.sims, 12345 + i), daemon=True)
            p.start()
            self.procs.append(p)
            self.conns.append(a)
        self.kind = [c.recv() for c in self.conns][0][1]
        self.workers = workers

    def step(self):
        """One ply of every worker's game.  -> (simulations, leaf evaluations)."""
        for c in self.conns:
            c.send("step")
        out = [c.recv() for c in self.conns]
        return sum(o[0] for o in out), sum(o[1] for o in out)

    def close(self):
        for c in self.conns:
            c.send("quit")
        for p in self.procs:
            p.join(timeout=10)


def host_workers():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 2)
    return max(1, min(n - 1, 64))


def reference_workload(args, workers, kind):
    w = workload(args)
    w["workload"] = (f"{args.rule} 15x15 self-play on the reference's CPU path ({kind}): {workers} worker processes x 1 torch thread, one game "
                     f"each, one ply per step = reference MCTS.run of {args.sims} sims/move with tree reuse via play_game_and_collect(max_moves=1), "
                     f"{args.blocks}x{args.channels} ResNet fp32, leaf queue 32; a bounded sample ({workers} games) of the "
                     f"{args.games}-games/GPU workload of the CUDA arm")
    w["games_per_gpu"] = workers
    w["parallelism"] = f"{workers} CPU worker processes (train.py:695-742 scheme)"
    w["cache"] = "n/a (CPU)"
    return w


def cpu_baseline(args, seconds):
    """The same harness for ~`seconds`: as many one-ply steps of all workers as fit (at least one)."""
    workers = host_workers()
    pool = ReferencePool(workers, args)
    sims = evals = steps = 0
    t0 = time.perf_counter()
    while steps == 0 or time.perf_counter() - t0 < seconds:
        s, e = pool.step()
        sims, evals, steps = sims + s, evals + e, steps + 1
    dt = time.perf_counter() - t0
    pool.close()
    src = "the unmodified reference (byte-compiled into oracle/_ref)" if pool.kind == "reference" else "oracle port of mcts/new_mcts_alpha.py + network.py"
    return {"value": sims / dt, "unit": UNIT, "cores": workers, "kind": pool.kind,
            "sample": f"{src}: {workers} worker processes x 1 torch thread, one self-play game each from a seeded position of 0..39 plies, "
                      f"{steps} plies per game x {args.sims} sims/move (play_game_and_collect(max_moves=1), tree reuse, root noise), "
                      f"{evals} leaf evals, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: see the comment above _ref_worker.  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = host_workers()
    pool = ReferencePool(workers, args)
    for _ in range(args.warmup):
        pool.step()
    t0 = time.perf_counter()
    total = evals = 0
    for _ in range(args.steps):
        s, e = pool.step()
        total, evals = total + s, evals + e
    dt = time.perf_counter() - t0
    pool.close()
    value = total / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": reference_workload(args, workers, pool.kind),
            "leaf_evals_per_s": evals / dt, "evals_per_sim": evals / max(total, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": pool.kind,
                             "sample": f"{workers} processes x {args.steps} plies x {args.sims} sims/move, {evals} leaf evals, {dt:.1f} s"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU engine
def scatter_start(sp, rng_seed):
    """Advance game g by (g mod 40) uniformly random legal plies so the batch mixes game phases."""
    import torch
    eng = sp.engine
    G = sp.G
    gen = torch.Generator(device=sp.device).manual_seed(rng_seed)
    pos = eng.roots()
    target = torch.arange(G, device=sp.device, dtype=torch.int32) % 40
    for t in range(39):
        legal = eng.rules.legal(pos)
        a = torch.multinomial(legal, 1, generator=gen).squeeze(1).to(torch.int32)
        a = torch.where(target > t, a, torch.full_like(a, -1))
        before = pos.clone()
        st = eng.rules.play(pos, a)
        over = (st & 4) != 0
        pos = torch.where(over[:, None], before, pos)          # never start from a finished game
    eng.set_roots(pos, clear_tree=True)
    return pos


def run_soak(args, sp, units, rank, world):
    """Tree-memory soak: full games from the empty board with respawn on termination, the garbage
    collector of ``azg_search_advance`` working at game length.  A dropped tree is a silent departure from the
    reference's results, so any drop fails the run."""
    import torch
    t0 = time.perf_counter()
    finished = 0
    for ply in range(args.soak):
        sp.step()
        finished += int(sum(int(u.done.sum().item()) for u in units))
        for u in units:
            u.cursor.zero_()
    torch.cuda.synchronize()
    per = [u.engine.stats() for u in units]
    line = {"soak_plies": args.soak, "config": workload(args), "rank": rank, "games_finished": finished,
            "max_nodes_per_game": max(x["max_nodes"] for x in per), "node_capacity": args.node_capacity,
            "dropped_trees": sum(x["dropped_trees"] for x in per), "games_in_error": sum(x["games_in_error"] for x in per),
            "sims": sum(x["sims"] for x in per), "seconds": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)
    if line["dropped_trees"] or line["games_in_error"]:
        sys.exit(1)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import PipelinedSelfPlay, SelfPlay, trunk_flops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rule = 1 if args.rule == "pente" else 0

    torch.manual_seed(0)
    model = PyTorchModel(board_size=15, n_res_blocks=args.blocks, channels=args.channels, device=str(dev))
    kw = dict(rule=rule, n_sims=args.sims, cpuct=1.0, queue_len=32, node_capacity=args.node_capacity, noise=True, alpha=0.05,
              eps=0.15, noise_plies=10, temp_threshold=10.0, example_capacity=1 << 18, seed=12345)
    if args.pipeline == 2:
        sp = PipelinedSelfPlay(model, n_games=args.games, game_base=rank * args.games, device=str(dev), **kw)
        units = sp.halves
    else:
        sp = SelfPlay(model, n_games=args.games, game_base=rank * args.games, device=str(dev), **kw)
        units = [sp]
    if args.soak:
        return run_soak(args, sp, units, rank, world)
    for i, u in enumerate(units):
        scatter_start(u, 777 + 2 * rank + i)
        if args.graph:
            u.enable_graph()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def totals():
        t = {k: sum(getattr(u, k) for u in units) for k in ("total_sims", "total_launches", "total_rounds")}
        t["total_evals"] = sum(u.engine.stats()["evals"] for u in units)      # device-side count (also valid in graph mode)
        return t

    def reset_cursors():
        for u in units:
            u.cursor.zero_()

    for _ in range(args.warmup):
        sp.step()
        reset_cursors()

    # ---- timed region A: device-resident self-play
    clocks = ClockSampler(local)
    t_before = totals()
    for u in units:
        u.net.profile(True)
        u.net.profile_read()
        u.net.profile_counters()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sp.step()
        reset_cursors()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    trunk_ms, conv_launches, pipe = 0.0, 0, {}
    for u in units:
        t_ms, n_l = u.net.profile_read()
        trunk_ms += t_ms
        conv_launches += n_l
        for k, v in u.net.profile_counters().items():
            if isinstance(v, dict):
                d = pipe.setdefault(k, {})
                for kk, vv in v.items():
                    d[kk] = d.get(kk, 0) + vv
            else:
                pipe[k] = pipe.get(k, 0) + v
        u.net.profile(False)
    t_after = totals()
    sims, evals = t_after["total_sims"] - t_before["total_sims"], t_after["total_evals"] - t_before["total_evals"]
    launches, rounds = t_after["total_launches"] - t_before["total_launches"], t_after["total_rounds"] - t_before["total_rounds"]
    clk = clocks.stop()
    per_unit = [u.engine.stats() for u in units]
    stats = {k: sum(x[k] for x in per_unit) for k in per_unit[0]}
    stats["max_nodes"] = max(x["max_nodes"] for x in per_unit)
    stats["error_bits"] = 0
    for x in per_unit:
        stats["error_bits"] |= x["error_bits"]

    # ---- timed region B: the same step driven through HOST buffers
    host = []
    for u in units:
        g = u.G
        host.append(dict(boards=torch.empty((g, 225), dtype=torch.int8).pin_memory(),
                         meta=torch.empty((g, 5), dtype=torch.int32).pin_memory(),          # player, last, caps0, caps1, plies
                         pi=torch.empty((g, 225), dtype=torch.float32).pin_memory(),
                         act=torch.empty((g, 2), dtype=torch.int32).pin_memory()))          # action, status

    def download(u, hb):
        boards, players, lasts, caps, plies = u.engine.rules.unpack(u.engine.roots())
        hb["boards"].copy_(boards, non_blocking=True)
        hb["meta"].copy_(torch.stack([players, lasts, caps[:, 0], caps[:, 1], plies], dim=1), non_blocking=True)

    for u, hb in zip(units, host):
        download(u, hb)
    torch.cuda.synchronize()
    h2d = sum(hb["boards"].numel() + hb["meta"].numel() * 4 for hb in host)
    d2h = sum(hb["pi"].numel() * 4 + hb["act"].numel() * 4 for hb in host) + h2d

    def e2e_step():
        for u, hb in zip(units, host):
            d_boards = hb["boards"].to(dev, non_blocking=True)
            d_meta = hb["meta"].to(dev, non_blocking=True)
            pos = u.engine.rules.pack(d_boards, d_meta[:, 0].contiguous(), d_meta[:, 1].contiguous(), d_meta[:, 2:4].contiguous(),
                                      d_meta[:, 4].contiguous())
            u.engine.set_roots(pos, clear_tree=False)
        status = sp.step()
        off = 0
        for u, hb in zip(units, host):
            hb["pi"].copy_(u.last_pi, non_blocking=True)
            hb["act"].copy_(torch.stack([u.actions, status[off:off + u.G]], dim=1), non_blocking=True)
            off += u.G
            download(u, hb)
        torch.cuda.synchronize()
        reset_cursors()

    e2e_step()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    t = torch.tensor([ms, ms_e2e, float(sims), float(evals)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, ms_e2e = float(mx[0]), float(mx[1])
        tot_sims, tot_evals = float(sm[2]), float(sm[3])
    else:
        tot_sims, tot_evals = float(sims), float(evals)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else "fallback 1.4 PF sustained (B200_PROFILING.md)"
        flops = float(evals) * trunk_flops(args.channels) * 2 * args.blocks
        achieved = flops / (trunk_ms * 1e-3) / 1e12 if trunk_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": tot_sims / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload(args),
            "leaf_evals_per_s": tot_evals / (ms * 1e-3), "evals_per_sim": tot_evals / max(tot_sims, 1.0),
            "node_visits_per_sim": stats["visits"] / max(stats["sims"], 1),
            "e2e": {"value": tot_sims / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "conv3x3_pair_kernel", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None,
                         # DRAM bytes per launch from the committed ncu --set full capture (65 536 positions per launch, the
                         # same size as the bench's rounds): mean of a layer without (8.56 GB) and with residual (13.01 GB)
                         "traffic": 10.787e9 if (args.channels == 128 and args.games * 32 == 65536) else None,
                         "traffic_source": "profiles/conv3x3_r02_ncu_full.csv: dram__bytes_read.sum + dram__bytes_write.sum, mean of a layer without (4.33 + 4.23 GB) and with residual (8.77 + 4.24 GB)",
                         "algorithmic_bytes_per_launch": int((evals / max(rounds, 1)) * 225 * args.channels * 2 * 2.5),
                         "peak_source": peak_src,
                         "launches": int(conv_launches), "avg_launch_ms": trunk_ms / max(conv_launches, 1),
                         "trunk_share_of_step": trunk_ms / ms if ms else None,
                         "algorithmic_flops_per_position_per_launch": trunk_flops(args.channels),
                         "pipeline_cycles_per_board": {k: round(v / max(pipe["boards"], 1), 1) for k, v in pipe.items()
                                                       if k != "boards" and not isinstance(v, dict)},
                         "pipeline_cycles_per_board_by_layer_type": {
                             t: {k: round(v / max(pipe[t]["boards"], 1), 1) for k, v in pipe[t].items() if k != "boards"}
                             for t in ("plain", "residual") if t in pipe}},
            "clocks": clk,
            "search": {"rounds": int(rounds), "game_groups": len(units), "games_in_error": stats["games_in_error"],
                       "error_bits": stats["error_bits"], "max_nodes_per_game": stats["max_nodes"],
                       "dropped_trees": stats["dropped_trees"], "engine_gb": sum(u.engine.memory_bytes for u in units) / 1e9,
                       "net_gb": sum(u.net.memory_bytes for u in units) / 1e9},
        }
        if world == 1 and not args.no_train_probe:
            try:
                line["train_step"] = train_probe(args, dev)
            except Exception as e:                      # a probe: never costs the bench line
                line["train_step"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def train_probe(args, dev):
    """Not part of the metric: the next row of SURVEY 8(f), PyTorchModel.train_batch (network.py:199-235), timed on the same
    network after the self-play measurement - the tensor-core training step replayed from CUDA graphs, batch resident in HBM,
    CUDA events over 20 steps after 5 warm-up steps.  Useful FLOPs = 3 x the forward trunk FLOPs on the 225 real pixels."""
    import torch
    from alphazero_gomoku_b200.network import PyTorchModel
    from alphazero_gomoku_b200.selfplay import trunk_flops
    torch.manual_seed(1)
    model = PyTorchModel(n_res_blocks=args.blocks, channels=args.channels, device=str(dev))
    g = torch.Generator(device=dev).manual_seed(1)
    flops_pos = 3 * trunk_flops(args.channels) * 2 * args.blocks
    rows = []
    for B in (128, 1024):
        x = (torch.rand((B, 3, 15, 15), device=dev, generator=g) < 0.15).float()
        x[:, 1] *= (1 - x[:, 0])
        x[:, 2] = 1.0
        pi = torch.softmax(torch.randn((B, 225), device=dev, generator=g), dim=1)
        z = torch.randint(-1, 2, (B, 1), device=dev, generator=g).float()
        for _ in range(5):
            model.train_batch_async(x, pi, z)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            losses = model.train_batch_async(x, pi, z)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / 20
        rows.append({"batch": B, "ms_per_step": round(ms, 4), "positions_per_s": round(B / (ms * 1e-3), 1),
                     "useful_tflops": round(B * flops_pos / (ms * 1e-3) / 1e12, 1), "last_losses": [round(float(v), 4) for v in losses.tolist()]})
    return {"what": "PyTorchModel.train_batch (forward, loss, backward, clip 3.0, Adam) on tensor cores, CUDA-graph replay, batch resident in HBM; "
                    "a probe next to the metric, not part of it", "net": f"{args.blocks}x{args.channels}", "steps": 20, "warmup": 5, "results": rows}


def main():
    args = parse()
    if not args.node_capacity:
        args.node_capacity = 24576 if args.rule == "pente" else 16384
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
