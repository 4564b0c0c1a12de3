#!/bin/bash
# 2-GPU: data-parallel train loop (NCCL all-gather of packed plies, flat gradient all-reduce between the two step graphs)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/train_loop_bench.py --games 512 --sims 200 --plies 64 --train-steps 40 --batch-per-gpu 128 > $O/f_loop_n${N}_b128.json 2> $O/f_loop_n${N}_b128.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/train_loop_bench.py --games 512 --sims 200 --plies 64 --train-steps 40 --batch-per-gpu 1024 > $O/f_loop_n${N}_b1024.json 2> $O/f_loop_n${N}_b1024.err
tail -2 $O/f_loop_n${N}_b128.json $O/f_loop_n${N}_b1024.json; tail -5 $O/f_loop_n${N}_b128.err
