#!/bin/bash
# round 2, GPU call B: training step parity, new self-play tests, 3x64 with two epilogue groups
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x -s > $O/b_train.log 2>&1; echo "rc=$?" >> $O/b_train.log
AZG_WGRAD_DESC=1 timeout 300 python -m pytest tests/test_train_gpu.py -q -x -s -k "forward_and_gradients" > $O/b_train_v1.log 2>&1; echo "rc=$?" >> $O/b_train_v1.log
timeout 600 python -m pytest tests/test_selfplay_gpu.py tests/test_net_gpu.py tests/test_dropin_gpu.py -q -x > $O/b_other.log 2>&1; echo "rc=$?" >> $O/b_other.log
timeout 300 python bench.py --steps 6 --warmup 3 --blocks 3 --channels 64 --no-cpu-baseline > $O/b_bench_3x64.json 2> $O/b_bench_3x64.err
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 --sims 64 > $O/b_ref.json 2> $O/b_ref.err
tail -3 $O/b_train.log $O/b_other.log
