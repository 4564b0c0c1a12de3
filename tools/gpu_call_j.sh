#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q > $O/j_tests.log 2>&1; echo "rc=$?" >> $O/j_tests.log
timeout 300 python tools/train_step_bench.py --batches 128,256,512,1024 --steps 20 --skip-autograd > $O/j_bench.jsonl 2>> $O/j_bench.err
timeout 300 python tools/train_step_bench.py --blocks 3 --channels 64 --batches 128,1024 --steps 20 --skip-autograd > $O/j_bench_3x64.jsonl 2>> $O/j_bench.err
AZG_TRAIN_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/j_train_launches.csv python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/j_ncu1.log 2>&1
tail -n 3 $O/j_tests.log; cat $O/j_bench.jsonl $O/j_bench_3x64.jsonl | cut -c1-210
