#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_selfplay_gpu.py -q > $O/e_tests.log 2>&1; echo "rc=$?" >> $O/e_tests.log
timeout 900 python tools/train_step_bench.py --batches 128,256,512,1024 --steps 20 --skip-autograd > $O/e_train_bench.jsonl 2> $O/e_train_bench.err
AZG_TRAIN_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/e_train_launches.csv python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/e_ncu1.log 2>&1
tail -4 $O/e_tests.log; cat $O/e_train_bench.jsonl
