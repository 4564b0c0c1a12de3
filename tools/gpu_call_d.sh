#!/bin/bash
# round 2, GPU call D: training-step throughput + profiles
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q > $O/d_train_tests.log 2>&1; echo "rc=$?" >> $O/d_train_tests.log
timeout 900 python tools/train_step_bench.py --batches 128,512,1024 --steps 20 > $O/d_train_bench.jsonl 2> $O/d_train_bench.err
timeout 600 python tools/train_step_bench.py --blocks 3 --channels 64 --batches 128,1024 --steps 20 > $O/d_train_bench_3x64.jsonl 2>> $O/d_train_bench.err
AZG_TRAIN_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/d_train_launches.csv python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/d_ncu1.log 2>&1
AZG_TRAIN_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wgrad3x3|conv3x3_pair" -s 60 -c 6 -o $O/d_train_tc python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/d_ncu2.log 2>&1
ncu -i $O/d_train_tc.ncu-rep --page raw --csv > $O/d_train_tc_raw.csv 2>/dev/null
tail -3 $O/d_train_tests.log; cat $O/d_train_bench.jsonl $O/d_train_bench_3x64.jsonl
