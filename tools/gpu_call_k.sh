#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
AZG_WGRAD_CLUSTER=1 timeout 300 python -m pytest tests/test_train_gpu.py -q -s -k "exact" > $O/k_exact_cluster.log 2>&1; echo "rc=$?" >> $O/k_exact_cluster.log
AZG_WGRAD_CLUSTER=1 timeout 600 python -m pytest tests/test_train_gpu.py -q > $O/k_tests_cluster.log 2>&1; echo "rc=$?" >> $O/k_tests_cluster.log
AZG_WGRAD_CLUSTER=0 timeout 600 python -m pytest tests/test_train_gpu.py -q > $O/k_tests_plain.log 2>&1; echo "rc=$?" >> $O/k_tests_plain.log
for c in 0 1; do
  AZG_WGRAD_CLUSTER=$c timeout 300 python tools/train_step_bench.py --batches 128,512,1024 --steps 20 --skip-autograd > $O/k_bench_c$c.jsonl 2>> $O/k_bench.err
done
AZG_WGRAD_CLUSTER=1 AZG_TRAIN_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/k_train_launches.csv python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/k_ncu1.log 2>&1
grep -E "^C=|passed|failed" $O/k_exact_cluster.log; tail -n 2 $O/k_tests_cluster.log $O/k_tests_plain.log; cat $O/k_bench_c0.jsonl $O/k_bench_c1.jsonl | cut -c1-200
