#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q > $O/i_tests.log 2>&1; echo "rc=$?" >> $O/i_tests.log
for rb in 148 160 296; do
  AZG_TRAIN_RED_BLOCKS=$rb timeout 300 python tools/train_step_bench.py --batches 512,1024 --steps 20 --skip-autograd > $O/i_bench_rb$rb.jsonl 2>> $O/i_bench.err
done
AZG_TRAIN_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_bwd_reduce|bn_bwd_apply|bn_apply_kernel" -s 20 -c 3 -o $O/i_bn python tools/train_step_bench.py --batches 512 --steps 2 --skip-autograd > $O/i_ncu.log 2>&1
ncu -i $O/i_bn.ncu-rep --page raw --csv > $O/i_bn_raw.csv 2>/dev/null
ncu -i $O/i_bn.ncu-rep --page details --csv > $O/i_bn_details.csv 2>/dev/null
tail -n 3 $O/i_tests.log; cat $O/i_bench_rb*.jsonl | cut -c1-200
