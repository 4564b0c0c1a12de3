#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_dropin_gpu.py tests/test_reference_pinned_gpu.py tests/test_fullsize_gpu.py tests/test_selfplay_gpu.py -q > $O/l_tests.log 2>&1; echo "rc=$?" >> $O/l_tests.log
for v in 0 1 0 1; do
  AZG_FILL_L1=$v timeout 300 python tools/player_latency.py --moves 4 >> $O/l_latency_l1_$v.jsonl 2>> $O/l_latency.err
done
for v in 0 1; do
  AZG_FILL_L1=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/l_bench_l1_$v.json 2>> $O/l_bench.err
done
tail -n 3 $O/l_tests.log; cat $O/l_latency_l1_0.jsonl $O/l_latency_l1_1.jsonl | cut -c1-300; for v in 0 1; do python -c "
import json;d=json.loads(open('$O/l_bench_l1_$v.json').read().strip().splitlines()[-1]);print($v, d['value'], d['roofline']['trunk_share_of_step'])"; done
