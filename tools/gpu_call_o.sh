#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests/test_search_gpu.py tests/test_dropin_gpu.py tests/test_reference_pinned_gpu.py tests/test_fullsize_gpu.py tests/test_selfplay_gpu.py tests/test_net_gpu.py -q > $O/o_tests.log 2>&1; echo "rc=$?" >> $O/o_tests.log
rm -f $O/o_latency_cc_*.jsonl
for v in 0 1 0 1; do
  AZG_CHILD_CODES=$v timeout 300 python tools/player_latency.py --moves 4 >> $O/o_latency_cc_$v.jsonl 2>> $O/o_latency.err
done
for v in 0 1 0 1; do
  AZG_CHILD_CODES=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/o_bench_cc_${v}.json 2>> $O/o_bench.err
  python -c "
import json;d=json.loads(open('$O/o_bench_cc_${v}.json').read().strip().splitlines()[-1]);print('cc=$v', d['value'], d['roofline']['trunk_share_of_step'], d['search']['engine_gb'])"
done
tail -n 3 $O/o_tests.log; cat $O/o_latency_cc_0.jsonl $O/o_latency_cc_1.jsonl | cut -c1-200
