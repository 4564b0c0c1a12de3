#!/bin/bash
# round 2, GPU call A: parity tests, bench lines for every BASELINE config, graph vs default, soak
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --durations=15 > $O/a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/a_pytest.log
timeout 300 python bench.py --steps 6 --warmup 3 > $O/a_bench_default.json 2> $O/a_bench_default.err
for i in 1 2; do
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --graph > $O/a_bench_graph_$i.json 2>> $O/a_bench_graph.err
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/a_bench_host_$i.json 2>> $O/a_bench_host.err
done
timeout 300 python bench.py --steps 6 --warmup 3 --rule pente --no-cpu-baseline > $O/a_bench_pente.json 2> $O/a_bench_pente.err
timeout 400 python bench.py --steps 3 --warmup 3 --blocks 10 --channels 256 --no-cpu-baseline > $O/a_bench_10x256.json 2> $O/a_bench_10x256.err
timeout 300 python bench.py --steps 6 --warmup 3 --blocks 3 --channels 64 --no-cpu-baseline > $O/a_bench_3x64.json 2> $O/a_bench_3x64.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/a_bench_reference.json 2> $O/a_bench_reference.err
timeout 600 python bench.py --soak 150 > $O/a_soak.json 2> $O/a_soak.err; echo "soak rc=$?" >> $O/a_soak.err
tail -5 $O/a_pytest.log
