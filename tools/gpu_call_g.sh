#!/bin/bash
# round 2, GPU call G: whole GPU suite, smoke, bench + launch list + full ncu captures of the trunk kernel (6x128 and 3x64)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $O/g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/g_smoke.log 2>&1; echo "smoke rc=$?" >> $O/g_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/g_bench.json 2> $O/g_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --blocks 3 --channels 64 --no-cpu-baseline > $O/g_bench_3x64.json 2>> $O/g_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/g_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_pair -s 30 -c 2 -o $O/conv3x3_r02 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/g_ncu2.log 2>&1
ncu -i $O/conv3x3_r02.ncu-rep --page raw --csv > $O/conv3x3_r02_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:conv3x3_pair -s 10 -c 2 -o $O/conv3x3_r02_3x64 python bench.py --steps 1 --warmup 1 --blocks 3 --channels 64 --no-cpu-baseline > $O/g_ncu3.log 2>&1
ncu -i $O/conv3x3_r02_3x64.ncu-rep --page raw --csv > $O/conv3x3_r02_3x64_raw.csv 2>/dev/null
tail -4 $O/g_pytest.log; tail -2 $O/g_smoke.log
