#!/bin/bash
# Everything the round's evidence rests on, in one gpurun call (single GPU):
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_validate.sh'
# parity suite, smoke, the bench line of every BASELINE config, the reference arm, the training-step bench.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $O/v_pytest.log 2>&1; echo "pytest rc=$?" >> $O/v_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/v_smoke.log 2>&1; echo "smoke rc=$?" >> $O/v_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/v_bench.json 2> $O/v_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --rule pente --no-cpu-baseline > $O/v_bench_pente.json 2>> $O/v_bench.err
timeout 300 python bench.py --steps 5 --warmup 3 --blocks 3 --channels 64 --no-cpu-baseline > $O/v_bench_3x64.json 2>> $O/v_bench.err
timeout 400 python bench.py --steps 3 --warmup 3 --blocks 10 --channels 256 --no-cpu-baseline > $O/v_bench_10x256.json 2>> $O/v_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $O/v_bench_reference.json 2>> $O/v_bench.err
timeout 600 python tools/train_step_bench.py --batches 128,256,512,1024,2048 --steps 20 > $O/v_train_bench.jsonl 2> $O/v_train_bench.err
timeout 300 python tools/train_step_bench.py --blocks 3 --channels 64 --batches 128,1024 --steps 20 >> $O/v_train_bench.jsonl 2>> $O/v_train_bench.err
tail -n 4 $O/v_pytest.log; tail -n 2 $O/v_smoke.log; cut -c1-220 $O/v_train_bench.jsonl
for f in v_bench v_bench_pente v_bench_3x64 v_bench_10x256 v_bench_reference; do python - <<PY
import json
d = json.loads(open("$O/$f.json").read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print("$f", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", r.get("frac"), (d.get("cpu_baseline") or {}).get("kind"), (d.get("search") or {}).get("dropped_trees"))
PY
done
