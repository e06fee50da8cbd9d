#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_backward.py -x -q -s > gpurun_out/r2b_pytest_bwd.log 2>&1; tail -12 gpurun_out/r2b_pytest_bwd.log
python bench.py --workload cfg4 --batch 8 --steps 5 --warmup 3 --dump-ops gpurun_out/r2b_train_ops_b8.txt > gpurun_out/r2b_train_b8.json 2> gpurun_out/r2b_train_b8.err; tail -c 400 gpurun_out/r2b_train_b8.err
python bench.py --workload cfg4 --steps 5 --warmup 3 --dump-ops gpurun_out/r2b_train_ops_b32.txt > gpurun_out/r2b_train_b32.json 2> gpurun_out/r2b_train_b32.err; tail -c 400 gpurun_out/r2b_train_b32.err
cat gpurun_out/r2b_train_ops_b32.txt | sort -k2 -n -r | head -30
python - <<PY
import json
for f in ("gpurun_out/r2b_train_b8.json","gpurun_out/r2b_train_b32.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline_parts"))
    except Exception as e: print(f, "ERR", e)
PY
