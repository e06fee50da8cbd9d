#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), the GPU tests, both bench arms
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2; fi
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_ref.json 2> gpurun_out/drv_ref.err
SECONDS=0; python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/drv_bench.json 2> gpurun_out/drv_bench.err; echo "bench wall ${SECONDS}s"
python - <<PY
import json
r=json.loads(open("gpurun_out/drv_ref.json").read().strip().splitlines()[-1]); d=json.loads(open("gpurun_out/drv_bench.json").read().strip().splitlines()[-1])
print("ref", round(r["value"],2), r["cpu_baseline"]["kind"], "| ours", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ratio e2e", round(d["e2e"]["value"]/r["value"],1), "same_config", r["config"]==d["config"])
print("clocks", d["clocks"], "launches", d["gpu_launches"], "frac", round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"])
for k,v in d["also"].items(): print("  ", k, round(v["value"],1) if isinstance(v,dict) else v)
PY
