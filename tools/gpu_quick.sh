#!/bin/bash
# usage: bash tools/gpu_quick.sh <tag> : full GPU suite + sampling bench (with cfg3/cfg5 records) + training bench
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu --maxfail=8 -q > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --dump-ops gpurun_out/${tag}_ops_cfg2.txt > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -c 300 gpurun_out/${tag}_bench.err
timeout 200 python bench.py --workload cfg4 --steps 10 --warmup 3 --dump-ops gpurun_out/${tag}_train_ops.txt > gpurun_out/${tag}_train.json 2> gpurun_out/${tag}_train.err; tail -c 300 gpurun_out/${tag}_train.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("cfg2", round(d["value"],1), "us", round(d["ms_per_step"]*1e3,1), "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3))
    for k,v in d["also"].items():
        if isinstance(v,dict): print("  ",k, round(v["value"],1), "frac", v.get("roofline",{}).get("frac"))
    print("  ", {k:(v["launches"],round(v["ms"]*1e3,1)) for k,v in d["kernels"].items()})
except Exception as e: print("bench ERR", e)
try:
    d=json.loads(open("gpurun_out/${tag}_train.json").read().strip().splitlines()[-1])
    print("cfg4", round(d["value"],2), "it/s", round(d["ms_per_step"],2), "ms", {k:(round(v["ms"],2), round(v["frac"],3)) for k,v in d["roofline_parts"].items()})
except Exception as e: print("train ERR", e)
PY
