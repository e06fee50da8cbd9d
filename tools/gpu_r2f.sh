#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2f_bench_n8.json 2> gpurun_out/r2f_bench_n8.err
tail -c 600 gpurun_out/r2f_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 1000 --warmup 20 --no-also > gpurun_out/r2f_bench_n8_1000.json 2> gpurun_out/r2f_bench_n8_1000.err
python - <<PY
import json
for f in ("gpurun_out/r2f_bench_n8.json","gpurun_out/r2f_bench_n8_1000.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"],1), "per_rank_ms", d["per_rank_ms"], "e2e", round(d["e2e"]["value"],1))
        for k,v in (d.get("also") or {}).items():
            if isinstance(v,dict): print("  ",k, round(v["value"],1), v.get("sample_steps_per_s"), v["per_rank_ms"])
    except Exception as e: print(f,"ERR",e)
PY
