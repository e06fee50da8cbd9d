#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload cfg4 --steps 10 --warmup 3 > gpurun_out/r2p_train_n8.json 2> gpurun_out/r2p_train_n8.err
tail -c 400 gpurun_out/r2p_train_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2p_train_n8.json").read().strip().splitlines()[-1]); print("N=8 train", round(d["value"],1), "it/s", round(d["ms_per_step"],2), "ms", d["per_rank_ms"], "samples/s", round(d["detail"]["samples_per_s"]))
PY
