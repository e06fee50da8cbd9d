# 8-GPU check: own arm at cfg2 (driver's scaling bench) and cfg5 (BASELINE configs[4]: 64 samples, 8 per GPU).  usage: bash tools/run_multi8.sh [N]
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 1000 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -c 700 gpurun_out/bench_n$N.log | head -c 400; echo; tail -c 600 gpurun_out/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus $N --workload cfg5 --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_n$N.log 2> gpurun_out/bench_cfg5_n$N.err
python - <<PY
import json
for f in ("bench_n$N", "bench_cfg5_n$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "steps/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), d["config"].get("sample_steps_per_s"))
    except Exception as e:
        print(f, "failed", e)
PY
