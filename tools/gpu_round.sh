# One GPU call of the edit-measure loop: parity tests, step timeline, bench line.  usage: bash tools/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -4 gpurun_out/pytest_$tag.log
timeout 200 python tools/trace_step.py --workload cfg2 > gpurun_out/trace_$tag.txt 2> gpurun_out/trace_$tag.err; tail -2 gpurun_out/trace_$tag.err
timeout 300 python bench.py --steps 1000 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_$tag.txt > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$tag.log").read().strip().splitlines()[-1])
    print("BENCH", "$tag", "steps/s", round(d["value"], 1), "us/step", round(d["ms_per_step"] * 1e3, 1), "e2e", round(d["e2e"]["value"], 1),
          "roof", d["roofline"] and round(d["roofline"]["frac"], 3))
except Exception as e:
    print("bench parse failed", e)
    print(open("gpurun_out/bench_$tag.err").read()[-2000:])
PY
