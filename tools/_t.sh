mkdir -p gpurun_out
tag=${1:-d2}
timeout 600 python -m pytest tests/test_gpu_decoder.py -m gpu -x -q > gpurun_out/pytest_dec_$tag.log 2>&1; tail -3 gpurun_out/pytest_dec_$tag.log
timeout 300 python tools/bench_decoder.py --reso 256 --no-cpu-baseline > gpurun_out/dec_r256_$tag.log 2> gpurun_out/dec_r256_$tag.err; tail -c 400 gpurun_out/dec_r256_$tag.err
timeout 300 python tools/bench_decoder.py --reso 256 --precision 1 --no-cpu-baseline > gpurun_out/dec_r256_p1_$tag.log 2>&1
python - <<PY
import json
for f in ("dec_r256_$tag","dec_r256_p1_$tag"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, "Mpts/s %.1f"%(d["value"]/1e6), "ms %.2f"%d["ms_per_step"], "e2e Mpts/s %.1f"%(d["e2e"]["value"]/1e6), "planes ms %.3f"%d["config"]["feature_planes_ms"], "roof %.3f"%d["roofline"]["frac"], d.get("parity_vs_oracle_rel_l2"))
    except Exception as e: print(f, "failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_dec_mlp_tc' -c 1 -o gpurun_out/dec_mlp_$tag -f python tools/bench_decoder.py --reso 128 --iters 1 --no-cpu-baseline > gpurun_out/ncu_dec_$tag.log 2>&1; tail -2 gpurun_out/ncu_dec_$tag.log
