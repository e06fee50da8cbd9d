#!/bin/bash
# round 2, first GPU visit: parity of the reworked loop, full-length chains, precision sweep, bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python tools/precision_sweep.py > gpurun_out/r2a_sweep.log 2>&1; tail -14 gpurun_out/r2a_sweep.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench20.json 2> gpurun_out/r2a_bench20.err; tail -c 600 gpurun_out/r2a_bench20.err
python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-also > gpurun_out/r2a_bench1000.json 2> gpurun_out/r2a_bench1000.err
for m in 2 4 1; do S3D_PRECISION=$m python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r2a_bench_mode$m.json 2> gpurun_out/r2a_bench_mode$m.err; done
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2>&1
head -c 1500 gpurun_out/r2a_bench20.json
