set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1000 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err
S3D_FUSED_BOUNDARY=1 timeout 300 python bench.py --steps 1000 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2> gpurun_out/bench_fused.err
timeout 300 python bench.py --workload cfg3 --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.log 2> gpurun_out/bench_cfg3.err
tail -c 3000 gpurun_out/bench.log; tail -c 1500 gpurun_out/bench_fused.log
