# Multi-GPU check of both bench arms exactly as the driver launches them.   usage: bash tools/run_multi.sh <N> <tag>
N=${1:-2}; tag=${2:-n2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.log 2> gpurun_out/bench_ref_${tag}.err
tail -c 600 gpurun_out/bench_ref_${tag}.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 1000 --warmup 3 > gpurun_out/bench_${tag}.log 2> gpurun_out/bench_${tag}.err
tail -c 3000 gpurun_out/bench_${tag}.log; tail -c 1500 gpurun_out/bench_${tag}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --workload cfg5 --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_${tag}.log 2> gpurun_out/bench_cfg5_${tag}.err
tail -c 1200 gpurun_out/bench_cfg5_${tag}.log; tail -c 800 gpurun_out/bench_cfg5_${tag}.err
