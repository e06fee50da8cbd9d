# One GPU call: parity tests, the default bench line (with the CPU baseline), the reference arm, the B=8 workloads,
# the ncu launch list of the bench command, and ncu --set full of one step.   usage: bash tools/run_profile.sh <tag>
tag=${1:-r1}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; tail -3 gpurun_out/pytest_$tag.log
timeout 400 python bench.py --steps 1000 --warmup 3 --dump-ops gpurun_out/ops_cfg2_$tag.txt > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.log 2> gpurun_out/bench_ref_$tag.err
timeout 300 python bench.py --workload cfg5 --steps 200 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_cfg5_$tag.txt > gpurun_out/bench_cfg5_$tag.log 2> gpurun_out/bench_cfg5_$tag.err
timeout 300 python bench.py --workload cfg3 --steps 100 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_cfg3_$tag.txt > gpurun_out/bench_cfg3_$tag.log 2> gpurun_out/bench_cfg3_$tag.err
# launch list: bench.py --steps 2 --warmup 3 launches 5 steps on the device path, then the e2e and profile legs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1
# one full step with the full metric set, after 2 warm steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc|k_gn_silu|k_boundary|k_upcat|k_avgpool2|k_sched' -s 40 -c 20 \
    -o gpurun_out/step_$tag -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out | tail -20
tail -c 1500 gpurun_out/bench_$tag.log
