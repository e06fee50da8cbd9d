# One GPU call: steady-state per-op times, the ncu launch list of the bench command, and ncu --set full of one step.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 1000 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_cfg2.txt > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 python bench.py --workload cfg5 --steps 200 --warmup 3 --no-cpu-baseline --dump-ops gpurun_out/ops_cfg5.txt > gpurun_out/bench_cfg5.log 2> gpurun_out/bench_cfg5.err
# launch list: bench.py --steps 2 --warmup 3 launches 5 steps on the device path, then the e2e and profile legs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
# one full step (21 launches) with the full metric set, after 2 warm steps
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_conv_tc|k_gn_silu|k_boundary|k_upcat|k_avgpool2' -s 42 -c 21 \
    -o gpurun_out/step_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
cat gpurun_out/ops_cfg2.txt
