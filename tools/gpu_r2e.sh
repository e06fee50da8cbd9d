#!/bin/bash
# r2e: full GPU suite on the current build, helper-kernel bandwidths, bench lines with per-op dumps, ncu launch list + full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
python tools/bench_aux.py > gpurun_out/r2e_bench_aux.jsonl 2> gpurun_out/r2e_bench_aux.err; cut -c1-200 gpurun_out/r2e_bench_aux.jsonl; tail -3 gpurun_out/r2e_bench_aux.err
python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r2e_ops_cfg2.txt > gpurun_out/r2e_bench20.json 2> gpurun_out/r2e_bench20.err; tail -c 300 gpurun_out/r2e_bench20.err
python bench.py --workload cfg5 --steps 100 --warmup 5 --no-cpu-baseline --no-also --dump-ops gpurun_out/r2e_ops_cfg5.txt > gpurun_out/r2e_bench_cfg5.json 2> gpurun_out/r2e_bench_cfg5.err
python bench.py --workload cfg3 --steps 100 --warmup 5 --no-cpu-baseline --no-also --dump-ops gpurun_out/r2e_ops_cfg3.txt > gpurun_out/r2e_bench_cfg3.json 2> gpurun_out/r2e_bench_cfg3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_cfg2.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2e_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv_tc -s 16 -c 8 -o gpurun_out/r2e_conv_full -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-also > gpurun_out/r2e_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_wgrad_tc|k_gn_bwd|k_grad_stage|k_roll_bwd" -c 14 -o gpurun_out/r2e_train_full -f python bench.py --workload cfg4 --batch 8 --steps 1 --warmup 3 > gpurun_out/r2e_ncu_train.log 2>&1
ls -la gpurun_out/*.ncu-rep
head -c 1200 gpurun_out/r2e_bench20.json
