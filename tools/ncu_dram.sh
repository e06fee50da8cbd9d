mkdir -p gpurun_out
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sector_hit_rate.pct -s 200 -c 44 --csv --log-file gpurun_out/ncu_dram.csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_dram.log 2>&1
tail -3 gpurun_out/ncu_dram.log
