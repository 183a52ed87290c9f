#!/bin/bash
# DRAM traffic of k_chains on the default bench workload (config 3, 100 M reads): one launch under ncu
mkdir -p gpurun_out
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_chains -s 1 -c 1 --csv --log-file gpurun_out/chains_traffic_c3.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/chains_traffic_c3.log 2>&1
echo "ncu traffic c3 exit $?"; tail -6 gpurun_out/chains_traffic_c3.csv | cut -c1-300
