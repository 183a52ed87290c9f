#!/bin/bash
# GPU-box visit: ncu captures of the chain kernel for profiles/ (full set at 4 M reads with source, DRAM traffic at 10 M)
set +e
mkdir -p gpurun_out
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/chains_full python tools/chain_profile.py 4000000 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
timeout 500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none -k regex:k_chains -c 1 --csv --log-file gpurun_out/chains_traffic.csv python tools/chain_profile.py 10000000 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $?"; tail -4 gpurun_out/chains_traffic.csv
