#!/bin/bash
# round 2, visit 5: profiles -- launch lists (our kernels + CUB only), ncu --set full of k_chains and of the secondary kernels
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 600 --csv --log-file gpurun_out/launches_c2_v5.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/launches_c2_v5.log 2>&1
echo "ncu launches c2 exit $? at $(( $(date +%s) - T0 )) s"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/r02_chains_full python tools/chain_profile.py 4000000 > gpurun_out/ncu_full_v5.log 2>&1
echo "ncu full chains exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/ncu_full_v5.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_consensus|k_align_singletons|k_insert_slots|k_fill_bins|k_gather_sorted|k_noise|k_extract_keys" -c 12 -f -o gpurun_out/r02_secondary_full python tools/chain_profile.py 4000000 > gpurun_out/ncu_sec_v5.log 2>&1
echo "ncu full secondary exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/ncu_sec_v5.log
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:k_chains -s 1 -c 1 --csv --log-file gpurun_out/chains_traffic_c2.csv python tools/chain_profile.py 10000000 > gpurun_out/chains_traffic_c2.log 2>&1
echo "ncu traffic exit $? at $(( $(date +%s) - T0 )) s"; tail -6 gpurun_out/chains_traffic_c2.csv | cut -c1-300
