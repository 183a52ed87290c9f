#!/bin/bash
# GPU-box visit: ncu launch list (our kernels + CUB only) of the hot path at 10 M reads
set +e
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|^Device' -c 900 --csv --log-file gpurun_out/launches.csv python tools/chain_profile.py 10000000 > gpurun_out/launches_run.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/launches_run.log
