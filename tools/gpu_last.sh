#!/bin/bash
# last GPU-box visit of a session: the whole GPU suite on the committed state + the ncu launch list of bench.py itself
set +e
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --maxfail=5 --tb=short > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k_|^Device' -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu bench exit $?"
