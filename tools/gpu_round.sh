#!/bin/bash
# One GPU-box visit: parity tests, bench (new / generic-update cross-check), ncu captures.  Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 --tb=short --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_n1.json
SPRING_B200_GENERIC_UPDATE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/bench_generic_update.json 2> gpurun_out/bench_generic_update.err
echo "generic-update bench exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_generic_update.json
if [ "$1" != "nocap" ]; then
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/chains_full python tools/chain_profile.py 4000000 > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "ncu launches exit $? at $(( $(date +%s) - T0 )) s"
fi
