#!/bin/bash
# round 2, visit 1: host facts, new verify tests + parity suite, bench on configs 2 and 3, reference arm on config 3
set +e
mkdir -p gpurun_out
T0=$(date +%s)
{ nproc; free -g | head -2; nvidia-smi -L; df -h /dev/shm | tail -1; } > gpurun_out/host.txt 2>&1
cat gpurun_out/host.txt
timeout 900 python -m pytest tests/test_gpu_verify.py tests/test_gpu_parity.py -m gpu -q --maxfail=8 --tb=short --durations=8 > gpurun_out/pytest_v1.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -15 gpurun_out/pytest_v1.log
timeout 300 python bench.py --config 2 --steps 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench c2 exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
timeout 600 python bench.py --config 3 --steps 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
echo "bench c3 exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_c3.json; tail -3 gpurun_out/bench_c3.err
SPRING_B200_CHAIN_DBG=1 timeout 300 python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/bench_c3_dbg.json 2> gpurun_out/bench_c3_dbg.err
tail -3 gpurun_out/bench_c3_dbg.err
timeout 700 python bench.py --impl reference --config 3 --steps 20 --warmup 5 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err
echo "ref c3 exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_ref_c3.json; tail -3 gpurun_out/bench_ref_c3.err
