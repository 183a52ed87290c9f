#!/bin/bash
# round 2, visit 38: final bench lines (config 3 as the driver runs it, config 2 with its files leg)
set +e
mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/bench_final_c3.json 2> gpurun_out/bench_final_c3.err; tail -c 300 gpurun_out/bench_final_c3.json
timeout 60 python bench.py --config 2 --no-cpu-baseline > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; tail -c 300 gpurun_out/bench_final_c2.json
