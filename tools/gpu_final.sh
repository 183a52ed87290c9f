#!/bin/bash
# GPU-box visit at the end of a session: smoke, the bench line for profiles/, launch list, ncu captures of k_chains
set +e
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -4 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; cut -c1-400 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref arm exit $?"; cut -c1-300 gpurun_out/bench_ref.json
bash tools/gpu_launches.sh
bash tools/gpu_profile.sh
