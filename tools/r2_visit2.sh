#!/bin/bash
# round 2, visit 2: chains2.cu (sub-warp chains, bit-sliced counts): parity, launch-bound variants, ncu capture
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_verify.py -m gpu -q --maxfail=6 --tb=short > gpurun_out/pytest_v2.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -40 gpurun_out/pytest_v2.log
B="timeout 300 python bench.py --config 2 --steps 3 --warmup 2 --no-cpu-baseline"
for v in "default:" "minb3:SPRING_B200_MINB=3" "minb2:SPRING_B200_MINB=2" "lanes32:SPRING_B200_LANES=32" "lanes32minb2:SPRING_B200_LANES=32 SPRING_B200_MINB=2" "v1:SPRING_B200_CHAINS_V1=1" "half:SPRING_B200_MAX_CHAINS=4736"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs $B > gpurun_out/bench_c2_$name.json 2> gpurun_out/bench_c2_$name.err
  echo "== $name rc=$? at $(( $(date +%s) - T0 )) s"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_c2_$name.json"))
    print("$name", "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f chains %d unmatched %d verify %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["chains"], d["unmatched"], d["verify"]["ok"], d["roofline"]["frac"]), d["stages_ms"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/bench_c2_$name.err").read()[-1500:])
PY
done
timeout 600 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_v2.json 2> gpurun_out/bench_c3_v2.err
echo "bench c3 exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/bench_c3_v2.json; tail -3 gpurun_out/bench_c3_v2.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/chains2_full python tools/chain_profile.py 4000000 > gpurun_out/ncu_full_v2.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/ncu_full_v2.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py > gpurun_out/memcheck_v2.log 2>&1
echo "memcheck exit $? at $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/memcheck_v2.log
