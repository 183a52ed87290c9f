#!/bin/bash
# round 2, visit 21: packed records / compile-time knobs; early slot prefetch and batch-0 size A/B (tunable instantiation); launch list of config 3; ncu of the production chain kernel
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "free_running or streams_match or reorder_stream or edge" > gpurun_out/pytest_v21.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v21.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f (counting %.2f) unmatched %d verify %s frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["counted"]["ms_chain_kernel_counting"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
run() { # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 3 $B > gpurun_out/bench_c${cfg}_$name.json 2> gpurun_out/bench_c${cfg}_$name.err; show c${cfg}_$name
}
for cfg in 2 3; do
  run v21 $cfg X=1
  run pf1 $cfg SPRING_B200_PREFETCH=1
  run pf2 $cfg SPRING_B200_PREFETCH=2
  run pf3 $cfg SPRING_B200_PREFETCH=3
  run b02 $cfg SPRING_B200_BATCH0=2
  run b02pf2 $cfg SPRING_B200_BATCH0=2 SPRING_B200_PREFETCH=2
done
echo "bench done at $(( $(date +%s) - T0 )) s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 700 --csv --log-file gpurun_out/launches_c3_v21.csv python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline --no-files-leg --no-verify > gpurun_out/launches_c3_v21.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - T0 )) s"; wc -l gpurun_out/launches_c3_v21.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/r02_chains_prod python tools/chain_profile.py 4000000 > gpurun_out/ncu_prod_v21.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/ncu_prod_v21.log
