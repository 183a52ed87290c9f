#!/bin/bash
# round 2, visit 23: branch-free count pass, first hashed key kept for pass 2 (A/B against a build without it); launch list of config 3; ncu of the production kernel at 10 M reads
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_verify.py tests/test_update_ref_model.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v23.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v23.log
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
for cfg in 2 3 5; do
  run v23 $cfg X=1
  run v23nohk $cfg SPRING_B200_LIB=$PWD/spring_b200/libspring_b200_nohk.so
done
echo "bench done at $(( $(date +%s) - T0 )) s"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3_v23.csv python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline --no-files-leg --no-verify --profile-after-setup > gpurun_out/launches_c3_v23.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - T0 )) s"; wc -l gpurun_out/launches_c3_v23.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/r02_chains_prod10M python tools/chain_profile.py 10000000 > gpurun_out/ncu_prod_v23.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/ncu_prod_v23.log
