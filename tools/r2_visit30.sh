#!/bin/bash
# round 2, visit 30: contig starts through the out-of-line reset_ref (one inlined copy of update_ref_fast instead of four): A/B
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_verify.py tests/test_update_ref_model.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not end_to_end and not drop_in" > gpurun_out/pytest_v30.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v30.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]))
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
  run v30 $cfg X=1
  run v30noreset $cfg SPRING_B200_LIB=$PWD/spring_b200/libspring_b200_noreset.so
done
echo "bench done at $(( $(date +%s) - T0 )) s"
