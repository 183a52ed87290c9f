#!/bin/bash
# round 2, visit 22: two-level key filter (first level L2-resident) -- parity, then a sweep of its size on configs 3, 5 and (forced) 2
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "two_level or free_running" > gpurun_out/pytest_v22.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v22.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
run() { # name config env...
  local name=$1 cfg=$2; shift 2
  env "$@" timeout 300 python bench.py --config $cfg --steps 3 $B > gpurun_out/bench_c${cfg}_$name.json 2> gpurun_out/bench_c${cfg}_$name.err; show c${cfg}_$name
}
run f1_0 3 SPRING_B200_FILTER1_BITS=0
run f1_15 3 SPRING_B200_FILTER1_BITS=1.5
run f1_2 3 SPRING_B200_FILTER1_BITS=2
run f1_3 3 SPRING_B200_FILTER1_BITS=3
run f1_4 3 SPRING_B200_FILTER1_BITS=4
run f1_0 5 SPRING_B200_FILTER1_BITS=0
run f1_2 5 SPRING_B200_FILTER1_BITS=2
run f1_3 5 SPRING_B200_FILTER1_BITS=3
run f1_0 2 SPRING_B200_FILTER1_BITS=0
run f1_3 2 SPRING_B200_FILTER1_BITS=3 SPRING_B200_FILTER1_MIN_MB=0
echo "bench done at $(( $(date +%s) - T0 )) s"
