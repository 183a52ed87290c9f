#!/bin/bash
# round 2, visit 17: final-candidate kernel (fast tail on, policies from the constant bank): suite + all configs
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v17.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -4 gpurun_out/pytest_v17.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_v17.json 2> gpurun_out/bench_c2_v17.err; show c2_v17
timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_v17.json 2> gpurun_out/bench_c3_v17.err; show c3_v17
timeout 300 python bench.py --config 5 --steps 3 $B > gpurun_out/bench_c5_v17.json 2> gpurun_out/bench_c5_v17.err; show c5_v17
echo "done at $(( $(date +%s) - T0 )) s"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/r02_chains_var python tools/chain_profile.py 4000000 var > gpurun_out/ncu_var_v17.log 2>&1
echo "ncu var exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/ncu_var_v17.log
