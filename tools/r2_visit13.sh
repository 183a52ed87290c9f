#!/bin/bash
# round 2, visit 13: free-running batches 8 / 16 / rest, filter L2 hint A/B; parity suite first
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v13.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/pytest_v13.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-verify --no-files-leg"
timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_v13.json 2> gpurun_out/bench_c2_v13.err; show c2_v13
timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_v13.json 2> gpurun_out/bench_c3_v13.err; show c3_v13
SPRING_B200_FILTER_HINT=0 timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_nohint.json 2> gpurun_out/bench_c3_nohint.err; show c3_nohint
SPRING_B200_FILTER_HINT=0 timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_nohint.json 2> gpurun_out/bench_c2_nohint.err; show c2_nohint
SPRING_B200_FILTER_BITS=12 timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_fb12.json 2> gpurun_out/bench_c2_fb12.err; show c2_fb12
SPRING_B200_FILTER_BITS=12 timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_fb12.json 2> gpurun_out/bench_c3_fb12.err; show c3_fb12
timeout 300 python bench.py --config 5 --steps 3 $B > gpurun_out/bench_c5_v13.json 2> gpurun_out/bench_c5_v13.err; show c5_v13
echo "done at $(( $(date +%s) - T0 )) s"
