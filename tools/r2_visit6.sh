#!/bin/bash
# round 2, visit 6: warm-up probe A/B, full GPU suite (repeat-rich 10 M test included)
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short --durations=5 -p no:cacheprovider > gpurun_out/pytest_v6.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -12 gpurun_out/pytest_v6.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f files %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"], (d.get("e2e_files") or {}).get("ms_per_step")), {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
SPRING_B200_PREFETCH=1 timeout 150 python bench.py --config 2 --steps 5 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c2_pf1.json 2> gpurun_out/bench_c2_pf1.err; show c2_pf1
SPRING_B200_PREFETCH=2 timeout 150 python bench.py --config 2 --steps 5 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c2_pf2.json 2> gpurun_out/bench_c2_pf2.err; show c2_pf2
SPRING_B200_PREFETCH=1 timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c3_pf1.json 2> gpurun_out/bench_c3_pf1.err; show c3_pf1
SPRING_B200_PREFETCH=2 timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c3_pf2.json 2> gpurun_out/bench_c3_pf2.err; show c3_pf2
timeout 300 python bench.py --config 2 --steps 5 --no-cpu-baseline > gpurun_out/bench_c2_v6.json 2> gpurun_out/bench_c2_v6.err; show c2_v6
echo "done at $(( $(date +%s) - T0 )) s"
