#!/bin/bash
# round 2, visit 3: hash-ordered dictionary build, file-level entry, verify tests, prefetch variant; tight timeouts
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 420 python -m pytest tests -m gpu -q -x --tb=short --durations=6 -p no:cacheprovider > gpurun_out/pytest_v3.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -25 gpurun_out/pytest_v3.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f chains %d unmatched %d verify %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["chains"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1200:])
PY
}
timeout 150 python bench.py --config 2 --steps 5 --no-cpu-baseline > gpurun_out/bench_c2_v3.json 2> gpurun_out/bench_c2_v3.err; show c2_v3
SPRING_B200_PREFETCH=1 timeout 150 python bench.py --config 2 --steps 5 --no-cpu-baseline --no-verify > gpurun_out/bench_c2_v3pf.json 2> gpurun_out/bench_c2_v3pf.err; show c2_v3pf
timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3_v3.json 2> gpurun_out/bench_c3_v3.err; show c3_v3
SPRING_B200_PREFETCH=1 timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-verify > gpurun_out/bench_c3_v3pf.json 2> gpurun_out/bench_c3_v3pf.err; show c3_v3pf
timeout 300 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/bench_c5_v3.json 2> gpurun_out/bench_c5_v3.err; show c5_v3
echo "benches done at $(( $(date +%s) - T0 )) s"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_v3.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/launches_c2_v3.log 2>&1
echo "ncu launches exit $? at $(( $(date +%s) - T0 )) s"
