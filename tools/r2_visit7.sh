#!/bin/bash
# round 2, visit 7: bit-sliced consensus + align ILP: parity suite, then the bench lines as the driver runs them
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v7.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -12 gpurun_out/pytest_v7.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f files %s cpu %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"], (d.get("e2e_files") or {}).get("ms_per_step"), (d.get("cpu_baseline") or {}).get("value")), {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
timeout 300 python bench.py --config 2 --steps 10 > gpurun_out/bench_c2_v7.json 2> gpurun_out/bench_c2_v7.err; show c2_v7
echo "at $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3_v7.json 2> gpurun_out/bench_c3_v7.err; show c3_v7
echo "at $(( $(date +%s) - T0 )) s"
timeout 400 python bench.py --config 5 --steps 5 --no-cpu-baseline > gpurun_out/bench_c5_v7.json 2> gpurun_out/bench_c5_v7.err; show c5_v7
echo "at $(( $(date +%s) - T0 )) s"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_|Device" -c 600 --csv --log-file gpurun_out/launches_c2_v7.csv python bench.py --config 2 --steps 1 --warmup 1 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/launches_c2_v7.log 2>&1
echo "ncu launches exit $? at $(( $(date +%s) - T0 )) s"
