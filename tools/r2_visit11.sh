#!/bin/bash
# round 2, visit 11: 32-bit-key dictionary sort: parity suite + bench lines
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v11.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -6 gpurun_out/pytest_v11.log
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
timeout 200 python bench.py --config 2 --steps 5 --no-cpu-baseline > gpurun_out/bench_c2_v11.json 2> gpurun_out/bench_c2_v11.err; show c2_v11
timeout 400 python bench.py --config 3 --steps 3 --no-cpu-baseline > gpurun_out/bench_c3_v11.json 2> gpurun_out/bench_c3_v11.err; show c3_v11
timeout 300 python bench.py --config 5 --steps 3 --no-cpu-baseline --no-files-leg > gpurun_out/bench_c5_v11.json 2> gpurun_out/bench_c5_v11.err; show c5_v11
echo "done at $(( $(date +%s) - T0 )) s"
