#!/bin/bash
# round 2, visit 19: N reads unpacked on the GPU, contig split, production chain kernel without counters: suite + configs 2, 3 + launch list of config 3
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v19.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -4 gpurun_out/pytest_v19.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f (counting %.2f) unmatched %d verify %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["counted"]["ms_chain_kernel_counting"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_h2d","ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_v19.json 2> gpurun_out/bench_c2_v19.err; show c2_v19
timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_v19.json 2> gpurun_out/bench_c3_v19.err; show c3_v19
timeout 300 python bench.py --config 5 --steps 3 $B > gpurun_out/bench_c5_v19.json 2> gpurun_out/bench_c5_v19.err; show c5_v19
echo "bench done at $(( $(date +%s) - T0 )) s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c3_v19.csv python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline --no-files-leg --no-verify > gpurun_out/launches_c3_v19.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - T0 )) s"; wc -l gpurun_out/launches_c3_v19.csv
