#!/bin/bash
# round 2, visit 12: key-filter size against the L2 (config 3: two filters of 100 M keys), parity suite, ratio policies
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 500 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v12.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -6 gpurun_out/pytest_v12.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f unmatched %d frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], d["roofline"]["frac"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
for fb in 3 4 5 6 8; do
  SPRING_B200_FILTER_BITS=$fb timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c3_fb$fb.json 2> gpurun_out/bench_c3_fb$fb.err; show c3_fb$fb
done
for fb in 4 6 8; do
  SPRING_B200_FILTER_BITS=$fb timeout 200 python bench.py --config 2 --steps 5 --no-cpu-baseline --no-verify --no-files-leg > gpurun_out/bench_c2_fb$fb.json 2> gpurun_out/bench_c2_fb$fb.err; show c2_fb$fb
done
echo "filter sweep done at $(( $(date +%s) - T0 )) s"
timeout 400 python tests/tools/ratio_check.py 4000000 --policy > gpurun_out/ratio_policy_4M.txt 2>&1; cat gpurun_out/ratio_policy_4M.txt | grep -v "stage seconds"
echo "done at $(( $(date +%s) - T0 )) s"
