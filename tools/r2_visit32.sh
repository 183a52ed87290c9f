#!/bin/bash
# round 2, visit 32: single-writer discipline for the chain's shared cold state: racecheck again, parity, bench lines
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_run.py se100_n pe100_illumina > gpurun_out/racecheck_final.log 2>&1 &
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_verify.py -m gpu -q -x --tb=short -p no:cacheprovider -k "not end_to_end and not drop_in" > gpurun_out/pytest_v32.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v32.log
wait
echo "racecheck at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/racecheck_final.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"]))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
for cfg in 2 3 5; do
  timeout 300 python bench.py --config $cfg --steps 3 $B > gpurun_out/bench_c${cfg}_v32.json 2> gpurun_out/bench_c${cfg}_v32.err; show c${cfg}_v32
done
echo "done at $(( $(date +%s) - T0 )) s"
