#!/bin/bash
# round 2, visit 37: singleton sweep templated on STITCH (run-time threshold had cost config 3's encoder 4 ms): stitch tests + config 3 line
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_verify.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "stitched_contigs or streams_match" > gpurun_out/pytest_v37.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -2 gpurun_out/pytest_v37.log
timeout 300 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-files-leg > gpurun_out/bench_c3_v37.json 2> gpurun_out/bench_c3_v37.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_c3_v37.json") if l.startswith("{")][-1])
print("c3 value %.1f e2e %.1f ms/step %.2f chain_ms %.2f verify %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["verify"]["ok"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_h2d","ms_dict","ms_chains","ms_encode")})
PY
echo "done at $(( $(date +%s) - T0 )) s"
