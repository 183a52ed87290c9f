#!/bin/bash
# GPU-box visit: parity tests, then the bench under each chain-kernel launch configuration.
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=6 --tb=short > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in 8x3 8x4 4x7; do
  SPRING_B200_KCFG=$cfg timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "cfg $cfg exit $?"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_$cfg.json"))
    print("$cfg", "value", round(j["value"], 1), "ms", round(j["ms_per_step"], 2), "chains", j["chains"], {k: round(v, 2) for k, v in j["stages_ms"].items()})
except Exception as e:
    print("$cfg failed", e)
PY
done
