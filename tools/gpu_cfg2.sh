#!/bin/bash
set +e
mkdir -p gpurun_out
for cfg in $CFGS; do
  SPRING_B200_KCFG=$cfg timeout 300 python bench.py --no-cpu-baseline --steps 4 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err
  echo "cfg $cfg exit $?"
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_$cfg.json"))
    print("$cfg", "value", round(j["value"], 1), "ms", round(j["ms_per_step"], 2), "chains", j["chains"], "unmatched", j["unmatched"], {k: round(v, 2) for k, v in j["stages_ms"].items()})
except Exception as e:
    print("$cfg failed", e)
PY
done
