#!/bin/bash
# round 2, visit 18: more chains per SM where the kernel is latency-bound (config 5: 30 % issue utilisation): launch-bound variants
set +e
mkdir -p gpurun_out
T0=$(date +%s)
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f chains %d unmatched %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["chains"], d["unmatched"]))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-verify --no-files-leg"
for k in 8x4 8x5 8x6 4x9; do
  SPRING_B200_KCFG=$k timeout 300 python bench.py --config 5 --reads 20000000 --steps 3 $B > gpurun_out/bench_c5_$k.json 2> gpurun_out/bench_c5_$k.err; show c5_$k
done
SPRING_B200_GENERIC_W=1 timeout 300 python bench.py --config 5 --reads 20000000 --steps 3 $B > gpurun_out/bench_c5_gw.json 2> gpurun_out/bench_c5_gw.err; show c5_gw
for k in 8x5 8x6; do
  SPRING_B200_KCFG=$k timeout 300 python bench.py --config 3 --reads 30000000 --steps 3 $B > gpurun_out/bench_c3_$k.json 2> gpurun_out/bench_c3_$k.err; show c3_$k
done
timeout 300 python bench.py --config 3 --reads 30000000 --steps 3 $B > gpurun_out/bench_c3_30M.json 2> gpurun_out/bench_c3_30M.err; show c3_30M
echo "done at $(( $(date +%s) - T0 )) s"
