#!/bin/bash
# round 2, visit 14: which of the two chain-kernel changes of visit 13 cost time: default (neither) vs fast tail
set +e
mkdir -p gpurun_out
T0=$(date +%s)
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f unmatched %d frac %.3f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], d["roofline"]["frac"]))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-verify --no-files-leg"
timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_v14.json 2> gpurun_out/bench_c2_v14.err; show c2_v14
SPRING_B200_FAST_TAIL=1 timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_ft.json 2> gpurun_out/bench_c2_ft.err; show c2_ft
timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_v14.json 2> gpurun_out/bench_c3_v14.err; show c3_v14
SPRING_B200_FAST_TAIL=1 timeout 300 python bench.py --config 3 --steps 3 $B > gpurun_out/bench_c3_ft.json 2> gpurun_out/bench_c3_ft.err; show c3_ft
timeout 300 python bench.py --config 5 --steps 3 $B > gpurun_out/bench_c5_v14.json 2> gpurun_out/bench_c5_v14.err; show c5_v14
SPRING_B200_FAST_TAIL=1 timeout 300 python bench.py --config 5 --steps 3 $B > gpurun_out/bench_c5_ft.json 2> gpurun_out/bench_c5_ft.err; show c5_ft
echo "done at $(( $(date +%s) - T0 )) s"
