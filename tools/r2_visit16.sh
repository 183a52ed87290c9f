#!/bin/bash
# round 2, visit 16: contig-start experiments behind runtime flags (SPRING_B200_OPT bit 1: loads under the claim, bit 2: first read in smem)
set +e
mkdir -p gpurun_out
T0=$(date +%s)
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f ms/step %.2f chain_ms %.2f unmatched %d" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"]))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-verify --no-files-leg"
for o in 0 1 2 3; do
  SPRING_B200_OPT=$o timeout 300 python bench.py --config 5 --reads 20000000 --steps 3 $B > gpurun_out/bench_c5_o$o.json 2> gpurun_out/bench_c5_o$o.err; show c5_o$o
  SPRING_B200_OPT=$o timeout 200 python bench.py --config 2 --steps 5 $B > gpurun_out/bench_c2_o$o.json 2> gpurun_out/bench_c2_o$o.err; show c2_o$o
done
SPRING_B200_OPT=3 SPRING_B200_FAST_TAIL=1 timeout 300 python bench.py --config 5 --reads 20000000 --steps 3 $B > gpurun_out/bench_c5_o3ft.json 2> gpurun_out/bench_c5_o3ft.err; show c5_o3ft
SPRING_B200_OPT=0 SPRING_B200_FAST_TAIL=1 timeout 300 python bench.py --config 5 --reads 20000000 --steps 3 $B > gpurun_out/bench_c5_o0ft.json 2> gpurun_out/bench_c5_o0ft.err; show c5_o0ft
echo "done at $(( $(date +%s) - T0 )) s"
