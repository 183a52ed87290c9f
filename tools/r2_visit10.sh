#!/bin/bash
# round 2, visit 10 (4 GPUs): bench.py under torchrun at N = 4 (config 2 at 10 M reads per GPU, config 3 at 50 M per GPU)
set +e
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L | head -8
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s exchange %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d.get("exchange_ms")), {k: round(v, 2) for k, v in d["stages_ms"].items()})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-2500:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 240 $TR bench.py --gpus 4 --config 2 --steps 5 > gpurun_out/bench_n4_c2.json 2> gpurun_out/bench_n4_c2.err; show n4_c2
timeout 400 $TR bench.py --gpus 4 --reads 50000000 --steps 3 > gpurun_out/bench_n4_c3_50M.json 2> gpurun_out/bench_n4_c3_50M.err; show n4_c3_50M
echo "done at $(( $(date +%s) - T0 )) s"
