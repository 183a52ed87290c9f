#!/bin/bash
# round 2, visit 15 (8 GPUs): bench.py under torchrun exactly as the driver's scaling run starts it, default config (100 M reads per GPU)
set +e
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench_n8_c3.json 2> gpurun_out/bench_n8_c3.err
echo "exit $? at $(( $(date +%s) - T0 )) s"
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_n8_c3.json") if l.startswith("{")][-1])
    print("n8_c3 value %.1f e2e %.1f ms/step %.2f chain_ms %.2f unmatched %d verify %s exchange %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d.get("exchange_ms")), {k: round(v, 2) for k, v in d["stages_ms"].items()})
    print(d.get("shard_layout_rank0"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/bench_n8_c3.err").read()[-3000:])
PY
