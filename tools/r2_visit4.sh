#!/bin/bash
# round 2, visit 4 (2 GPUs): the library's exchange / finalize / merge on hardware, torchrun bench at N = 2
set +e
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v4.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -30 gpurun_out/pytest_v4.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/bench_{n}.json"))
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f chains %d unmatched %d verify %s exchange %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["chains"], d["unmatched"], (d.get("verify") or {}).get("ok"), d.get("exchange_ms")), {k: round(v, 2) for k, v in d["stages_ms"].items()})
    print("   layout", d.get("shard_layout_rank0"))
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-2500:])
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --config 2 --steps 5 > gpurun_out/bench_n2_c2.json 2> gpurun_out/bench_n2_c2.err; show n2_c2
echo "at $(( $(date +%s) - T0 )) s"
timeout 400 $TR bench.py --gpus 2 --config 3 --steps 3 > gpurun_out/bench_n2_c3.json 2> gpurun_out/bench_n2_c3.err; show n2_c3
echo "at $(( $(date +%s) - T0 )) s"
NCCL_DEBUG=INFO timeout 300 $TR bench.py --gpus 2 --config 2 --steps 2 --warmup 1 --no-verify 2>&1 | grep -E "NVLS|P2P|via|Connected|channels" | head -12 > gpurun_out/nccl_info_n2.txt; head -12 gpurun_out/nccl_info_n2.txt
echo "done at $(( $(date +%s) - T0 )) s"
