#!/bin/bash
# round 2, visit 27 (2 GPUs): the 2-GPU tests (merged job decodes, spliced binary with SPRING_B200_GPUS=2) and bench.py under torchrun at N = 2
set +e
mkdir -p gpurun_out
T0=$(date +%s)
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v27.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_v27.log
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
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 240 $TR bench.py --gpus 2 --config 2 --steps 5 --no-cpu-baseline > gpurun_out/bench_n2_c2_v27.json 2> gpurun_out/bench_n2_c2_v27.err; show n2_c2_v27
timeout 400 $TR bench.py --gpus 2 --config 3 --steps 3 --no-cpu-baseline > gpurun_out/bench_n2_c3_v27.json 2> gpurun_out/bench_n2_c3_v27.err; show n2_c3_v27
echo "done at $(( $(date +%s) - T0 )) s"
