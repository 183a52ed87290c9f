#!/bin/bash
# round 2, visit 9 (2 GPUs): multi-GPU tests with the final kernels, N = 2 bench lines, sanitizer on GPU 0
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 400 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_parity.py::test_bucket_kernel_matches_numpy_mirror" -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_v9.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -8 gpurun_out/pytest_v9.log
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
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --config 2 --steps 5 > gpurun_out/bench_n2_c2_v9.json 2> gpurun_out/bench_n2_c2_v9.err; show n2_c2_v9
timeout 400 $TR bench.py --gpus 2 --steps 3 > gpurun_out/bench_n2_c3_v9.json 2> gpurun_out/bench_n2_c3_v9.err; show n2_c3_v9
echo "benches at $(( $(date +%s) - T0 )) s"
CUDA_VISIBLE_DEVICES=0 timeout 420 compute-sanitizer --tool memcheck python tools/sanitize_run.py se100_n var64_noisy pe100_illumina > gpurun_out/memcheck_v9.log 2>&1 &
CUDA_VISIBLE_DEVICES=1 timeout 420 compute-sanitizer --tool racecheck python tools/sanitize_run.py se100_n pe100_illumina > gpurun_out/racecheck_v9.log 2>&1 &
wait
echo "sanitizers done at $(( $(date +%s) - T0 )) s"; tail -4 gpurun_out/memcheck_v9.log; tail -4 gpurun_out/racecheck_v9.log
