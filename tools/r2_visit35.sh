#!/bin/bash
# round 2, visit 35: contig stitching by contig ends (head and tail, 64 bases) -- tests, Reads: bytes with and without (4 M reads), bench config 2 with and without
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests/test_gpu_verify.py tests/test_gpu_parity.py -m gpu -q -x --tb=short -p no:cacheprovider -k "stitch or free_running or end_to_end or drop_ins or deep" > gpurun_out/pytest_v35.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -25 gpurun_out/pytest_v35.log | cut -c1-220
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v35.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke_v35.log
timeout 300 python tests/tools/ratio_check.py 4000000 --stitch > gpurun_out/ratio_stitch_4M.txt 2>&1; echo "ratio exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/ratio_stitch_4M.txt
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f ms/step %.2f chain_ms %.2f verify %s contigs %d stitched %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], (d.get("verify") or {}).get("ok"), d["contigs"], d["contigs_stitched"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
B="--no-cpu-baseline --no-files-leg"
timeout 300 python bench.py --config 2 --steps 3 $B > gpurun_out/bench_c2_v35auto.json 2> gpurun_out/bench_c2_v35auto.err; show c2_v35auto
SPRING_B200_STITCH=0 timeout 300 python bench.py --config 2 --steps 3 $B > gpurun_out/bench_c2_v35off.json 2> gpurun_out/bench_c2_v35off.err; show c2_v35off
echo "done at $(( $(date +%s) - T0 )) s"
