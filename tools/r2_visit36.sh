#!/bin/bash
# round 2, visit 36 (final state): full suite, bench lines of configs 3 and 2 as the driver runs them, Reads: bytes with every co-resident chain + stitching
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_final.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f files %s ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f stitched %d/%d" % (d["value"], d["e2e"]["value"], (d.get("e2e_files") or {}).get("value"), d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"], d["contigs_stitched"], d["contigs"]), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_h2d","ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
timeout 600 python bench.py > gpurun_out/bench_final_c3.json 2> gpurun_out/bench_final_c3.err; show final_c3
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; show final_c2
echo "bench done at $(( $(date +%s) - T0 )) s"
timeout 200 python tests/tools/ratio_check.py 4000000 --stitch --all-chains > gpurun_out/ratio_stitch_allchains_4M.txt 2>&1; echo "ratio exit $? at $(( $(date +%s) - T0 )) s"; cat gpurun_out/ratio_stitch_allchains_4M.txt
