#!/bin/bash
# round 2, visit 31 (final): full suite, smoke, bench lines of configs 3 / 2 / 5 as the driver runs them, launch list of config 3,
# ncu --set full + DRAM traffic of the production chain kernel, compute-sanitizer
set +e
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $? after $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/pytest_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke_final.log
show() { python - "$1" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/bench_{n}.json") if l.startswith("{")][-1])
    print(n, "value %.1f e2e %.1f files %s ms/step %.2f chain_ms %.2f unmatched %d verify %s frac %.3f cpu %s" % (d["value"], d["e2e"]["value"], (d.get("e2e_files") or {}).get("value"), d["ms_per_step"], d["roofline"]["kernel_ms"], d["unmatched"], (d.get("verify") or {}).get("ok"), d["roofline"]["frac"], (d.get("cpu_baseline") or {}).get("value")), {k: round(v, 2) for k, v in d["stages_ms"].items() if k in ("ms_h2d","ms_dict","ms_chains","ms_encode")})
except Exception as e:
    print(n, "failed", e); print(open(f"gpurun_out/bench_{n}.err").read()[-1500:])
PY
}
timeout 600 python bench.py > gpurun_out/bench_final_c3.json 2> gpurun_out/bench_final_c3.err; show final_c3
timeout 300 python bench.py --config 2 --no-cpu-baseline > gpurun_out/bench_final_c2.json 2> gpurun_out/bench_final_c2.err; show final_c2
timeout 300 python bench.py --config 5 --steps 3 --no-cpu-baseline --no-files-leg > gpurun_out/bench_final_c5.json 2> gpurun_out/bench_final_c5.err; show final_c5
echo "bench done at $(( $(date +%s) - T0 )) s"
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_c3_final.csv python bench.py --config 3 --steps 1 --warmup 0 --no-cpu-baseline --no-files-leg --no-verify --profile-after-setup > gpurun_out/launches_c3_final.log 2>&1
echo "launch list exit $? at $(( $(date +%s) - T0 )) s"; wc -l gpurun_out/launches_c3_final.csv
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_chains -c 1 -f -o gpurun_out/r02_chains_final python tools/chain_profile.py 10000000 > gpurun_out/ncu_final.log 2>&1
echo "ncu full exit $? at $(( $(date +%s) - T0 )) s"
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct
timeout 200 ncu --metrics $M --clock-control none -k regex:k_chains -s 1 -c 1 --csv --log-file gpurun_out/chains_traffic_c2_final.csv python tools/chain_profile.py 10000000 > /dev/null 2>&1
timeout 300 ncu --profile-from-start off --metrics $M --clock-control none -k regex:k_chains -s 1 -c 1 --csv --log-file gpurun_out/chains_traffic_c3_final.csv python bench.py --config 3 --steps 1 --warmup 1 --no-cpu-baseline --no-files-leg --no-verify --profile-after-setup > /dev/null 2>&1
echo "traffic exit $? at $(( $(date +%s) - T0 )) s"; tail -5 gpurun_out/chains_traffic_c3_final.csv | cut -d, -f13-15
SPRING_B200_CONTIG_SPLIT=7 timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_run.py se100_n pe100_illumina > gpurun_out/memcheck_final.log 2>&1
echo "memcheck exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/memcheck_final.log
timeout 300 compute-sanitizer --tool racecheck python tools/sanitize_run.py se100_n > gpurun_out/racecheck_final.log 2>&1
echo "racecheck exit $? at $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/racecheck_final.log
