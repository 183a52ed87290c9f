"""Aggregate an ncu report of one kernel per CUDA source line (needs -lineinfo and --import-source on).

usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [top] [--by inst|samples]
Uses `ncu --page source --print-source cuda,sass --csv`: the rows whose Address is "-" are ncu's own
per-source-line aggregates (warp-stall samples, warp instructions executed), one section per file.
"""
import csv, io, os, subprocess, sys

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, kname = args[0], args[1]
top = int(args[2]) if len(args) > 2 else 40
by = "inst" if "--by" in sys.argv and sys.argv[sys.argv.index("--by") + 1] == "inst" else "samples"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{kname}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, agg = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    si, ii, ti = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    try:
        s, i, t = int(r[si]), int(r[ii]), int(r[ti])
    except ValueError:
        continue
    k = (cur_file, int(r[0]))
    a = agg.setdefault(k, [0, 0, 0, r[1].strip()])
    a[0] += s; a[1] += i; a[2] += t
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
tot_t = sum(a[2] for a in agg.values())
print(f"kernel {kname}: {tot_s} stall samples, {tot_i} warp instructions, {tot_t / tot_i:.1f} active threads per instruction")
key = (lambda kv: -kv[1][1]) if by == "inst" else (lambda kv: -kv[1][0])
print("samples%  inst%   thr/inst  file:line  source")
for k, a in sorted(agg.items(), key=key)[:top]:
    print(f"{100 * a[0] / tot_s:6.1f}  {100 * a[1] / tot_i:6.1f}  {a[2] / max(a[1], 1):6.1f}  {k[0]}:{k[1]}  {a[3][:100]}")
