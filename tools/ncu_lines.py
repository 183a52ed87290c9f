"""Aggregate ncu warp-stall samples of one kernel per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
ncu's CLI prints per-SASS-instruction samples but no line numbers; nvdisasm -g prints the same
instruction sequence with '//## File "...", line N' markers.  The two are aligned by index.
"""
import csv, io, os, re, subprocess, sys, tempfile

rep, lib, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kname}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, rows = rows[hdr_i], rows[hdr_i + 1:]
si = hdr.index("# Samples")
samples = [int(r[si]) for r in rows if len(r) > si]
ii = hdr.index("Instructions Executed")
rows = [r for r in rows if len(r) > si and r[si].isdigit()]
insts = [int(r[ii]) for r in rows if len(r) > si]
sass = [r[1].strip() for r in rows if len(r) > si]

d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
lines = None
for f in sorted(os.listdir(d)):
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(d, f)], capture_output=True, text=True).stdout
    if kname not in txt:
        continue
    # isolate the function
    m = re.search(r"\.text\.[^\n]*" + re.escape(kname) + r"[^\n]*:\n", txt)
    start = m.end() if m else txt.index(kname)
    body = txt[start:]
    end = re.search(r"\n\s*\.section|\n//-+ \.text\.", body)
    body = body[: end.start()] if end else body
    cur, lines = None, []
    for ln in body.splitlines():
        mm = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if mm:
            cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            lines.append(cur)
    break
if lines is None:
    sys.exit("kernel not found in cubins")
print(f"sass rows {len(sass)}  disasm instrs {len(lines)}  total samples {sum(samples)}")
n = min(len(sass), len(lines))
agg = {}
iagg = {}
for i in range(n):
    agg[lines[i]] = agg.get(lines[i], 0) + samples[i]
    iagg[lines[i]] = iagg.get(lines[i], 0) + insts[i]
tot = sum(samples)
itot = sum(insts)
print(f"warp instructions executed: {itot}")
src_cache = {}
for (k, v) in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    text = ""
    if k:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", k[0])
        if p not in src_cache and os.path.exists(p):
            src_cache[p] = open(p).read().splitlines()
        if p in src_cache and k[1] - 1 < len(src_cache[p]):
            text = src_cache[p][k[1] - 1].strip()[:110]
    print(f"{100*v/tot:5.1f}%  inst {100*iagg[k]/itot:5.1f}%  {k}  {text}")
