"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one pass of the hot path
(the launches between two consecutive k_pack_seq), shares per kernel.
usage: launch_summary.py launches.csv [pass_index]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, rows = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for r in rows:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[r[mu]]
    seq.append((r[kn], v))
ends = [i for i, (n, _) in enumerate(seq) if "k_pack_seq" in n]
want = int(sys.argv[2]) if len(sys.argv) > 2 else len(ends) - 1
lo = ends[want - 1] + 1 if want > 0 else 0
one = seq[lo: ends[want] + 1]
tot = sum(v for _, v in one)
ag, cn = collections.Counter(), collections.Counter()
for n, v in one:
    short = re.sub(r"^void ", "", n)
    short = re.sub(r"\(.*", "", short)
    short = re.sub(r"sb::(<unnamed>::)?", "", short)
    short = short[:100]
    ag[short] += v; cn[short] += 1
print(f"pass {want}: {len(one)} launches, {tot:.3f} ms summed (serialised, cold cache: compare shares)")
for k, v in ag.most_common(40):
    print(f"{100 * v / tot:6.2f}% {v:8.3f} ms x{cn[k]:3d}  {k}")
