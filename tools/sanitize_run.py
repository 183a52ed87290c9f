"""Small run of every kernel for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import CASES, make_input
from spring_b200 import capi
ctx = capi.Context(0)
for name in ("se100_n", "var64_noisy", "long511", "heavy_bins"):
    kw = dict(CASES[name]); kw["num_reads"] = min(kw["num_reads"], 3000)
    hp = make_input(**kw)
    for det in (False, True):
        ctx.set_schedule(det)
        for chains in (1, 16):
            s = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
    print(name, "ok", s.num_aligned)
