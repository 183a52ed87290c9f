"""Debug helper (GPU box): first divergence between the CUDA reorder stream and the oracle, per case."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from helpers import CASES, make_input
from oracle import pyoracle as po
from spring_b200 import capi

ctx = capi.Context(0)
names = sys.argv[1:] or sorted(CASES)
for name in names:
    hp = make_input(**CASES[name])
    for chains in (1, 5, 64):
        try:
            order, flag, pos, rev, s_order = ctx.reorder(hp.packed, hp.lengths, hp.max_readlen, chains)
        except Exception as e:
            print(name, chains, "ERROR", e); continue
        st = ctx.stats()
        ro = po.reorder(hp.packed, hp.lengths, hp.max_readlen, st["num_chains"])
        n = min(len(order), len(ro.order))
        bad = np.nonzero((order[:n] != ro.order[:n]) | (flag[:n] != ro.flag[:n]) | (pos[:n] != ro.pos[:n]) | (rev[:n] != ro.rc[:n]))[0]
        ok = len(order) == len(ro.order) and len(bad) == 0 and (s_order == ro.s_order).all()
        print(f"{name:16s} C={st['num_chains']:3d} L={hp.max_readlen} n={len(hp.lengths)} gpu_aligned={len(order)} orc_aligned={len(ro.order)} "
              f"rounds={st['rounds']}/{ro.counters['rounds']} probes_seq={st['probes_seq']}/{ro.counters['probes']} "
              f"cmp={st['compares']}/{ro.counters['compares']} unmatched={st['unmatched']}/{ro.counters['unmatched']} {'OK' if ok else 'MISMATCH'}")
        if not ok and len(bad):
            i = int(bad[0])
            lo = max(0, i - 2)
            print("   first diff at", i)
            for j in range(lo, min(n, i + 3)):
                print(f"   [{j}] gpu rid={order[j]} f={flag[j]} pos={pos[j]} rc={chr(rev[j])} len={hp.lengths[order[j]]} | "
                      f"orc rid={ro.order[j]} f={ro.flag[j]} pos={ro.pos[j]} rc={chr(ro.rc[j])} len={hp.lengths[ro.order[j]]}")
