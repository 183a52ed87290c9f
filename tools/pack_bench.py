"""GPU box: throughput of spring_b200_pack_reads (preprocess's read path) at 10 M reads x 150 bp, host text in,
packed rows left in HBM; next to the oracle's plain-Python packer on a small sample (for scale only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from spring_b200 import capi, dnaio, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
rs = synth.generate(n, 150, genome_len=n * 5, seed=3, sub_rate=0.005, n_frac=0.002, device="cuda")
codes = rs.codes.cpu().numpy()
bases_np = dnaio.CODE4CHAR[codes].reshape(-1)
del codes
bases = torch.empty(bases_np.shape, dtype=torch.uint8, pin_memory=True); bases.numpy()[:] = bases_np
offs = np.arange(n + 1, dtype=np.uint64) * np.uint64(150)
ctx = capi.Context(0)
ts = []
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = ctx.pack_reads(bases.numpy(), offs, keep_on_device=True)
    ts.append(time.perf_counter() - t0)
ms = 1e3 * min(ts[2:])
print({"reads": n, "ms_incl_h2d_of_text": round(ms, 2), "mreads_s": round(n / ms / 1e3, 1), "text_gb_s": round(bases_np.size / ms / 1e6, 1),
       "num_clean": r["num_clean"], "num_n": r["num_n"], "launches": ctx.stats()["gpu_launches"]})
