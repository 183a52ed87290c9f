"""GPU box: per-phase cycle breakdown of the chain kernel on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spring_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
chains = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
genome = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else n * 150 // 30
dev = torch.device("cuda", 0)
var = "var" in sys.argv  # config 5's shape: 35-250 bp reads, 46 % of them end a contig (contig-start dominated)
L = 250 if var else 150
if var and len(sys.argv) <= 3:
    genome = int(n * 142.5 / 30)
rs = synth.generate(n, L, genome_len=genome, seed=6 if var else 3, sub_rate=0.005, device=dev, var_len=(35, 250) if var else None)
d_reads = synth.pack_reads(rs.codes, rs.lengths, L).contiguous(); d_lens = rs.lengths.to(torch.int16).contiguous(); del rs
ctx = capi.Context(0, torch.cuda.current_stream().cuda_stream)
inp = ctx.make_input(d_reads.data_ptr(), d_lens.data_ptr(), n, L)
for it in range(3):
    ctx.reorder_encode_raw(inp, chains, device=True)
    st = ctx.stats()
tot = st["rounds"] * ((st["num_chains"] + 7) // 8)
print({k: st[k] for k in ("num_chains", "rounds", "unmatched", "lost_proposals", "probes_issued", "probes_seq", "slot_probes", "compares")})
print({k: round(st[k], 3) for k in st if k.startswith("ms_")})
print("avg cycles per block-round: search %.0f  waitA %.0f  commit %.0f  waitB %.0f  | round %.0f cycles; kernel us/round %.2f" % (
    st["cyc_search"] / tot, st["cyc_wait_a"] / tot, st["cyc_commit"] / tot, st["cyc_wait_b"] / tot,
    (st["cyc_search"] + st["cyc_wait_a"] + st["cyc_commit"] + st["cyc_wait_b"]) / tot, 1e3 * st["ms_chain_kernel"] / st["rounds"]))
