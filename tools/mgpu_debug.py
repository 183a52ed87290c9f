"""torchrun debug of the multi-GPU exchange: are the received reads intact and on the right rank?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from spring_b200 import capi, multigpu, synth
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rs = synth.generate(n, 150, genome_len=n * world * 5, seed=3, sub_rate=0.005, device=dev, read_seed=3000 + rank)
reads = synth.pack_reads(rs.codes, rs.lengths, 150).contiguous(); lens = rs.lengths.to(torch.int16).contiguous()
ids = (torch.arange(n, device=dev, dtype=torch.int32) + rank * n)
ctx = capi.Context(local, torch.cuda.current_stream().cuda_stream)
bfn = multigpu.gpu_bucket_fn(ctx, 150)
b = bfn(reads, lens, world)
print(f"rank {rank}: send bucket counts {torch.bincount(b.long(), minlength=world).tolist()} lens min/max {int(lens.min())}/{int(lens.max())}", flush=True)
r, l, g = multigpu.exchange_by_bucket(reads, lens, 150, world, ids=ids, bucket_fn=bfn)
torch.cuda.synchronize()
b2 = bfn(r.contiguous(), l.contiguous(), world)
print(f"rank {rank}: received {r.shape[0]} reads; on right rank: {float((b2 == rank).float().mean()):.4f}; lens min/max {int(l.min())}/{int(l.max())}; "
      f"ids from ranks {torch.bincount((g.long() // n), minlength=world).tolist()}", flush=True)
# integrity: checksum of (id, row) pairs must be conserved globally
cs_send = torch.stack([(reads.sum(dim=1) ^ ids.long()).sum(), torch.tensor(n, device=dev)])
cs_recv = torch.stack([(r.sum(dim=1) ^ g.long()).sum(), torch.tensor(r.shape[0], device=dev)])
dist.all_reduce(cs_send); dist.all_reduce(cs_recv)
if rank == 0: print("checksum send/recv", cs_send.tolist(), cs_recv.tolist(), "OK" if cs_send.tolist() == cs_recv.tolist() else "MISMATCH", flush=True)
inp = ctx.make_input(r.data_ptr(), l.data_ptr(), r.shape[0], 150)
s = ctx.reorder_encode_raw(inp, 0, device=True); st = ctx.stats()
print(f"rank {rank}: aligned {s.num_aligned} of {s.num_reads}, unmatched {st['unmatched']}", flush=True)
dist.destroy_process_group()
