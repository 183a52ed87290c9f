"""Multi-GPU plumbing of the hot path: one process per GPU, one exchange.

The reference is single-process (OpenMP only).  Its on-disk format is
already sharded -- decompress concatenates ``read_seq.bin.<t>`` for t < cp.num_thr and all
positions are absolute in that concatenation (src/decompress.cpp:106-120, src/encoder.h:473-487) --
so a GPU plays the role of one reference thread:

  1. every rank holds a block of the clean reads (global id = rank offset + local index)
  2. owner(read) = minimizer bucket mod world (CUDA kernel, csrc/bucket.cu)
  3. ONE all-to-all(v) of fused {packed read, length[, global id]} byte records over NCCL / NVLink
     (preceded by the tiny all-to-all of the split sizes)
  4. each rank runs the unchanged single-GPU reorder + encode on the reads it owns
  5. the per-rank streams are merged on the host: consensus shards concatenated, positions offset,
     local indices mapped back to global ids, and -- the invariant the downstream stages rely on
     (src/reorder_compress_streams.cpp:254-270) -- all aligned reads of all ranks before any
     unaligned read.

`exchange_by_bucket` is torch.distributed only (works on NCCL and, for the CPU tests, on gloo);
the bucket function is injected so the tests can run it without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def gpu_bucket_fn(ctx, max_readlen: int):
    """bucket function backed by the library's CUDA kernel (device tensors in, device tensor out)."""
    def fn(reads: torch.Tensor, lens: torch.Tensor, world: int) -> torch.Tensor:
        out = torch.empty((reads.shape[0],), dtype=torch.int32, device=reads.device)
        ctx.bucket_reads(reads.data_ptr(), lens.data_ptr(), reads.shape[0], max_readlen, world, out.data_ptr())
        return out
    return fn


_default_ctx = {}


def exchange_by_bucket(reads: torch.Tensor, lens: torch.Tensor, max_readlen: int, world: int,
                       ids: torch.Tensor | None = None, bucket_fn=None, group=None):
    """reads int64[n, W], lens int16[n] (device or CPU) -> the reads this rank owns after the
    all-to-all, as (reads, lens) or (reads, lens, ids) when global ids are passed."""
    if world == 1:
        return (reads, lens) if ids is None else (reads, lens, ids)
    if bucket_fn is None:
        from . import capi
        key = reads.device.index
        if key not in _default_ctx:
            _default_ctx[key] = capi.Context(key, torch.cuda.current_stream().cuda_stream)
        bucket_fn = gpu_bucket_fn(_default_ctx[key], max_readlen)
    bucket = bucket_fn(reads, lens, world)
    # stable sort of one-byte keys (a single radix pass) instead of an int64 argsort: rank order = input order
    if world > 256:
        raise ValueError("exchange_by_bucket: one-byte owner keys, world <= 256")
    order = torch.argsort(bucket.to(torch.uint8), stable=True)
    send_counts = torch.bincount(bucket.long(), minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rc = send_counts.tolist(), recv_counts.tolist()
    n_recv = int(sum(rc))

    # ONE all-to-all for everything a read carries: the 8W-byte row, its u16 length and (optionally) its
    # u32 global id travel as one fused byte record (every backend moves uint8; gloo has no int16 all-to-all)
    fields = [reads, lens] + ([ids] if ids is not None else [])
    n = reads.shape[0]
    widths = [int(x.element_size() * int(np.prod(x.shape[1:], dtype=np.int64))) for x in fields]
    if n:  # explicit widths: reshape(n, -1) cannot infer a width for n == 0 (a rank without reads must still join the collective)
        rec = torch.cat([x[order].contiguous().view(torch.uint8).reshape(n, w) for x, w in zip(fields, widths)], dim=1)
    else:
        rec = torch.empty((0, sum(widths)), dtype=torch.uint8, device=reads.device)
    got = torch.empty((n_recv, sum(widths)), dtype=torch.uint8, device=reads.device)
    dist.all_to_all_single(got, rec, output_split_sizes=rc, input_split_sizes=sc, group=group)
    out, off = [], 0
    for x, w in zip(fields, widths):
        out.append(got[:, off:off + w].contiguous().view(x.dtype).reshape((n_recv,) + tuple(x.shape[1:])))
        off += w
    return tuple(out)


@dataclass
class MergedStreams:
    """Whole-job streams in the reference's layout: one consensus shard per rank."""
    seq_shards: list            # per rank: (packed bytes uint8, seq_len)
    pos: np.ndarray
    noise: np.ndarray
    noisepos: np.ndarray
    rc: np.ndarray
    order: np.ndarray
    lengths: np.ndarray
    unaligned: np.ndarray
    unaligned_len: int
    num_aligned: int


def merge_rank_streams(parts: list, id_maps: list) -> MergedStreams:
    """parts[r]: that rank's streams (fields seq_packed/seq_len or seq, pos, noise, noisepos, rc, order,
    lengths, unaligned, unaligned_len, num_aligned); id_maps[r]: uint32 global id of local read i.
    What the reference's thread-file merge does (src/encoder.h:386-487), with ranks as threads."""
    pos, noise, noisepos, rc, order_a, len_a, order_u, len_u, unal, shards = [], [], [], [], [], [], [], [], [], []
    base = 0
    ul = 0
    for s, ids in zip(parts, id_maps):
        na = int(s.num_aligned)
        pos.append(np.asarray(s.pos, dtype=np.uint64) + np.uint64(base))
        noise.append(np.asarray(s.noise)); noisepos.append(np.asarray(s.noisepos)); rc.append(np.asarray(s.rc))
        o = np.asarray(ids, dtype=np.uint32)[np.asarray(s.order, dtype=np.int64)]
        order_a.append(o[:na]); order_u.append(o[na:])
        ln = np.asarray(s.lengths)
        len_a.append(ln[:na]); len_u.append(ln[na:])
        unal.append(np.asarray(s.unaligned))
        seq_len = int(s.seq_len) if hasattr(s, "seq_len") else len(s.seq)
        shards.append((np.asarray(s.seq_packed) if hasattr(s, "seq_packed") else None, seq_len, getattr(s, "seq", None)))
        base += seq_len
        ul += int(s.unaligned_len)
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return MergedStreams(shards, cat(pos, np.uint64), cat(noise, np.uint8), cat(noisepos, np.uint16), cat(rc, np.uint8),
                         cat(order_a + order_u, np.uint32), cat(len_a + len_u, np.uint16), cat(unal, np.uint8), ul,
                         int(sum(int(s.num_aligned) for s in parts)))


def init_comm(ctx, rank: int, world: int, device=None) -> None:
    """Create the library's NCCL communicator on every rank: rank 0 makes the unique id, torch.distributed (any
    backend) broadcasts its 128 bytes."""
    from . import capi
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    ctx.comm_init(bytes(buf.cpu().numpy().tobytes()), rank, world)


def global_ids(n_local: int, rank: int, world: int, paired: bool, device) -> torch.Tensor:
    """Global index (position in the whole job's FASTQ order, file 2 after file 1) of local read i when every rank holds
    one block of n_local reads -- for paired input a block of n_local / 2 pairs, file-1 mates first."""
    i = torch.arange(n_local, device=device, dtype=torch.int64)
    if not paired:
        return (rank * n_local + i).to(torch.int32)
    half_l, half_g = n_local // 2, n_local * world // 2
    return torch.where(i < half_l, rank * half_l + i, half_g + rank * half_l + (i - half_l)).to(torch.int32)
