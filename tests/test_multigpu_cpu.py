"""N > 1 path on CPU: world_size-2 gloo run of the bucket exchange + per-rank hot path + merge.

The per-rank compute is the oracle here (tests may use it as a stand-in; the product path on a GPU
box is the CUDA library, see bench.py --gpus N and tests/test_gpu_parity.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import make_input, original_reads
from oracle import pyoracle as po
from spring_b200 import dnaio, multigpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M64 = (1 << 64) - 1


def mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64)          # mirror of sb::mix64 (spring_b200/csrc/common.cuh)
    x ^= x >> np.uint64(31); x *= np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(29)
    return x


def bucket_numpy(packed: np.ndarray, lengths: np.ndarray, num_buckets: int) -> np.ndarray:
    """numpy mirror of minimizer_bucket (spring_b200/csrc/common.cuh): canonical 16-mer minimizer bucket."""
    n = len(lengths)
    lmax = int(lengths.max()) if n else 0
    codes = dnaio.unpack_codes(packed, max(lmax, 1)).astype(np.uint64)
    M32 = np.uint64(0xFFFFFFFF)

    def h32(x):
        x = (x * np.uint64(0x9E3779B1)) & M32
        return x ^ (x >> np.uint64(15))
    best = np.full(n, 0xFFFFFFFF, dtype=np.uint64)
    fwd = np.zeros(n, np.uint64); rc = np.zeros(n, np.uint64)
    with np.errstate(over="ignore"):
        for j in range(lmax):
            c = codes[:, j]
            fwd = ((fwd << np.uint64(2)) | c) & M32
            rc = (rc >> np.uint64(2)) | ((np.uint64(3) - c) << np.uint64(30))
            if j >= 15:
                h = h32(np.minimum(fwd, rc))
                upd = (j < lengths) & (h < best)
                best = np.where(upd, h, best)
        short = lengths < 16
        best = np.where(short, h32(lengths.astype(np.uint64)), best)
        return ((mix64(best) >> np.uint64(16)) % np.uint64(num_buckets)).astype(np.int32)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    hp = make_input(num_reads=6000, read_len=100, seed=41)          # every rank regenerates the same set
    n = len(hp.lengths)
    lo, hi = rank * n // world, (rank + 1) * n // world              # this rank's input block
    reads = torch.from_numpy(hp.packed[lo:hi].view(np.int64).copy())
    lens = torch.from_numpy(hp.lengths[lo:hi].view(np.int16).copy())
    ids = torch.arange(lo, hi, dtype=torch.int32)

    def bfn(r, l, w):
        return torch.from_numpy(bucket_numpy(r.numpy().view(np.uint64), l.numpy().view(np.uint16), w))

    r, l, gid = multigpu.exchange_by_bucket(reads, lens, hp.max_readlen, world, ids=ids, bucket_fn=bfn)
    owned = bfn(r, l, world)
    assert (owned == rank).all(), "a read landed on the wrong rank"
    _, er = po.reorder_encode(r.numpy().view(np.uint64), l.numpy().view(np.uint16), hp.max_readlen, num_chains=3)
    q.put((rank, er, gid.numpy().astype(np.uint32)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_multi_rank_exchange_and_merge(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    merged = multigpu.merge_rank_streams([r[1] for r in res], [r[2] for r in res])
    hp = make_input(num_reads=6000, read_len=100, seed=41)
    # all aligned reads of all ranks precede every unaligned read (reorder_compress_streams.cpp:254-270)
    assert merged.num_aligned == sum(r[1].num_aligned for r in res)
    assert (np.sort(merged.order) == np.arange(len(hp.lengths), dtype=np.uint32)).all()
    # decode the merged job exactly as decompress does: concatenated consensus, absolute positions
    class S: pass
    s = S()
    s.seq = np.concatenate([np.asarray(sh[2]) for sh in merged.seq_shards])
    s.pos, s.noise, s.noisepos, s.rc, s.lengths = merged.pos, merged.noise, merged.noisepos, merged.rc, merged.lengths
    s.unaligned, s.num_aligned = merged.unaligned, merged.num_aligned
    dec = po.decode(s)
    orig = original_reads(hp)
    for i, o in enumerate(merged.order):
        assert dec[i] == orig[int(o)]
    # the merged job goes through the stages after the encoder like a single-shard one: re-blocking
    # (reorder_compress_streams.cpp:83-361) and the decompressor's block decode give the reads back
    s.order, s.unaligned_len = merged.order, merged.unaligned_len
    blocks = po.reblock(s, False, False, 1500)
    back = po.decode_blocks(blocks, s.seq, len(hp.lengths), False, False, 1500)
    assert back == [orig[int(o)] for o in merged.order]
    # sharding costs matches but must stay in the same range as one shard
    _, one = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, num_chains=3)
    assert merged.num_aligned > 0.9 * one.num_aligned


def test_bucket_is_strand_canonical_and_shift_tolerant():
    hp = make_input(num_reads=4000, read_len=100, seed=42, sub_rate=0.0)
    b = bucket_numpy(hp.packed, hp.lengths, 8)
    seqs = dnaio.packed_to_seqs(hp.packed, hp.lengths)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    rc_packed, rc_len = dnaio.seqs_to_packed([s.translate(comp)[::-1] for s in seqs], hp.max_readlen)
    assert (bucket_numpy(rc_packed, rc_len, 8) == b).all()            # a read and its reverse complement agree
    assert np.bincount(b, minlength=8).min() > 0.5 * len(b) / 8        # roughly balanced
