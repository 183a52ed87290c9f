"""N > 1 path on real GPUs (skipped with fewer than 2): the library's own exchange (fused bucket / scatter kernels +
one NCCL group), the single-GPU path per rank, spring_b200_finalize_shard, and the C++ merge -- then the MERGED job is
decoded and compared with the input, pairs kept together (the reference's own -r check, util/test_script.sh:78-82).

One process per GPU (torch.multiprocessing spawn, NCCL rendezvous on 127.0.0.1), like bench.py under torchrun."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_READS, READ_LEN, SEED = 60000, 100, 51


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _job(paired):
    from spring_b200 import synth
    return synth.generate(N_READS, READ_LEN, seed=SEED, n_frac=0.01, paired=paired, error_model="illumina", device="cuda")


def _local_block(rs, rank, world):
    """This rank's block of the job: single end a block of reads, paired end a block of pairs (file-1 mates, then their
    file-2 mates), as bench.py generates them."""
    import torch
    from spring_b200 import synth
    n = rs.num_reads
    if rs.paired:
        half, hl = n // 2, n // 2 // world
        idx = torch.cat([torch.arange(rank * hl, (rank + 1) * hl), half + torch.arange(rank * hl, (rank + 1) * hl)]).to(rs.codes.device)
    else:
        nl = n // world
        idx = torch.arange(rank * nl, (rank + 1) * nl, device=rs.codes.device)
    return synth.ReadSet(rs.codes[idx], rs.lengths[idx], rs.max_readlen, rs.paired), idx


def _worker(rank, world, port, paired, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from spring_b200 import capi, multigpu, synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rs = _job(paired)                                  # every rank regenerates the same job (same seed)
    local, idx = _local_block(rs, rank, world)
    di = synth.to_device_input(local)
    n_local = local.num_reads
    gid = multigpu.global_ids(n_local, rank, world, paired, dev)
    assert (gid.long() == idx.long()).all()            # global id == position in the whole job's FASTQ order
    isn = torch.zeros(n_local, dtype=torch.bool, device=dev)
    if len(di.order_n):
        isn[torch.from_numpy(di.order_n.astype(np.int64)).to(dev)] = True
    ids_clean = gid[~isn].contiguous()
    ids_n = gid[isn].cpu().numpy().astype(np.uint32)
    ctx = capi.Context(rank, torch.cuda.current_stream().cuda_stream)
    multigpu.init_comm(ctx, rank, world, dev)
    n_clean = int(di.reads.shape[0])
    for rep in range(2):                               # twice: buffers are reused
        x = ctx.exchange_reads(di.reads.data_ptr(), di.lengths.data_ptr(), ids_clean.data_ptr(), n_clean, READ_LEN)
        n_own = int(x.num_reads)
        on = (n_own + np.arange(len(ids_n))).astype(np.uint32)
        inp = ctx.make_input(x.reads, x.lengths, n_own, READ_LEN, di.n_records, on, n_own + len(on))
        ctx.reorder_encode_raw(inp, 0, device=True)
        lay = ctx.finalize_shard(x.ids, n_own, ids_n)
    s = ctx.fetch_streams()
    torch.cuda.synchronize()
    st = ctx.stats()
    q.put((rank, s, lay, st["ms_exchange"], n_own, int(x.sent_to_peers), int(x.received_from_peers)))
    dist.barrier()
    ctx.comm_free()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("paired", [False, True])
def test_merged_multi_gpu_job_decodes_to_the_input(ctx, paired):
    import torch
    import torch.multiprocessing as mp
    from oracle import pyoracle as po
    from spring_b200 import capi, dnaio
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    if world == 3:
        world = 2
    port = _free_port()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, port, paired, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    parts = [r[1] for r in res]
    lays = [r[2] for r in res]
    merged, shards = capi.merge_shards(parts)
    # layout bookkeeping of finalize_shard == what the merge did
    assert lays[0]["total_reads"] == N_READS == len(merged.order)
    assert lays[-1]["total_aligned"] == merged.num_aligned == sum(int(p.num_aligned) for p in parts)
    assert sum(r[5] for r in res) == sum(r[6] for r in res) > 0          # reads really crossed between the GPUs
    assert sum(r[4] for r in res) + sum(len(p.order) - r[4] for p, r in zip(parts, res)) == N_READS
    order = np.asarray(merged.order).astype(np.int64)
    assert (np.sort(order) == np.arange(N_READS)).all(), "global ids are not a permutation"
    # decode the merged job as decompress does: concatenated consensus, absolute positions (decompress.cpp:106-120)
    rs = _job(paired)
    codes, lens = rs.codes.cpu().numpy(), rs.lengths.cpu().numpy()
    orig = [dnaio.CODE4CHAR[codes[i, : lens[i]]].tobytes() for i in range(N_READS)]

    class S:
        pass
    s = S()
    s.seq = np.concatenate([p.seq for p in parts])
    s.pos, s.noise, s.noisepos, s.rc, s.lengths = merged.pos, merged.noise, merged.noisepos, merged.rc, merged.lengths
    s.unaligned, s.num_aligned, s.order, s.unaligned_len = merged.unaligned, merged.num_aligned, merged.order, merged.unaligned_len
    dec = po.decode(s)
    assert all(dec[i] == orig[int(o)] for i, o in enumerate(order)), "merged streams do not decode to the input"
    # ... and through the stages after the encoder on the GPU: pe_encode + re-blocking, then the block decode
    cp = capi.CP.from_buffer_copy(dnaio.CompressionParams(paired_end=paired, preserve_order=False, num_reads=N_READS,
                                                          max_readlen=READ_LEN, num_reads_per_block=7000).pack())
    blocks = ctx.reblock_streams(cp, merged)
    acgt = np.zeros(256, np.uint8); acgt[list(b"ACGT")] = [0, 1, 2, 3]
    c2 = acgt[s.seq]
    pad = (-len(c2)) % 4
    c2 = np.concatenate([c2, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    seq_packed = (c2[:, 0] | (c2[:, 1] << 2) | (c2[:, 2] << 4) | (c2[:, 3] << 6)).astype(np.uint8)
    bases, offs = ctx.decode_blocks(blocks, seq_packed, len(s.seq), cp)
    got = [bases[int(offs[i]): int(offs[i + 1])].tobytes() for i in range(N_READS)]
    if paired:
        half = N_READS // 2
        assert sorted(zip(got[:half], got[half:])) == sorted(zip(orig[:half], orig[half:]))
    else:
        assert sorted(got) == sorted(orig)
    # sharding costs matches but stays in the range of one GPU
    assert merged.num_aligned > 0.85 * N_READS


@pytest.mark.parametrize("paired", [False, True])
def test_spliced_reference_binary_on_two_gpus(paired, tmp_path):
    """The reference's own `spring -c -r` host pipeline with call_reorder / call_encoder from libspring_b200.so and
    SPRING_B200_GPUS=2 (one process, two GPUs: spring_b200_reorder_encode_files_multi): the archive must decode with the
    UNMODIFIED reference to the input, pairs kept together (util/test_script.sh:78-82)."""
    import subprocess
    import torch
    from oracle import pyoracle as po
    from spring_b200 import synth
    from test_gpu_parity import _fastq_records
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    if not (os.path.exists(po.SPLICE_BIN) and po.have_reference()):
        pytest.skip("oracle/_ref binaries not built")
    rs = synth.generate(40000, 120, seed=33, paired=paired, n_frac=0.01, var_len=(60, 120), error_model="illumina")
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(rs, f1, f2 if paired else None)
    arc = str(tmp_path / "out.spring")
    ins = [f1, f2] if paired else [f1]
    env = dict(os.environ, SPRING_B200_GPUS="2")
    r = subprocess.run([po.SPLICE_BIN, "-c", "-r", "-i", *ins, "-o", arc, "-t", "4", "-w", str(tmp_path)], capture_output=True, text=True,
                       env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "were unmatched" in r.stdout and "singleton reads were aligned" in r.stdout
    out = str(tmp_path / "dec")
    r = subprocess.run([po.REF_BIN, "-d", "-i", arc, "-o", out, "-t", "3", "-w", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    if not paired:
        got, want = sorted(_fastq_records(out)), sorted(_fastq_records(f1))
    else:
        got = sorted(zip(_fastq_records(out + ".1"), _fastq_records(out + ".2")))
        want = sorted(zip(_fastq_records(f1), _fastq_records(f2)))
    assert got == want
