"""GPU parity tests of spring_b200_decode_blocks (decompress_short's block decode) against
oracle/reblock_oracle.c:orc_decode_blocks, which tests/test_decode_oracle.py pins against the reference
decompressor; and the device round trip FASTQ bases -> pack -> reorder + encode -> re-block -> decode."""
import numpy as np
import pytest

from helpers import CASES, make_input, original_reads
from oracle import pyoracle as po
from spring_b200 import capi, dnaio, synth
from test_gpu_reblock import make_cp
from test_reblock_oracle import FAMILIES, encoder_streams

pytestmark = pytest.mark.gpu


def split(bases, offs):
    return [bases[int(offs[i]): int(offs[i + 1])].tobytes() for i in range(len(offs) - 1)]


def packed_seq(er):
    """read_seq.bin's coding of the whole consensus (encoder.cpp:126-141), the tail packed as well."""
    code = np.zeros(256, np.uint8); code[list(b"ACGT")] = [0, 1, 2, 3]
    c = code[np.asarray(er.seq)]
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8), len(er.seq)


@pytest.mark.parametrize("name,paired,preserve,block", FAMILIES)
def test_decode_blocks_matches_oracle(ctx, name, paired, preserve, block):
    hp, er = encoder_streams(name)
    order = po.pe_encode(er.order) if paired and not preserve else er.order
    blocks = po.reblock(er, paired, preserve, block, order=order)
    want = po.decode_blocks(blocks, er.seq, hp.num_reads, paired, preserve, block)
    sp, sl = packed_seq(er)
    bases, offs = ctx.decode_blocks(blocks, sp, sl, make_cp(hp.num_reads, hp.max_readlen, paired, preserve, block))
    assert split(bases, offs) == want
    assert ctx.stats()["gpu_launches"] > 0


def test_decode_blocks_with_position_escapes(ctx):
    """Deltas >= 65535 between consecutive aligned reads (65535 + absolute u64), several in one 32-unit group,
    next to independently coded mates: the one sequential stream of the decoder."""
    rng = np.random.default_rng(9)
    n_units, L = 3000, 50
    seq = rng.integers(0, 4, size=3_000_000).astype(np.uint8)
    from types import SimpleNamespace
    for paired in (False, True):
        n = n_units * (2 if paired else 1)
        jump = rng.random(n) < 0.2
        pos = np.cumsum(np.where(jump, rng.integers(65535, 200000, n), rng.integers(0, 300, n))).astype(np.uint64)
        pos = pos % np.uint64(len(seq) - L)          # not monotone: backward jumps wrap and take the escape too
        er = SimpleNamespace(seq=np.frombuffer(b"ACGT", np.uint8)[seq], pos=pos, noise=np.full(n, ord("\n"), np.uint8),
                             noisepos=np.zeros(0, np.uint16), rc=rng.choice(np.frombuffer(b"dr", np.uint8), n),
                             order=rng.permutation(n).astype(np.uint32), lengths=np.full(n, L, np.uint16),
                             unaligned=np.zeros(0, np.uint8), unaligned_len=0, num_aligned=n)
        order = po.pe_encode(er.order) if paired else None
        blocks = po.reblock(er, paired, False, 1000, order=order)
        assert len(blocks.data["pos"]) > 2 * n_units + 8 * 300                                 # escapes are present
        want = po.decode_blocks(blocks, er.seq, n, paired, False, 1000)
        sp, sl = packed_seq(er)
        bases, offs = ctx.decode_blocks(blocks, sp, sl, make_cp(n, L, paired, False, 1000))
        assert split(bases, offs) == want


def test_decode_refuses_inconsistent_streams(ctx):
    hp, er = encoder_streams("se150")
    blocks = po.reblock(er, False, False, 1000)
    sp, sl = packed_seq(er)
    cp = make_cp(hp.num_reads, hp.max_readlen, False, False, 1000)
    from copy import deepcopy
    bad = deepcopy(blocks); bad.data["flag"] = blocks.data["flag"].copy(); bad.data["flag"][5] = ord("2") if blocks.data["flag"][5] == ord("0") else ord("0")
    with pytest.raises(capi.SpringB200Error):
        ctx.decode_blocks(bad, sp, sl, cp)
    with pytest.raises(capi.SpringB200Error):
        ctx.decode_blocks(blocks, sp[: len(sp) // 2], sl // 2, cp)   # positions beyond the consensus
    with pytest.raises(capi.SpringB200Error):
        ctx.decode_blocks(blocks, sp, sl, make_cp(hp.num_reads, hp.max_readlen, False, False, 999))


@pytest.mark.parametrize("paired", [False, True])
def test_device_round_trip_full_size(ctx, paired):
    """2 M reads: bases -> pack_reads -> reorder + encode (free-running, all chains) -> re-block (streams stay
    in HBM) -> decode_blocks gives the input back as a multiset (pairs kept together) -- the reference's own
    -r check (util/test_script.sh:78-82), every stage of it on the GPU."""
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=12, paired=paired, n_frac=0.002,
                        error_model="illumina" if paired else "uniform", device="cuda")
    codes, lens = rs.codes.cpu().numpy(), rs.lengths.cpu().numpy()
    assert (lens == 150).all()
    bases_in = dnaio.CODE4CHAR[codes].reshape(-1)
    offs_in = np.arange(rs.num_reads + 1, dtype=np.uint64) * np.uint64(150)
    half = rs.num_reads // 2 if paired else rs.num_reads
    pk = ctx.pack_reads(bases_in, offs_in, half, keep_on_device=True)
    inp = ctx.make_input(pk["reads_ptr"], pk["lengths_ptr"], pk["num_clean"], pk["max_readlen"], pk["n_records"], pk["order_n"], pk["num_reads"])
    ctx.reorder_encode_raw(inp, 0, device=True)
    st = ctx.fetch_streams()
    cp = make_cp(rs.num_reads, 150, paired, False, 256000)
    blocks = ctx.reblock_streams(cp, None)
    bases, offs = ctx.decode_blocks(blocks, st.seq_packed, st.seq_len, cp)
    assert (np.diff(offs.astype(np.int64)) == 150).all()
    got = bases.reshape(-1, 150)
    want = bases_in.reshape(-1, 150)
    if paired:
        got = np.concatenate([got[:half], got[half:]], axis=1)
        want = np.concatenate([want[:half], want[half:]], axis=1)
    key = lambda m: m[np.lexsort(m.T[::-1])]
    assert (key(got) == key(want)).all()
