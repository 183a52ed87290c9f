"""GPU parity tests of spring_b200_pack_reads (preprocess's read path) against the
oracle's file images, which tests/test_pack_oracle.py pins against the reference's own preprocess."""
import os
import tempfile

import numpy as np
import pytest

from helpers import assert_streams_equal
from oracle import pyoracle as po
from spring_b200 import capi, dnaio, synth
from test_pack_oracle import SETS, read_seqs

pytestmark = pytest.mark.gpu


def flat(seqs):
    offs = np.zeros(len(seqs) + 1, np.uint64)
    offs[1:] = np.cumsum([len(s) for s in seqs])
    return np.frombuffer(b"".join(seqs), dtype=np.uint8), offs


def file_image(packed, lengths):
    """write_dna_in_bits records of packed rows (util.cpp:269-294)."""
    if len(lengths) == 0:
        return b""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.dna")
        dnaio.write_dna_file(p, packed, lengths)
        return open(p, "rb").read()


@pytest.mark.parametrize("name", sorted(SETS))
def test_pack_reads_matches_oracle(ctx, name):
    rs = synth.generate(**SETS[name])
    seqs = read_seqs(rs)
    half = rs.num_reads // 2 if rs.paired else rs.num_reads
    want = po.pack_reads(seqs, half)
    bases, offs = flat(seqs)
    got = ctx.pack_reads(bases, offs, half)
    assert (got["num_clean_file1"], got["num_clean"] - got["num_clean_file1"]) == tuple(want["num_reads_clean"])
    assert got["max_readlen"] == want["max_readlen"] and got["num_reads"] == want["num_reads"]
    c0 = got["num_clean_file1"]
    assert file_image(got["packed"][:c0], got["lengths"][:c0]) == want["clean_1"]
    assert file_image(got["packed"][c0:], got["lengths"][c0:]) == want["clean_2"]
    assert got["n_records"] == want["n_records"] and (got["order_n"] == want["order_n"]).all()
    assert ctx.stats()["gpu_launches"] > 0


def test_pack_reads_edge_cases(ctx):
    # no reads; only empty reads; a single base
    got = ctx.pack_reads(np.zeros(0, np.uint8), np.zeros(1, np.uint64))
    assert got["num_clean"] == 0 and got["num_n"] == 0 and got["max_readlen"] == 0
    bases, offs = flat([b"", b"", b"N", b"", b"G"])
    got = ctx.pack_reads(bases, offs)
    want = po.pack_reads([b"", b"", b"N", b"", b"G"])
    assert got["num_clean"] == 4 and list(got["lengths"]) == [0, 0, 0, 1] and got["packed"][3, 0] == 1
    assert got["n_records"] == want["n_records"] and list(got["order_n"]) == [2]
    # characters the reference's tables do not define are refused, as is a read beyond MAX_READ_LEN
    for bad in (b"ACGTacgt", b"ACGR", b"AC.GT"):
        with pytest.raises(capi.SpringB200Error) as e:
            ctx.pack_reads(*flat([b"ACGT", bad]))
        assert e.value.code == -1
    with pytest.raises(capi.SpringB200Error) as e:
        ctx.pack_reads(*flat([b"A" * 512]))
    assert "Too long read length" in str(e.value)
    assert ctx.pack_reads(*flat([b"A" * 511]))["max_readlen"] == 511


def test_fastq_bases_to_streams_without_dna_files(ctx):
    """pack_reads(keep_on_device) -> reorder_encode_device: the same streams as the host-array path."""
    rs = synth.generate(60000, 150, seed=61, paired=True, n_frac=0.01, error_model="illumina")
    hp = synth.to_hotpath_input(rs)
    seqs = read_seqs(rs)
    bases, offs = flat(seqs)
    ctx.set_schedule(True)
    try:
        want = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 64)
        pk = ctx.pack_reads(bases, offs, rs.num_reads // 2, keep_on_device=True)
        assert pk["max_readlen"] == hp.max_readlen and pk["num_clean"] == len(hp.lengths)
        inp = ctx.make_input(pk["reads_ptr"], pk["lengths_ptr"], pk["num_clean"], pk["max_readlen"], pk["n_records"],
                             pk["order_n"], pk["num_reads"])
        ctx.reorder_encode_raw(inp, 64, device=True)
        got = ctx.fetch_streams()
    finally:
        ctx.set_schedule(False)
    assert_streams_equal(got, want, "bases -> pack on the GPU -> hot path")


def test_pack_reads_full_size(ctx):
    """Config-2 shape at 2 M reads: the GPU-packed rows equal the generator's own packing."""
    import torch
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=3, sub_rate=0.005, n_frac=0.002, device="cuda")
    hp = synth.to_hotpath_input(rs)
    codes = rs.codes.cpu().numpy()
    assert (rs.lengths == 150).all()
    bases = dnaio.CODE4CHAR[codes].reshape(-1)
    offs = (np.arange(rs.num_reads + 1, dtype=np.uint64) * np.uint64(150))
    got = ctx.pack_reads(bases, offs)
    assert (got["packed"] == hp.packed).all() and (got["lengths"] == hp.lengths).all()
    assert got["n_records"] == hp.n_records and (got["order_n"] == hp.order_n).all()
