"""Pins oracle/reblock_oracle.c (pe_encode + the re-blocking of reorder_compress_streams)
byte for byte against the reference itself: oracle/_ref/spring_ref --reblock runs the unmodified
src/pe_encode.cpp and src/reorder_compress_streams.cpp on the same encoder streams and BSC-decodes the
per-block files it wrote."""
import os
import tempfile

import numpy as np
import pytest

from helpers import CASES, make_input
from oracle import pyoracle as po
from spring_b200 import dnaio

needs_ref = pytest.mark.skipif(not po.have_reference(), reason="oracle/_ref/spring_ref not built")

# (case, paired, preserve_order, reads per block)
FAMILIES = [
    ("se100_n", False, False, 256000),
    ("se100_n", False, False, 1000),       # many blocks: first-read-of-block absolute positions
    ("se150", False, True, 3000),          # order-preserving mode: absolute u64 positions, order applied
    ("pe100_illumina", True, False, 700),  # pe_encode + pair flags 0..4
    ("pe100_illumina", True, True, 256000),
    ("pe_var", True, False, 512),
    ("mostly_n", False, False, 400),       # mostly unaligned reads
    ("lowcov", False, False, 999),         # mostly singletons
    ("var64_noisy", False, False, 777),
]


def encoder_streams(name):
    hp = make_input(**CASES[name])
    _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    return hp, er


def assert_blocks_equal(a: po.BlockStreams, b: po.BlockStreams, what=""):
    assert a.num_blocks == b.num_blocks, what
    for s in po.BLOCK_STREAMS:
        assert (np.asarray(a.off[s], np.uint64) == np.asarray(b.off[s], np.uint64)).all(), f"{what}: block offsets of {s}"
        assert a.data[s].tobytes() == b.data[s].tobytes(), f"{what}: stream {s}"


@needs_ref
@pytest.mark.parametrize("name,paired,preserve,block", FAMILIES)
def test_reblock_oracle_equals_reference(name, paired, preserve, block):
    hp, er = encoder_streams(name)
    order = er.order
    if paired and not preserve:
        order = po.pe_encode(er.order)
    got = po.reblock(er, paired, preserve, block, order=order)
    cp = dnaio.CompressionParams(paired_end=paired, preserve_order=preserve, num_reads=hp.num_reads, max_readlen=hp.max_readlen,
                                 num_reads_per_block=block, num_thr=2)
    with tempfile.TemporaryDirectory() as d:
        po.write_encoder_streams(d, er, cp.pack())
        units = hp.num_reads // 2 if paired else hp.num_reads
        ref = po.run_reference_reblock(d, paired, (units + block - 1) // block)
        left = set(os.listdir(d))
        assert "read_order.bin" not in left and "read_pos.bin" not in left  # consumed, as the reference does
    assert_blocks_equal(got, ref, f"{name} paired={paired} preserve={preserve} block={block}")
    if paired:
        assert set(got.data["flag"].tolist()) <= set(b"01234")


def test_pe_encode_properties():
    """pe_encode: file-1 reads keep their stream order, every file-2 read gets its mate's slot + half."""
    rng = np.random.default_rng(5)
    n = 2000
    order = rng.permutation(n).astype(np.uint32)
    new = po.pe_encode(order)
    half = n // 2
    f1 = order < half
    assert (new[f1] == np.arange(half)).all()
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    for i in np.nonzero(~f1)[0][:200]:
        assert new[i] == new[inv[order[i] - half]] + half
    assert (np.sort(new) == np.arange(n)).all()


def test_reblock_empty_and_tiny():
    from dataclasses import replace
    hp, er = encoder_streams("short40")
    b = po.reblock(er, False, False, 10 ** 9)
    assert b.num_blocks == 1 and len(b.data["flag"]) == hp.num_reads and len(b.data["lengths"]) == 2 * hp.num_reads
    # stream sizes are conserved by the re-blocking
    assert len(b.data["noisepos"]) == 2 * len(er.noisepos) and len(b.data["noise"]) == len(er.noise)
    assert len(b.data["unaligned"]) == er.unaligned_len and len(b.data["rc"]) == er.num_aligned


def test_reblock_oracle_matches_golden_vectors():
    """tests/golden/reblock_*.npz: block files written by the reference (make_golden_reblock.py); this
    check also runs where oracle/_ref is absent."""
    import json
    from types import SimpleNamespace
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    names = sorted(f for f in os.listdir(gdir) if f.startswith("reblock_") and f.endswith(".npz"))
    assert len(names) >= 3
    for fn in names:
        g = np.load(os.path.join(gdir, fn))
        meta = json.loads(bytes(g["meta"]).decode())
        er = SimpleNamespace(pos=g["pos"], noise=g["noise"], noisepos=g["noisepos"], rc=g["rc"], order=g["order"],
                             lengths=g["lengths"], unaligned=g["unaligned"], unaligned_len=meta["unaligned_len"],
                             num_aligned=meta["num_aligned"])
        order = po.pe_encode(er.order) if meta["paired"] and not meta["preserve"] else er.order
        got = po.reblock(er, meta["paired"], meta["preserve"], meta["block"], order=order)
        for s in po.BLOCK_STREAMS:
            assert got.data[s].tobytes() == g["blk_" + s].tobytes(), f"{fn}: {s}"
            assert (got.off[s] == g["off_" + s]).all(), f"{fn}: offsets of {s}"
