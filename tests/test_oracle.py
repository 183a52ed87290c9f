"""CPU tests of the oracle: pinned against the reference build and the committed golden vectors."""
import json
import os
import tempfile

import numpy as np
import pytest

from helpers import CASES, assert_streams_equal, check_roundtrip, make_input
from oracle import pyoracle as po
from spring_b200 import dnaio

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_roundtrip(name):
    hp = make_input(**CASES[name])
    for chains in (1, 7):
        ro, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
        assert len(ro.order) + len(ro.s_order) == len(hp.lengths)
        check_roundtrip(er, hp, po.decode)


@pytest.mark.skipif(not po.have_reference(), reason="oracle/_ref/spring_ref not built (no /root/reference here)")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_equals_reference_single_thread(name):
    """Byte-for-byte equality of every output stream with call_reorder + call_encoder at -t 1."""
    hp = make_input(**CASES[name])
    _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    with tempfile.TemporaryDirectory() as d:
        dnaio.write_hotpath_inputs(d, hp.packed, hp.lengths, max_readlen=hp.max_readlen, n_seqs=hp.n_seqs,
                                   order_n=hp.order_n, num_reads=hp.num_reads,
                                   paired_split=hp.num_clean[0] if hp.paired else None)
        po.run_reference_hotpath(d, 1)
        ref = po.load_reference_streams(d, 1)
    assert_streams_equal(er, ref, name)


def test_oracle_matches_golden_vectors():
    """tests/golden/*.npz were produced by the reference build (tests/golden/make_golden.py)."""
    files = sorted(f for f in os.listdir(GOLDEN) if f.endswith(".npz") and not f.startswith("reblock_"))
    assert files, "no golden vectors committed"
    for fn in files:
        g = np.load(os.path.join(GOLDEN, fn), allow_pickle=False)
        meta = json.loads(bytes(g["meta"]).decode())
        _, er = po.reorder_encode(g["packed"], g["lengths"], meta["max_readlen"], bytes(g["n_records"]), g["order_n"],
                                  meta["num_reads"], 1)
        for f in ("seq", "pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned"):
            assert (np.asarray(getattr(er, f)) == g["ref_" + f]).all(), f"{fn}: {f}"
        assert er.unaligned_len == meta["unaligned_len"]


def test_reference_fixture_has_no_matches():
    """util/test_1.fastq + test_2.fastq (config 1): 39 clean reads, none overlap."""
    g = np.load(os.path.join(GOLDEN, "ref_fixture_pe.npz"), allow_pickle=False)
    meta = json.loads(bytes(g["meta"]).decode())
    ro, er = po.reorder_encode(g["packed"], g["lengths"], meta["max_readlen"], bytes(g["n_records"]), g["order_n"],
                               meta["num_reads"], 1)
    assert len(ro.order) == 0 and er.num_aligned == 0 and len(ro.s_order) == 39


def test_chain_schedule_is_deterministic():
    hp = make_input(**CASES["se100_n"])
    a = po.reorder(hp.packed, hp.lengths, hp.max_readlen, 13)
    b = po.reorder(hp.packed, hp.lengths, hp.max_readlen, 13)
    assert (a.order == b.order).all() and (a.pos == b.pos).all() and (a.s_order == b.s_order).all()


def test_empty_and_tiny_inputs():
    empty = np.zeros((0, 4), np.uint64)
    ro, er = po.reorder_encode(empty, np.zeros(0, np.uint16), 100)
    assert len(er.order) == 0 and len(er.seq) == 0
    p, l = dnaio.seqs_to_packed([b"ACGTACGTAC"], 10)
    ro, er = po.reorder_encode(p, l, 10)
    assert list(er.order) == [0] and er.num_aligned == 0 and po.decode(er) == [b"ACGTACGTAC"]
    # N-only input
    nrec = dnaio.write_dnaN_records([b"ACGNNACGT", b"NNNN"])
    ro, er = po.reorder_encode(np.zeros((0, 1), np.uint64), np.zeros(0, np.uint16), 9, nrec, np.array([0, 1], np.uint32), 2)
    assert po.decode(er) == [b"ACGNNACGT", b"NNNN"]


def test_cut_contigs_still_decode(monkeypatch):
    """encoder.h:215 cuts a contig after 10 000 001 reads; with the oracle's limit lowered the pieces are real and the
    streams must still decode to the input (the GPU test compares the CUDA path with this setting bit for bit)."""
    hp = make_input(**CASES["se150"])
    _, whole = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    monkeypatch.setenv("SPRING_ORACLE_MAX_LIST", "7")
    _, cut = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    assert len(cut.seq) > len(whole.seq)
    check_roundtrip(cut, hp, po.decode)
