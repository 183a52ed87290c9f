"""Pins oracle.pyoracle.pack_reads (preprocess's read path: N split + 2-bit / 4-bit record packing,
SURVEY 8f rank 2) against the reference itself: oracle/_ref/spring_ref --preprocess runs the unmodified
src/preprocess.cpp on FASTQ files and the files it leaves are compared byte for byte."""
import os
import tempfile

import numpy as np
import pytest

from oracle import pyoracle as po
from spring_b200 import dnaio, synth

needs_ref = pytest.mark.skipif(not po.have_reference(), reason="oracle/_ref/spring_ref not built")

SETS = {
    "se_n": dict(num_reads=3000, read_len=100, seed=51, n_frac=0.05),
    "pe_var": dict(num_reads=4000, read_len=151, seed=52, paired=True, var_len=(20, 151), n_frac=0.03, error_model="illumina"),
    "mostly_n": dict(num_reads=800, read_len=37, seed=53, n_frac=0.8),
    "long511": dict(num_reads=300, read_len=511, seed=54, var_len=(400, 511), genome_len=20000, n_frac=0.01),
    "tiny": dict(num_reads=500, read_len=9, seed=55, var_len=(1, 9), genome_len=300, n_frac=0.1),
}


def read_seqs(rs):
    codes, lens = rs.codes.cpu().numpy(), rs.lengths.cpu().numpy()
    return [dnaio.CODE4CHAR[codes[i, : lens[i]]].tobytes() for i in range(rs.num_reads)]


def assert_same(a: dict, b: dict, what=""):
    for k in ("clean_1", "clean_2", "n_records"):
        assert a[k] == b[k], f"{what}: {k}"
    assert (np.asarray(a["order_n"]) == np.asarray(b["order_n"])).all(), what
    assert tuple(a["num_reads_clean"]) == tuple(b["num_reads_clean"]) and a["max_readlen"] == b["max_readlen"], what
    assert a["num_reads"] == b["num_reads"], what


@needs_ref
@pytest.mark.parametrize("name", sorted(SETS))
def test_pack_oracle_equals_reference_preprocess(name, tmp_path):
    rs = synth.generate(**SETS[name])
    f1, f2 = str(tmp_path / "a_1.fastq"), str(tmp_path / "a_2.fastq")
    synth.write_fastq(rs, f1, f2 if rs.paired else None)
    d = str(tmp_path / "tmp"); os.makedirs(d)
    ref = po.run_reference_preprocess(d, f1, f2 if rs.paired else None)
    half = rs.num_reads // 2 if rs.paired else rs.num_reads
    got = po.pack_reads(read_seqs(rs), half)
    assert_same(got, ref, name)


@needs_ref
@pytest.mark.skipif(not os.path.exists("/root/reference/util/test_1.fastq"), reason="reference fixtures not present")
def test_pack_oracle_on_the_reference_fixture(tmp_path):
    """util/test_1.fastq + test_2.fastq: lengths 0-100, 161 of 200 reads with N."""
    def seqs_of(p):
        lines = open(p, "rb").read().split(b"\n")
        return [lines[i] for i in range(1, len(lines) - 1, 4)]
    f1, f2 = "/root/reference/util/test_1.fastq", "/root/reference/util/test_2.fastq"
    d = str(tmp_path / "tmp"); os.makedirs(d)
    ref = po.run_reference_preprocess(d, f1, f2)
    s1, s2 = seqs_of(f1), seqs_of(f2)
    assert_same(po.pack_reads(s1 + s2, len(s1)), ref, "reference fixture")


def test_pack_oracle_agrees_with_the_hotpath_layout():
    """The file images parse back (readDnaFile, reorder.h:222-244) to the arrays the tests feed the hot path."""
    rs = synth.generate(**SETS["pe_var"])
    hp = synth.to_hotpath_input(rs)
    o = po.pack_reads(read_seqs(rs), rs.num_reads // 2)
    with tempfile.TemporaryDirectory() as d:
        for nm, k in (("c1", "clean_1"), ("c2", "clean_2")):
            open(os.path.join(d, nm), "wb").write(o[k])
        p1, l1 = dnaio.read_dna_file(os.path.join(d, "c1"), o["num_reads_clean"][0], o["max_readlen"])
        p2, l2 = dnaio.read_dna_file(os.path.join(d, "c2"), o["num_reads_clean"][1], o["max_readlen"])
    assert (np.concatenate([p1, p2]) == hp.packed).all() and (np.concatenate([l1, l2]) == hp.lengths).all()
    assert o["n_records"] == hp.n_records and (o["order_n"] == hp.order_n).all()
