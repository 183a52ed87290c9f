"""Pins oracle/reblock_oracle.c:orc_decode_blocks (decompress_short's block decode)
against the reference decompressor: an archive written by the unmodified reference (`spring -c`) is opened,
its block files BSC-decoded with the reference's own BSC_decompress, decoded by the oracle -- and must give
exactly the reads `spring -d` writes, in the same order."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from spring_b200 import dnaio, synth

needs_ref = pytest.mark.skipif(not po.have_reference(), reason="oracle/_ref/spring_ref not built")


def seq_lines(path):
    lines = open(path, "rb").read().split(b"\n")
    return [lines[i] for i in range(1, len(lines) - 1, 4)]


@needs_ref
@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("reorder", [True, False])
def test_decode_oracle_equals_reference_decompressor(paired, reorder, tmp_path):
    rs = synth.generate(9000, 110, seed=71 + paired, paired=paired, n_frac=0.02, var_len=(30, 110), error_model="illumina")
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(rs, f1, f2 if paired else None)
    arc = str(tmp_path / "a.spring")
    ins = [f1, f2] if paired else [f1]
    r = subprocess.run([po.REF_BIN, "-c", *(["-r"] if reorder else []), "-i", *ins, "-o", arc, "-t", "2", "-w", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = str(tmp_path / "dec")
    r = subprocess.run([po.REF_BIN, "-d", "-i", arc, "-o", out, "-t", "2", "-w", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    want = seq_lines(out + ".1") + seq_lines(out + ".2") if paired else seq_lines(out)
    work = str(tmp_path / "x"); os.makedirs(work)
    cpb, blocks, seq = po.load_archive_blocks(arc, work)
    cp = dnaio.CompressionParams.unpack(cpb)
    assert cp.paired_end == paired and cp.preserve_order == (not reorder)
    got = po.decode_blocks(blocks, seq, cp.num_reads, cp.paired_end, cp.preserve_order, cp.num_reads_per_block)
    assert got == want
    if not reorder:  # order-preserving mode gives the input back
        assert got == (seq_lines(f1) + (seq_lines(f2) if paired else []))


def test_decode_inverts_reblock():
    """decode(reblock(encoder streams)) == the reads the encoder streams stand for, slot by slot."""
    from helpers import CASES, make_input, original_reads
    for name, paired, preserve, block in (("pe100_illumina", True, False, 700), ("se100_n", False, False, 1000), ("se150", False, True, 3000)):
        hp = make_input(**CASES[name])
        _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
        slot = po.pe_encode(er.order) if paired and not preserve else (er.order if (paired or preserve) else np.arange(hp.num_reads))
        blocks = po.reblock(er, paired, preserve, block, order=slot if (paired or preserve) else None)
        got = po.decode_blocks(blocks, er.seq, hp.num_reads, paired, preserve, block)
        stream_reads = po.decode(er)                       # per stream index
        want = [None] * hp.num_reads
        for i, s in enumerate(np.asarray(slot)):
            want[int(s)] = stream_reads[i]
        assert got == want, name
        if preserve:
            assert got == original_reads(hp)
