"""Shared helpers of the parity tests."""
from __future__ import annotations

import numpy as np

from spring_b200 import dnaio, synth

STREAM_FIELDS = ("seq", "pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned")

# (name, generate kwargs) -- small cases the oracle finishes in well under a second each
CASES = {
    "se100_n": dict(num_reads=6000, read_len=100, seed=1, n_frac=0.01),
    "se150": dict(num_reads=8000, read_len=150, seed=2, n_frac=0.002),
    "var250": dict(num_reads=5000, read_len=250, seed=3, var_len=(35, 250), n_frac=0.005),
    "pe100_illumina": dict(num_reads=6000, read_len=100, seed=4, paired=True, n_frac=0.01, error_model="illumina"),
    "short40": dict(num_reads=3000, read_len=40, seed=5, n_frac=0.02),
    "var64_noisy": dict(num_reads=3000, read_len=64, seed=6, var_len=(1, 64), sub_rate=0.03, n_frac=0.05),
    "long511": dict(num_reads=1500, read_len=511, seed=7, var_len=(300, 511), genome_len=20000),
    "heavy_bins": dict(num_reads=30000, read_len=40, seed=8, genome_len=60, sub_rate=0.01),
    "lowcov": dict(num_reads=4000, read_len=100, seed=9, genome_len=4000000),
    # window-formula boundaries: reorder.h:752-759 switches at L = 100, encoder.h:610-620 at L = 50
    "len50": dict(num_reads=2500, read_len=50, seed=10, n_frac=0.02),
    "len51": dict(num_reads=2500, read_len=51, seed=11, n_frac=0.02),
    "len100": dict(num_reads=3000, read_len=100, seed=12, sub_rate=0.01),
    "len101": dict(num_reads=3000, read_len=101, seed=13, sub_rate=0.01),
    "fixed12": dict(num_reads=1500, read_len=12, seed=14, genome_len=400),
    "tiny16": dict(num_reads=1500, read_len=16, seed=14, genome_len=4000, var_len=(3, 16)),
    "dups": dict(num_reads=4000, read_len=80, seed=15, genome_len=500, sub_rate=0.0),
    "mostly_n": dict(num_reads=1500, read_len=90, seed=16, n_frac=0.9),
    "pe_var": dict(num_reads=4000, read_len=140, seed=17, paired=True, var_len=(50, 140), n_frac=0.02, error_model="illumina"),
}


def make_input(**kw) -> synth.HotpathInput:
    return synth.to_hotpath_input(synth.generate(**kw))


def original_reads(hp: synth.HotpathInput) -> list[bytes]:
    """Reads in original FASTQ order (clean + N merged back), as ASCII."""
    clean = dnaio.packed_to_seqs(hp.packed, hp.lengths)
    out = [None] * hp.num_reads
    is_n = np.zeros(hp.num_reads, dtype=bool)
    is_n[hp.order_n] = True
    for i, s in zip(hp.order_n, hp.n_seqs):
        out[int(i)] = s
    it = iter(clean)
    for i in range(hp.num_reads):
        if not is_n[i]:
            out[i] = next(it)
    return out


def assert_streams_equal(a, b, what=""):
    for f in STREAM_FIELDS:
        x, y = np.asarray(getattr(a, f)), np.asarray(getattr(b, f))
        assert x.shape == y.shape, f"{what}: stream {f}: {x.shape} vs {y.shape}"
        if not (x == y).all():
            i = int(np.nonzero(x != y)[0][0])
            raise AssertionError(f"{what}: stream {f} differs first at {i}: {x[max(0,i-3):i+4]} vs {y[max(0,i-3):i+4]}")
    assert a.unaligned_len == b.unaligned_len, what
    assert a.num_aligned == b.num_aligned, what


def check_roundtrip(er, hp: synth.HotpathInput, decode):
    """decode(streams)[i] must be original read order[i]; order must be a permutation."""
    dec = decode(er)
    orig = original_reads(hp)
    order = np.asarray(er.order)
    assert len(dec) == hp.num_reads
    assert (np.sort(order) == np.arange(hp.num_reads, dtype=np.uint32)).all(), "order is not a permutation"
    for i, o in enumerate(order):
        assert dec[i] == orig[int(o)], f"read {i} (original {o}) does not round-trip"
