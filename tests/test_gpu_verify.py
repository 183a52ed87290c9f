"""spring_b200_verify_roundtrip: re-block -> block decode -> exact compare with the input, all in HBM.

The same check as the reference's own -r test (util/test_script.sh:78-82), made exact through read_order.bin,
and cheap enough for the bench's full sizes.  Here it is (a) cross-checked against the host-side round trip on
small cases, (b) shown to catch a corrupted read, (c) run at BASELINE sizes."""
import numpy as np
import pytest

from helpers import CASES, assert_streams_equal, check_roundtrip, make_input
from oracle import pyoracle as po
from spring_b200 import capi, dnaio

pytestmark = pytest.mark.gpu


def _cp(hp_or_n, paired, max_readlen, block=256000):
    n = hp_or_n if isinstance(hp_or_n, int) else hp_or_n.num_reads
    return capi.CP.from_buffer_copy(dnaio.CompressionParams(paired_end=paired, preserve_order=False, num_reads=n,
                                                            max_readlen=max_readlen, num_reads_per_block=block).pack())


@pytest.mark.parametrize("name", ["se100_n", "var250", "pe100_illumina", "mostly_n", "heavy_bins", "pe_var", "tiny16"])
def test_verify_small_cases(ctx, name):
    hp = make_input(**CASES[name])
    got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 0)
    check_roundtrip(got, hp, po.decode)                       # host-side decode of the same streams
    for block in (256000, 700):
        v = ctx.verify_roundtrip(_cp(hp, hp.paired, hp.max_readlen, block))
        assert v["ok"] == 1 and v["reads_checked"] == hp.num_reads, v
    if hp.paired:                                            # the same streams read as a single-end job decode too
        assert ctx.verify_roundtrip(_cp(hp, False, hp.max_readlen))["ok"] == 1


def test_verify_catches_a_corrupted_read(ctx):
    """The device path compares with the caller's arrays: flip one base of one input read after the
    encode call and exactly that read must be reported."""
    import torch
    from spring_b200 import synth
    rs = synth.generate(50000, 100, seed=9, n_frac=0.01, device="cuda")
    di = synth.to_device_input(rs)
    inp = ctx.make_input(di.reads.data_ptr(), di.lengths.data_ptr(), int(di.reads.shape[0]), 100, di.n_records, di.order_n, di.num_reads)
    ctx.reorder_encode_raw(inp, 0, device=True)
    cp = _cp(di.num_reads, False, 100)
    assert ctx.verify_roundtrip(cp)["ok"] == 1
    di.reads[1234, 0] ^= 1
    torch.cuda.synchronize()
    v = ctx.verify_roundtrip(cp)
    assert v["ok"] == 0 and v["base_mismatch_reads"] == 1 and v["length_mismatch_reads"] == 0 and v["bad_order"] == 0, v
    di.reads[1234, 0] ^= 1
    torch.cuda.synchronize()
    assert ctx.verify_roundtrip(cp)["ok"] == 1


def test_deep_column_does_not_abort(ctx):
    """More than 65 535 reads stacked on one column: the reference counts in int (reorder.h:383-384); the
    packed u16 counts saturate at 65535 instead of failing.  With identical reads every count of a
    column sits in one field, so even the single-chain result stays bit-exact against the oracle."""
    rng = np.random.default_rng(5)
    L = 60
    base = rng.integers(0, 4, L, dtype=np.uint8)
    codes = np.tile(base, (70000, 1))
    other = rng.integers(0, 4, (3000, L), dtype=np.uint8)
    allc = np.concatenate([codes[:40000], other, codes[40000:]])
    lens = np.full(len(allc), L, np.uint16)
    packed = dnaio.pack_codes(allc, lens, L)
    ctx.set_schedule(True)
    try:
        got = ctx.reorder_encode(packed, lens, L, num_chains=1)
    finally:
        ctx.set_schedule(False)
    _, er = po.reorder_encode(packed, lens, L, num_chains=1)
    assert_streams_equal(got, er, "70k identical reads, one chain")
    got = ctx.reorder_encode(packed, lens, L, num_chains=0)       # default schedule, many chains
    assert ctx.verify_roundtrip(_cp(len(lens), False, L))["ok"] == 1
    assert got.num_aligned >= 69990


def _device_job(ctx, rs, paired):
    from spring_b200 import synth
    di = synth.to_device_input(rs)
    inp = ctx.make_input(di.reads.data_ptr(), di.lengths.data_ptr(), int(di.reads.shape[0]), rs.max_readlen, di.n_records,
                         di.order_n, di.num_reads)
    s = ctx.reorder_encode_raw(inp, 0, device=True)
    v = ctx.verify_roundtrip(_cp(di.num_reads, paired, rs.max_readlen))
    return s, v, di


def test_config2_full_size_verified(ctx):
    """BASELINE config 2 at full size (10 M SE 150 bp, 30x): every read of the default path's output decodes
    to its original."""
    from spring_b200 import synth
    rs = synth.generate(10_000_000, 150, genome_len=50_000_000, seed=3, sub_rate=0.005, device="cuda")
    s, v, _ = _device_job(ctx, rs, False)
    assert v["ok"] == 1 and v["reads_checked"] == 10_000_000, v
    assert s.num_aligned > 0.95 * 10_000_000


def test_config3_shape_20M_pairs_verified(ctx):
    """Config 3's shape (paired-end, Illumina error model, 0.2 % reads with N) at 20 M reads, through
    pe_encode + the paired re-blocking + the paired block decode."""
    from spring_b200 import synth
    rs = synth.generate(20_000_000, 150, genome_len=100_000_000, seed=4, paired=True, n_frac=0.002, error_model="illumina", device="cuda")
    s, v, _ = _device_job(ctx, rs, True)
    assert v["ok"] == 1 and v["reads_checked"] == 20_000_000, v
    assert s.n_reads_aligned > 0


def test_config5_shape_10M_verified(ctx):
    """Config 5's shape (35-250 bp: 512-bit reorder rows) at 10 M reads."""
    from spring_b200 import synth
    rs = synth.generate(10_000_000, 250, genome_len=47_500_000, seed=6, var_len=(35, 250), sub_rate=0.005, device="cuda")
    s, v, _ = _device_job(ctx, rs, False)
    assert v["ok"] == 1 and v["reads_checked"] == 10_000_000, v


def test_repeat_rich_10M_verified(ctx):
    """Real genomes are not i.i.d.: interspersed and tandem repeats, low-complexity stretches and coverage spikes
    (5 % of 10 M reads start inside four 200 bp hot spots: ~90 000 reads per column there, dictionary bins of hundreds of reads).  The
    reference counts in int and never fails on such input (reorder.h:383); neither may the GPU path, and every read
    must still decode to its original."""
    from spring_b200 import synth
    rs = synth.generate(10_000_000, 150, genome_len=50_000_000, seed=8, sub_rate=0.005, device="cuda", repeats=True)
    s, v, _ = _device_job(ctx, rs, False)
    assert v["ok"] == 1 and v["reads_checked"] == 10_000_000, v
    st = ctx.stats()
    assert s.num_aligned > 0.9 * 10_000_000, (s.num_aligned, st)


@pytest.mark.parametrize("name", ["se150", "pe100_illumina", "var250", "dups", "lowcov", "se100_n"])
def test_stitched_contigs_decode(ctx, name):
    """Contig stitching (spring_b200_set_stitch): contigs whose head fits into another contig's consensus are laid into it, flipped
    when it fits reversed.  Only positions / orientations change, so every read must still decode to its original -- checked with
    the oracle's decoder on the host and by the round trip in HBM -- and the consensus must not grow."""
    from oracle import pyoracle as po
    from helpers import CASES, check_roundtrip, make_input
    from spring_b200 import dnaio
    hp = make_input(**CASES[name])
    try:
        for chains in (7, 64, 300):
            ctx.set_stitch(0)
            plain = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
            ctx.set_stitch(1)
            got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
            st = ctx.stats()
            check_roundtrip(got, hp, po.decode)
            cp = dnaio.CompressionParams(paired_end=hp.paired, preserve_order=False, num_reads=hp.num_reads, max_readlen=hp.max_readlen,
                                         num_reads_per_block=700)
            v = ctx.verify_roundtrip(capi.CP.from_buffer_copy(cp.pack()))
            assert v["ok"] == 1 and v["reads_checked"] == hp.num_reads, (name, chains, v)
            assert st["contigs_stitched"] <= st["contigs"]
            if name == "se150" and chains == 300:  # 8000 reads over 300 chains: most contigs abut another chain's
                assert st["contigs_stitched"] > 0 and got.seq_len < plain.seq_len, (st["contigs_stitched"], got.seq_len, plain.seq_len)
    finally:
        ctx.set_stitch(0)


def test_stitching_2M_reads_verified(ctx):
    """2 M reads over 4736 chains (422 reads per chain: the regime where the chains' contig starts cost most): stitched, verified in HBM,
    and the consensus shrinks towards the genome's length."""
    from spring_b200 import dnaio, synth
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=41, sub_rate=0.005, device="cuda")
    di = synth.to_device_input(rs)
    n_clean = int(di.reads.shape[0])
    inp = ctx.make_input(di.reads.data_ptr(), di.lengths.data_ptr(), n_clean, 150, di.n_records, di.order_n, rs.num_reads)
    cp = capi.CP.from_buffer_copy(dnaio.CompressionParams(paired_end=False, preserve_order=False, num_reads=rs.num_reads, max_readlen=150).pack())
    try:
        lens = {}
        for mode in (0, 1):
            ctx.set_stitch(mode)
            s = ctx.reorder_encode_raw(inp, 4736, device=True)
            v = ctx.verify_roundtrip(cp)
            assert v["ok"] == 1 and v["reads_checked"] == rs.num_reads, (mode, v)
            lens[mode] = (int(s.seq_len), ctx.stats()["contigs_stitched"])
        assert lens[1][1] > 0 and lens[1][0] < lens[0][0], lens
    finally:
        ctx.set_stitch(0)
