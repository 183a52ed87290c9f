"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle."""
import os
import tempfile

import numpy as np
import pytest

from helpers import CASES, assert_streams_equal, check_roundtrip, make_input
from oracle import pyoracle as po
from spring_b200 import capi, dnaio

pytestmark = pytest.mark.gpu

CHAINS = (1, 5, 64)


@pytest.fixture
def det(ctx):
    """ctx with the deterministic (round-synchronous) chain schedule; restored afterwards."""
    ctx.set_schedule(True)
    yield ctx
    ctx.set_schedule(False)


@pytest.mark.parametrize("name", ["se150", "var250", "short40", "long511"])
def test_dictionary_matches_oracle(ctx, name):
    """constructdictionary: same unique keys, same bins, read ids ascending per bin (bit-exact)."""
    hp = make_input(**CASES[name])
    for which in (0, 1):
        k, b, r = ctx.build_dictionary(hp.packed, hp.lengths, hp.max_readlen, which)
        ok, ob, orr = po.reorder_dict(hp.packed, hp.lengths, hp.max_readlen, which)
        assert (k == ok).all() and (b == ob).all() and (r == orr).all()


def test_dictionary_3M_reads_matches_oracle(ctx):
    """At 3 M keys about a thousand pairs of different keys share their top 32 hash bits: the sort's fix-up of
    multi-key runs (dict.cu:k_fix_runs) is exercised on real collisions; keys, bins and ids must still be the oracle's."""
    from spring_b200 import synth
    hp = synth.to_hotpath_input(synth.generate(3_000_000, 100, genome_len=10_000_000, seed=19, sub_rate=0.005, device="cuda"))
    for which in (0, 1):
        k, b, r = ctx.build_dictionary(hp.packed, hp.lengths, hp.max_readlen, which)
        ok, ob, orr = po.reorder_dict(hp.packed, hp.lengths, hp.max_readlen, which)
        assert (k == ok).all() and (b == ob).all() and (r == orr).all()


def test_multi_key_runs_everywhere(tmp_path):
    """SPRING_B200_DICT_SORT_BITS=6: the radix sort looks at 6 hash bits only, so EVERY run holds many keys and the
    fix-up does the real sorting -- dictionaries and streams must not change.  (The knob is read once per process.)"""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        from helpers import CASES, assert_streams_equal, make_input
        from oracle import pyoracle as po
        from spring_b200 import capi
        ctx = capi.Context(0)
        for name in ("se150", "var250", "heavy_bins", "pe100_illumina"):
            hp = make_input(**CASES[name])
            for which in (0, 1):
                k, b, r = ctx.build_dictionary(hp.packed, hp.lengths, hp.max_readlen, which)
                ok, ob, orr = po.reorder_dict(hp.packed, hp.lengths, hp.max_readlen, which)
                assert (k == ok).all() and (b == ob).all() and (r == orr).all(), name
            got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
            _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
            assert_streams_equal(got, er, name)
        print("ok")
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, SPRING_B200_DICT_SORT_BITS="6"))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("name", sorted(CASES))
def test_reorder_stream_matches_oracle(det, name):
    """reorder<>(), deterministic schedule: order / flag / pos / rev / singleton lists bit-exact
    for 1 and many chains."""
    ctx = det
    hp = make_input(**CASES[name])
    for chains in CHAINS:
        order, flag, pos, rev, s_order = ctx.reorder(hp.packed, hp.lengths, hp.max_readlen, chains)
        st = ctx.stats()
        ro = po.reorder(hp.packed, hp.lengths, hp.max_readlen, st["num_chains"])
        assert (order == ro.order).all() and (flag == ro.flag).all() and (pos == ro.pos).all()
        assert (rev == ro.rc).all() and (s_order == ro.s_order).all()
        assert st["unmatched"] == ro.counters["unmatched"]
        assert st["rounds"] == ro.counters["rounds"]
        assert st["probes_seq"] == ro.counters["probes"]      # sequential-equivalent lookups
        assert st["compares"] == ro.counters["compares"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_streams_match_oracle(det, name):
    """call_reorder + call_encoder, deterministic schedule: every output stream bit-exact against
    the oracle for 1 and many chains."""
    ctx = det
    hp = make_input(**CASES[name])
    for chains in CHAINS:
        got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
        c = ctx.stats()["num_chains"]
        _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, c)
        assert_streams_equal(got, er, f"{name} chains={c}")
        assert got.matched_s == er.matched_s and got.matched_N == er.matched_N
        sp, tail = er.packed_seq()
        assert got.seq_packed[: len(sp)].tobytes() == sp


def test_golden_vectors(ctx):
    """Streams produced by the reference itself at -t 1 (tests/golden/make_golden.py)."""
    import json
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    for fn in sorted(f for f in os.listdir(gdir) if f.endswith(".npz") and not f.startswith("reblock_")):
        g = np.load(os.path.join(gdir, fn))
        meta = json.loads(bytes(g["meta"]).decode())
        got = ctx.reorder_encode(g["packed"], g["lengths"], meta["max_readlen"], bytes(g["n_records"]), g["order_n"],
                                 meta["num_reads"], 1)
        for f in ("seq", "pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned"):
            assert (np.asarray(getattr(got, f)) == g["ref_" + f]).all(), f"{fn}: {f}"
        assert got.unaligned_len == meta["unaligned_len"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_free_running_schedule(ctx, name):
    """Default (free-running) schedule.  One chain: bit-exact against the oracle (= the reference at
    -t 1).  Many chains: timing-dependent like the reference at -t > 1, so (a) the reorder output is
    a valid partition of the reads, (b) the encoder is bit-exact against the oracle's encoder fed
    with the very same reorder stream, (c) the streams decode back to the input."""
    hp = make_input(**CASES[name])
    got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    assert_streams_equal(got, er, f"{name} free-running, 1 chain")
    for chains in (7, 0):
        got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
        order, flag, pos, rev, s_order = ctx.fetch_reorder()
        allr = np.concatenate([order, s_order])
        assert (np.sort(allr) == np.arange(len(hp.lengths), dtype=np.uint32)).all()
        assert len(flag) == 0 or flag[0] == 0
        ro = po.ReorderResult(order, flag, pos, rev, s_order, {})
        er = po.encode(hp.packed, hp.lengths, hp.max_readlen, ro, hp.n_records, hp.order_n, hp.num_reads)
        assert_streams_equal(got, er, f"{name} free-running, chains={chains}: encoder vs oracle on the same stream")
        check_roundtrip(got, hp, po.decode)


@pytest.mark.parametrize("name", ["se150", "dups", "pe100_illumina"])
def test_long_contigs_are_cut_like_the_reference(ctx, name, monkeypatch):
    """encoder.h:215: a contig whose read list grows beyond 10 000 000 reads is written in pieces of 10 000 001 stream
    records.  With the limit lowered on both sides (SPRING_B200_CONTIG_SPLIT / SPRING_ORACLE_MAX_LIST = 7) most contigs
    are cut: the streams must still equal the oracle's bit for bit and decode to the input."""
    monkeypatch.setenv("SPRING_B200_CONTIG_SPLIT", "7")
    monkeypatch.setenv("SPRING_ORACLE_MAX_LIST", "7")
    hp = make_input(**CASES[name])
    got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    assert_streams_equal(got, er, f"{name}, contigs cut every 8 reads")
    check_roundtrip(got, hp, po.decode)
    monkeypatch.delenv("SPRING_B200_CONTIG_SPLIT")
    monkeypatch.delenv("SPRING_ORACLE_MAX_LIST")
    whole = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    assert whole.seq_len < got.seq_len  # the cut really happened: pieces repeat consensus the whole contig shares


def test_auto_chains_roundtrip_and_determinism(det):
    """Default chain count (as many as co-reside), deterministic schedule: decode == input, and two
    runs are identical."""
    ctx = det
    hp = make_input(num_reads=200000, read_len=150, seed=21, n_frac=0.002)
    a = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 0)
    st = ctx.stats()
    assert st["num_chains"] > 64
    check_roundtrip(a, hp, po.decode)
    b = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 0)
    assert_streams_equal(a, b, "determinism")
    # match rate stays in the reference's range: oracle with one chain is the reference at -t 1
    ro = po.reorder(hp.packed, hp.lengths, hp.max_readlen, 1)
    assert st["unmatched"] < 1.5 * ro.counters["unmatched"] + st["num_chains"]


def test_edge_cases(ctx):
    # empty input
    got = ctx.reorder_encode(np.zeros((0, 4), np.uint64), np.zeros(0, np.uint16), 100)
    assert len(got.order) == 0 and got.seq_len == 0
    # one read
    p, l = dnaio.seqs_to_packed([b"ACGTACGTAC"], 10)
    got = ctx.reorder_encode(p, l, 10)
    assert list(got.order) == [0] and got.num_aligned == 0 and po.decode(got) == [b"ACGTACGTAC"]
    # only reads with N
    nrec = dnaio.write_dnaN_records([b"ACGNNACGT", b"NNNN"])
    got = ctx.reorder_encode(np.zeros((0, 1), np.uint64), np.zeros(0, np.uint16), 9, nrec, np.array([0, 1], np.uint32), 2)
    assert po.decode(got) == [b"ACGNNACGT", b"NNNN"]
    # identical reads (every shift-0 candidate passes), zero-length read among them
    seqs = [b"ACGTTGCAACGTTGCAACGTTGCAACGTTGCAACGTTGCAACGT"] * 50 + [b""]
    p, l = dnaio.seqs_to_packed(seqs, 44)
    ctx.set_schedule(True)
    got = ctx.reorder_encode(p, l, 44, num_chains=3)
    ctx.set_schedule(False)
    _, er = po.reorder_encode(p, l, 44, num_chains=ctx.stats()["num_chains"])
    assert_streams_equal(got, er, "identical reads")
    # bad arguments are refused with the reference's message
    with pytest.raises(capi.SpringB200Error) as e:
        ctx.reorder_encode(np.zeros((1, 16), np.uint64), np.array([600], np.uint16), 600)
    assert e.value.code == -1 and "Wrong bitset size" in str(e.value)


@pytest.mark.parametrize("name", ["pe100_illumina", "pe_var", "se150"])
def test_file_level_drop_in(ctx, name):
    """spring_b200_reorder_encode_files on a temp_dir laid out by preprocess: same files the
    reference's call_reorder + call_encoder leave, inputs consumed (fixed-length records take the
    sliced multi-threaded copy, mixed lengths the header walk)."""
    hp = make_input(**CASES[name])
    _, er = po.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 1)
    with tempfile.TemporaryDirectory() as d:
        cpy = dnaio.write_hotpath_inputs(d, hp.packed, hp.lengths, max_readlen=hp.max_readlen, n_seqs=hp.n_seqs,
                                         order_n=hp.order_n, num_reads=hp.num_reads,
                                         paired_split=hp.num_clean[0] if hp.paired else None, num_thr=3)
        os.remove(os.path.join(d, "cp_in.bin"))
        cp = capi.CP.from_buffer_copy(cpy.pack())
        ctx.reorder_encode_files(d, cp, 1)
        left = sorted(os.listdir(d))
        assert "input_clean_1.dna" not in left and "input_N.dna" not in left and "read_order_N.bin" not in left
        for t in range(3):
            assert f"read_seq.bin.{t}" in left and f"read_seq.bin.{t}.tail" in left
        got = po.load_reference_streams(d, 3)
    assert_streams_equal(got, er, "file-level")


def _fastq_records(path):
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    return [b"\n".join(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]


@pytest.mark.skipif(not (os.path.exists(po.SPLICE_BIN) and os.path.exists(po.SPLICE2_BIN) and po.have_reference()),
                    reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("splice", ["hotpath", "hotpath+reblock"])
@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("reorder", [True, False])
def test_end_to_end_archive_decodes_with_reference(paired, reorder, splice, tmp_path):
    """The reference's own `spring -c [-r]` host pipeline with call_reorder / call_encoder replaced by
    libspring_b200.so (oracle/_ref/spring_b200_ref), and with pe_encode / reorder_compress_streams
    replaced as well (oracle/_ref/spring_b200_ref2), writes an archive; the UNMODIFIED reference
    (`spring -d`) must decode it to the input: with -r as a multiset of records (pairs kept together,
    the check of util/test_script.sh:78-82), without -r byte for byte (util/test_script.sh:5-21)."""
    import subprocess
    from spring_b200 import synth
    rs = synth.generate(30000, 120, seed=31, paired=paired, n_frac=0.01, var_len=(60, 120), error_model="illumina")
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(rs, f1, f2 if paired else None)
    arc = str(tmp_path / "out.spring")
    ins = [f1, f2] if paired else [f1]
    binary = po.SPLICE_BIN if splice == "hotpath" else po.SPLICE2_BIN
    r = subprocess.run([binary, "-c", *(["-r"] if reorder else []), "-i", *ins, "-o", arc, "-t", "4", "-w", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "were unmatched" in r.stdout and "singleton reads were aligned" in r.stdout and "reads with N were aligned" in r.stdout
    out = str(tmp_path / "dec")
    r = subprocess.run([po.REF_BIN, "-d", "-i", arc, "-o", out, "-t", "3", "-w", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not paired:
        got, want = _fastq_records(out), _fastq_records(f1)
    else:
        got = list(zip(_fastq_records(out + ".1"), _fastq_records(out + ".2")))
        want = list(zip(_fastq_records(f1), _fastq_records(f2)))
    if reorder:
        got, want = sorted(got), sorted(want)
    assert got == want


@pytest.mark.skipif(not (os.path.exists(po.SPLICE3_BIN) and po.have_reference()), reason="oracle/_ref binaries not built")
@pytest.mark.parametrize("paired", [False, True])
@pytest.mark.parametrize("reorder", [True, False])
def test_preprocess_and_decompress_drop_ins(paired, reorder, tmp_path):
    """oracle/_ref/spring_b200_ref3: the reference's host pipeline with preprocess (N split + packing on the GPU, reads
    handed to call_reorder in HBM) and decompress_short (block decode on the GPU) replaced as well.  Three legs:
    (a) its archive decodes with the UNMODIFIED reference, (b) its archive decodes with its own `-d` (GPU decode),
    (c) an archive the unmodified reference wrote (consensus shards of any length, so shards start inside a byte) decodes
    with its `-d`.  With SPRING_B200_PREPROCESS_FILES=1 the drop-in writes the .dna files instead: same archive contents."""
    import subprocess
    from spring_b200 import synth
    rs = synth.generate(30000, 120, seed=33, paired=paired, n_frac=0.01, var_len=(60, 120), error_model="illumina")
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(rs, f1, f2 if paired else None)
    ins = [f1, f2] if paired else [f1]
    flags = ["-r"] if reorder else []

    def run(binary, args, env=None):
        r = subprocess.run([binary, *args, "-w", str(tmp_path)], capture_output=True, text=True, env=dict(os.environ, **(env or {})))
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    def records(prefix):
        if not paired:
            return _fastq_records(prefix)
        return list(zip(_fastq_records(prefix + ".1"), _fastq_records(prefix + ".2")))

    want = _fastq_records(f1) if not paired else list(zip(_fastq_records(f1), _fastq_records(f2)))
    same = (lambda got: sorted(got) == sorted(want)) if reorder else (lambda got: got == want)
    arc3, arc3f, arcr = str(tmp_path / "b200.spring"), str(tmp_path / "b200_files.spring"), str(tmp_path / "ref.spring")
    out = run(po.SPLICE3_BIN, ["-c", *flags, "-i", *ins, "-o", arc3, "-t", "4"])
    assert "were unmatched" in out and "Total number of reads without N" in out
    run(po.SPLICE3_BIN, ["-c", *flags, "-i", *ins, "-o", arc3f, "-t", "4"], {"SPRING_B200_PREPROCESS_FILES": "1"})
    run(po.REF_BIN, ["-c", *flags, "-i", *ins, "-o", arcr, "-t", "5"])
    for name, binary, arc, thr in (("a", po.REF_BIN, arc3, "3"), ("a-files", po.REF_BIN, arc3f, "3"), ("b", po.SPLICE3_BIN, arc3, "3"),
                                   ("c", po.SPLICE3_BIN, arcr, "2")):
        dec = str(tmp_path / ("dec_" + name))
        run(binary, ["-d", "-i", arc, "-o", dec, "-t", thr])
        assert same(records(dec)), f"leg {name}"


@pytest.mark.skipif(not (os.path.exists(po.SPLICE3_BIN) and po.have_reference()), reason="oracle/_ref binaries not built")
def test_decompress_drop_in_over_several_steps(tmp_path):
    """1.1 M paired reads = 550 k pairs = 3 blocks of 256 000 pairs: `-d -t 2` takes two steps (two blocks, then one),
    each ONE spring_b200_decode_blocks call; archive written by the unmodified reference, and by the drop-in binary."""
    import subprocess
    from spring_b200 import synth
    rs = synth.generate(1_100_000, 100, genome_len=4_000_000, seed=35, paired=True, n_frac=0.005, error_model="illumina", device="cuda")
    f1, f2 = str(tmp_path / "in_1.fastq"), str(tmp_path / "in_2.fastq")
    synth.write_fastq(rs, f1, f2)
    want = sorted(zip(_fastq_records(f1), _fastq_records(f2)))
    for who, binary in (("reference", po.REF_BIN), ("drop-in", po.SPLICE3_BIN)):
        arc, dec = str(tmp_path / (who + ".spring")), str(tmp_path / ("dec_" + who))
        r = subprocess.run([binary, "-c", "-r", "-i", f1, f2, "-o", arc, "-t", "8", "-w", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        r = subprocess.run([po.SPLICE3_BIN, "-d", "-i", arc, "-o", dec, "-t", "2", "-w", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        assert sorted(zip(_fastq_records(dec + ".1"), _fastq_records(dec + ".2"))) == want, who


def test_bucket_kernel_matches_numpy_mirror(ctx):
    """k_bucket (multi-GPU owner of a read) against the numpy restatement used by the gloo CPU tests."""
    import torch
    from test_multigpu_cpu import bucket_numpy
    hp = make_input(**CASES["var250"])
    r = torch.from_numpy(hp.packed.view(np.int64).copy()).cuda()
    l = torch.from_numpy(hp.lengths.view(np.int16).copy()).cuda()
    out = torch.empty(len(hp.lengths), dtype=torch.int32, device="cuda")
    for g in (2, 8):
        ctx.bucket_reads(r.data_ptr(), l.data_ptr(), len(hp.lengths), hp.max_readlen, g, out.data_ptr())
        torch.cuda.synchronize()
        assert (out.cpu().numpy() == bucket_numpy(hp.packed, hp.lengths, g)).all()


def _full_size_roundtrip(ctx, rs, chains=0):
    """Size-independent property at BASELINE sizes: decode(streams)[i] == original read order[i],
    order is a permutation, aligned reads first.  Vectorised (no Python loop over reads)."""
    import torch
    from spring_b200 import synth
    hp = synth.to_hotpath_input(rs)
    got = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
    order = np.asarray(got.order).astype(np.int64)
    assert (np.sort(order) == np.arange(hp.num_reads)).all()
    L = rs.max_readlen
    dec = po.decode_matrix(got, L)
    ascii_of = np.frombuffer(b"AGCTN", dtype=np.uint8)
    codes = rs.codes.cpu().numpy()
    lens = rs.lengths.cpu().numpy()
    orig = ascii_of[codes[order]]
    valid = np.arange(L)[None, :] < lens[order][:, None]
    assert (np.asarray(got.lengths) == lens[order]).all()
    assert (np.where(valid, orig, 0) == np.where(valid, dec, 0)).all()
    return got, ctx.stats()


def test_config2_scale_roundtrip(ctx):
    """BASELINE config 2 at 1/5 scale (2 M SE 150 bp, 30x, --no-quality) through the default path."""
    from spring_b200 import synth
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=3, sub_rate=0.005, device="cuda")
    got, st = _full_size_roundtrip(ctx, rs)
    assert got.num_aligned > 0.95 * 2_000_000 and st["num_chains"] > 900


def test_config3_like_roundtrip(ctx):
    """Config 3 shape (paired-end, Illumina error model, 0.2 % reads with N) at 2 M reads."""
    from spring_b200 import synth
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=4, paired=True, n_frac=0.002, error_model="illumina", device="cuda")
    got, st = _full_size_roundtrip(ctx, rs)
    assert got.matched_N > 0


def test_config5_like_roundtrip(ctx):
    """Config 5 shape (variable length 35-250 bp: 512-bit reorder rows) at 1 M reads."""
    from spring_b200 import synth
    rs = synth.generate(1_000_000, 250, genome_len=5_000_000, seed=6, var_len=(35, 250), sub_rate=0.005, device="cuda")
    _full_size_roundtrip(ctx, rs)
