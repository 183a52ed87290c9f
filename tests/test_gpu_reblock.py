"""GPU parity tests of the stages after the encoder: spring_b200_pe_encode and
spring_b200_reblock_streams / _files, called through the C ABI, against oracle/reblock_oracle.c
(pinned against the reference by tests/test_reblock_oracle.py) and against golden vectors made by
the reference itself (tests/golden/make_golden_reblock.py).  Bit-exact: byte streams and offsets."""
import json
import os
import tempfile
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import CASES, make_input
from oracle import pyoracle as po
from spring_b200 import capi, dnaio
from test_reblock_oracle import FAMILIES, assert_blocks_equal, encoder_streams

pytestmark = pytest.mark.gpu


def make_cp(num_reads, max_readlen, paired, preserve, block) -> capi.CP:
    cp = dnaio.CompressionParams(paired_end=paired, preserve_order=preserve, num_reads=num_reads, max_readlen=max_readlen,
                                 num_reads_per_block=block, num_thr=1)
    return capi.CP.from_buffer_copy(cp.pack())


@pytest.mark.parametrize("n", [0, 2, 10, 2000, 1_000_000])
def test_pe_encode_matches_oracle(ctx, n):
    order = np.random.default_rng(n).permutation(n).astype(np.uint32)
    assert (ctx.pe_encode(order) == po.pe_encode(order)).all()
    if n:
        assert ctx.stats()["gpu_launches"] > 0


def test_pe_encode_refuses_bad_input(ctx):
    with pytest.raises(capi.SpringB200Error):
        ctx.pe_encode(np.array([0, 1, 2], np.uint32))       # odd
    with pytest.raises(capi.SpringB200Error):
        ctx.pe_encode(np.array([0, 7], np.uint32))          # not a permutation


@pytest.mark.parametrize("name,paired,preserve,block", FAMILIES)
def test_reblock_matches_oracle(ctx, name, paired, preserve, block):
    """Host streams (the oracle's encoder output) through spring_b200_reblock_streams."""
    hp, er = encoder_streams(name)
    order = po.pe_encode(er.order) if paired and not preserve else er.order
    want = po.reblock(er, paired, preserve, block, order=order)
    got = ctx.reblock_streams(make_cp(hp.num_reads, hp.max_readlen, paired, preserve, block), er)
    assert_blocks_equal(got, want, f"{name} paired={paired} preserve={preserve} block={block}")
    if paired or preserve:
        assert (got.order == order).all()
    else:
        assert got.order is None
    assert ctx.stats()["gpu_launches"] > 0


def test_reblock_golden_vectors(ctx):
    """Block files written by the reference's own pe_encode + reorder_compress_streams."""
    gdir = os.path.join(os.path.dirname(__file__), "golden")
    names = sorted(f for f in os.listdir(gdir) if f.startswith("reblock_") and f.endswith(".npz"))
    assert names
    for fn in names:
        g = np.load(os.path.join(gdir, fn))
        meta = json.loads(bytes(g["meta"]).decode())
        er = SimpleNamespace(pos=g["pos"], noise=g["noise"], noisepos=g["noisepos"], rc=g["rc"], order=g["order"],
                             lengths=g["lengths"], unaligned=g["unaligned"], unaligned_len=meta["unaligned_len"],
                             num_aligned=meta["num_aligned"])
        got = ctx.reblock_streams(make_cp(meta["num_reads"], 511, meta["paired"], meta["preserve"], meta["block"]), er)
        for s in po.BLOCK_STREAMS:
            assert got.data[s].tobytes() == g["blk_" + s].tobytes(), f"{fn}: {s}"
            assert (got.off[s] == g["off_" + s]).all(), f"{fn}: offsets of {s}"


@pytest.mark.parametrize("paired", [False, True])
def test_reblock_device_resident_full_size(ctx, paired):
    """BASELINE shapes at 2 M reads: the encoder's streams stay in HBM (streams=None) and are re-blocked
    there; bit-exact against the oracle run on the fetched streams, default block size 256000."""
    from spring_b200 import synth
    rs = synth.generate(2_000_000, 150, genome_len=10_000_000, seed=8, paired=paired, n_frac=0.002,
                        error_model="illumina" if paired else "uniform", device="cuda")
    hp = synth.to_hotpath_input(rs)
    st = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 0)
    got = ctx.reblock_streams(make_cp(hp.num_reads, hp.max_readlen, paired, False, 256000), None)
    assert ctx.stats()["ms_reblock"] > 0
    order = po.pe_encode(st.order) if paired else st.order
    want = po.reblock(st, paired, False, 256000, order=order)
    assert_blocks_equal(got, want, f"device-resident, paired={paired}")
    units = hp.num_reads // 2 if paired else hp.num_reads
    assert got.num_blocks == (units + 255999) // 256000
    if paired:
        assert (got.order == order).all()
        flags = np.bincount(got.data["flag"] - ord("0"), minlength=5)
        assert flags[0] > 0.5 * units      # most pairs: both mates aligned within 32767 of each other


def test_reblock_edge_cases(ctx):
    # no reads at all
    empty = SimpleNamespace(pos=np.zeros(0, np.uint64), noise=np.zeros(0, np.uint8), noisepos=np.zeros(0, np.uint16),
                            rc=np.zeros(0, np.uint8), order=np.zeros(0, np.uint32), lengths=np.zeros(0, np.uint16),
                            unaligned=np.zeros(0, np.uint8), unaligned_len=0, num_aligned=0)
    got = ctx.reblock_streams(make_cp(0, 100, False, False, 256000), empty)
    assert got.num_blocks == 0 and all(len(v) == 0 for v in got.data.values())
    # only unaligned reads (nothing matched): flags all '2', text = the reads
    hp = make_input(**CASES["lowcov"])
    _, er = po.reorder_encode(hp.packed[:50], hp.lengths[:50], hp.max_readlen, b"", None, 50, 1)
    got = ctx.reblock_streams(make_cp(50, hp.max_readlen, False, False, 7), er)
    assert_blocks_equal(got, po.reblock(er, False, False, 7), "unaligned only")
    # inconsistent streams are refused, not mis-blocked
    hp, er = encoder_streams("se150")
    bad = SimpleNamespace(**{k: getattr(er, k) for k in ("pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned",
                                                          "unaligned_len", "num_aligned")})
    bad.noise = er.noise.copy(); bad.noise[bad.noise == ord("\n")] = ord("0")
    with pytest.raises(capi.SpringB200Error):
        ctx.reblock_streams(make_cp(hp.num_reads, hp.max_readlen, False, False, 1000), bad)
    with pytest.raises(capi.SpringB200Error):
        ctx.reblock_streams(make_cp(hp.num_reads + 1, hp.max_readlen, False, False, 1000), er)


def test_reblock_files_drop_in(ctx):
    """spring_b200_reblock_files on a temp_dir laid out by the encoder: inputs consumed, raw block files
    identical to what the reference writes before BSC."""
    hp, er = encoder_streams("pe100_illumina")
    block = 900
    cpd = dnaio.CompressionParams(paired_end=True, preserve_order=False, num_reads=hp.num_reads, max_readlen=hp.max_readlen,
                                  num_reads_per_block=block, num_thr=1)
    want = po.reblock(er, True, False, block, order=po.pe_encode(er.order))
    with tempfile.TemporaryDirectory() as d:
        po.write_encoder_streams(d, er, cpd.pack())
        os.remove(os.path.join(d, "cp_in.bin"))
        ctx.reblock_files(d, capi.CP.from_buffer_copy(cpd.pack()))
        left = set(os.listdir(d))
        assert not ({"read_pos.bin", "read_order.bin", "read_noise.txt", "read_unaligned.txt.count"} & left)
        for s, fn in zip(po.BLOCK_STREAMS, po.BLOCK_FILES):
            for b in range(want.num_blocks):
                with open(os.path.join(d, f"{fn}.{b}"), "rb") as f:
                    assert f.read() == want.block(s, b), f"{fn}.{b}"
