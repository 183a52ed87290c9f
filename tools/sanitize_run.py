"""Small run of every kernel for compute-sanitizer (memcheck / racecheck): the hot path under both chain
schedules, pe_encode + re-blocking (host and device-resident streams), preprocess packing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import CASES, make_input
from spring_b200 import capi, dnaio
ctx = capi.Context(0)
for name in (sys.argv[1:] or ("se100_n", "var64_noisy", "long511", "heavy_bins", "pe100_illumina")):
    kw = dict(CASES[name]); kw["num_reads"] = min(kw["num_reads"], 3000)
    hp = make_input(**kw)
    for det in (False, True):
        ctx.set_schedule(det)
        for chains in (1, 16):
            s = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, chains)
    ctx.set_schedule(False)
    for mode in (1, 0):  # contig stitching on (the stitch kernels and the second layout pass), then off for the re-blocking checks below
        ctx.set_stitch(mode)
        s = ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 16)
    paired = hp.paired
    for preserve, block in ((False, 700), (True, 256000)):
        cp = capi.CP.from_buffer_copy(dnaio.CompressionParams(paired_end=paired, preserve_order=preserve, num_reads=hp.num_reads,
                                                              max_readlen=hp.max_readlen, num_reads_per_block=block).pack())
        b1 = ctx.reblock_streams(cp, None)      # streams resident in HBM
        b2 = ctx.reblock_streams(cp, s)         # host streams
        assert all(b1.data[k].tobytes() == b2.data[k].tobytes() for k in b1.data)
    # round trip in HBM (verify.cu: re-block -> decode.cu -> compare) on the resident streams of the last call
    ctx.reorder_encode(hp.packed, hp.lengths, hp.max_readlen, hp.n_records, hp.order_n, hp.num_reads, 16)
    v = ctx.verify_roundtrip(capi.CP.from_buffer_copy(dnaio.CompressionParams(paired_end=paired, preserve_order=False, num_reads=hp.num_reads,
                                                                               max_readlen=hp.max_readlen, num_reads_per_block=700).pack()))
    assert v["ok"] == 1, v
    seqs = dnaio.packed_to_seqs(hp.packed, hp.lengths) + list(hp.n_seqs)
    offs = np.zeros(len(seqs) + 1, np.uint64); offs[1:] = np.cumsum([len(x) for x in seqs])
    pk = ctx.pack_reads(np.frombuffer(b"".join(seqs), np.uint8), offs)
    assert pk["num_clean"] == len(hp.lengths)
    print(name, "ok", s.num_aligned, b1.num_blocks)
