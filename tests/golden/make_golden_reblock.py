"""Golden vectors for pe_encode + the re-blocking of reorder_compress_streams, made by the REFERENCE itself
(oracle/_ref/spring_ref --hotpath for the encoder streams, then --reblock: src/pe_encode.cpp and
src/reorder_compress_streams.cpp unmodified, block files BSC-decoded).  Run in the build container:
    python tests/golden/make_golden_reblock.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_input  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from spring_b200 import dnaio  # noqa: E402

CASES = {
    "reblock_pe": (dict(num_reads=3000, read_len=75, seed=41, paired=True, n_frac=0.02, error_model="illumina"), True, False, 400),
    "reblock_se": (dict(num_reads=2500, read_len=100, seed=42, n_frac=0.02, var_len=(40, 100)), False, False, 600),
    "reblock_se_ordered": (dict(num_reads=2000, read_len=60, seed=43, n_frac=0.01), False, True, 256000),
}

for name, (kw, paired, preserve, block) in CASES.items():
    hp = make_input(**kw)
    with tempfile.TemporaryDirectory() as d:
        dnaio.write_hotpath_inputs(d, hp.packed, hp.lengths, max_readlen=hp.max_readlen, n_seqs=hp.n_seqs, order_n=hp.order_n,
                                   num_reads=hp.num_reads, paired_split=hp.num_clean[0] if paired else None, num_thr=1)
        po.run_reference_hotpath(d, 1, unbsc=True)
        er = po.load_reference_streams(d, 1)
        cp = dnaio.CompressionParams(paired_end=paired, preserve_order=preserve, num_reads=hp.num_reads, max_readlen=hp.max_readlen,
                                     num_reads_per_block=block, num_thr=1)
        for f in os.listdir(d):
            if f.startswith("read_seq"):
                os.remove(os.path.join(d, f))
        with open(os.path.join(d, "cp_in.bin"), "wb") as f:
            f.write(cp.pack())
        units = hp.num_reads // 2 if paired else hp.num_reads
        ref = po.run_reference_reblock(d, paired, (units + block - 1) // block, num_thr=2)
    out = {"meta": np.frombuffer(json.dumps(dict(paired=paired, preserve=preserve, block=block, num_reads=hp.num_reads,
                                                 num_aligned=int(er.num_aligned), unaligned_len=int(er.unaligned_len))).encode(), np.uint8),
           "pos": er.pos, "noise": er.noise, "noisepos": er.noisepos, "rc": er.rc, "order": er.order, "lengths": er.lengths,
           "unaligned": er.unaligned}
    for s in po.BLOCK_STREAMS:
        out["blk_" + s] = ref.data[s]
        out["off_" + s] = ref.off[s]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), name + ".npz"), **out)
    print(name, hp.num_reads, "reads,", ref.num_blocks, "blocks")
