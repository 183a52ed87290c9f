"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/spring_ref, built from
/root/reference by oracle/Makefile).  Run in the build container only; the vectors are committed
because /root/reference does not exist on the GPU box.

Each file holds the hot path's inputs (packed clean reads, lengths, input_N.dna bytes,
read_order_N.bin) and every stream call_reorder + call_encoder left behind at -t 1.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_input  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from spring_b200 import dnaio  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = ("seq", "pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned")


def save(name, packed, lengths, n_records, order_n, max_readlen, num_reads, ref, extra=None):
    meta = dict(max_readlen=int(max_readlen), num_reads=int(num_reads), unaligned_len=int(ref.unaligned_len), **(extra or {}))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), packed=packed, lengths=lengths,
                        n_records=np.frombuffer(n_records, dtype=np.uint8), order_n=order_n,
                        meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8),
                        **{"ref_" + f: getattr(ref, f) for f in FIELDS})


def from_synthetic(name, **kw):
    hp = make_input(**kw)
    with tempfile.TemporaryDirectory() as d:
        dnaio.write_hotpath_inputs(d, hp.packed, hp.lengths, max_readlen=hp.max_readlen, n_seqs=hp.n_seqs, order_n=hp.order_n,
                                   num_reads=hp.num_reads, paired_split=hp.num_clean[0] if hp.paired else None)
        po.run_reference_hotpath(d, 1)
        ref = po.load_reference_streams(d, 1)
    save(name, hp.packed, hp.lengths, hp.n_records, hp.order_n, hp.max_readlen, hp.num_reads, ref, dict(gen=kw))


def from_reference_fixture():
    """BASELINE config 1: util/test_1.fastq + util/test_2.fastq through the reference's own preprocess."""
    u = "/root/reference/util"
    with tempfile.TemporaryDirectory() as d:
        subprocess.check_call([po.REF_BIN, "--preprocess", "-i", f"{u}/test_1.fastq", f"{u}/test_2.fastq", "--temp", d, "-r", "-t", "1"],
                              stdout=subprocess.DEVNULL)
        cp = dnaio.CompressionParams.unpack(open(os.path.join(d, "cp_in.bin"), "rb").read())
        p1, l1 = dnaio.read_dna_file(os.path.join(d, "input_clean_1.dna"), cp.num_reads_clean_0, cp.max_readlen)
        p2, l2 = dnaio.read_dna_file(os.path.join(d, "input_clean_2.dna"), cp.num_reads_clean_1, cp.max_readlen)
        packed, lengths = np.concatenate([p1, p2]), np.concatenate([l1, l2])
        nrec = open(os.path.join(d, "input_N.dna"), "rb").read()
        order_n = np.fromfile(os.path.join(d, "read_order_N.bin"), dtype=np.uint32)
        for f in os.listdir(d):  # -r mode leaves raw id/quality text for later stages; not part of the hot path
            if f.startswith(("id_", "quality_")):
                os.remove(os.path.join(d, f))
        po.run_reference_hotpath(d, 1)
        ref = po.load_reference_streams(d, 1)
    save("ref_fixture_pe", packed, lengths, nrec, order_n, cp.max_readlen, cp.num_reads, ref,
         dict(num_clean=[cp.num_reads_clean_0, cp.num_reads_clean_1]))


if __name__ == "__main__":
    po.build()
    from_reference_fixture()
    from_synthetic("syn_se100", num_reads=4000, read_len=100, seed=11, n_frac=0.01)
    from_synthetic("syn_var150", num_reads=3000, read_len=150, seed=12, var_len=(30, 150), n_frac=0.01)
    from_synthetic("syn_pe75", num_reads=3000, read_len=75, seed=13, paired=True, n_frac=0.01, error_model="illumina")
    print(sorted(os.listdir(HERE)))
