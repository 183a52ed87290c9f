"""CPU tests of the boundary: the library loads, exports every declared symbol, and refuses to
compute without a GPU (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from spring_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spring_b200.h")).read()
    return sorted(set(re.findall(r"\b(spring_b200_[a-z_]+)\s*\(", hdr)))


def test_header_symbols_are_exported():
    lib = capi.load()
    syms = declared_symbols()
    assert set(syms) == set(capi.EXPORTS), "capi.EXPORTS out of sync with include/spring_b200.h"
    for s in syms:
        assert hasattr(lib, s), f"libspring_b200.so does not export {s}"


def test_version_and_struct_sizes():
    lib = capi.load()
    assert b"sm_100a" in lib.spring_b200_version()
    assert ctypes.sizeof(capi.CP) == 64          # compression_params, util.h:30-51
    assert capi.CP.num_reads.offset == 28 and capi.CP.max_readlen.offset == 40 and capi.CP.num_thr.offset == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_gpu_fails_loudly():
    with pytest.raises(capi.SpringB200Error) as e:
        capi.Context(0)
    assert e.value.code == -2


def test_ctypes_structs_match_the_c_header(tmp_path):
    """Every ctypes mirror in spring_b200/capi.py has the size and field offsets gcc gives the struct of
    include/spring_b200.h (a silent layout drift would corrupt arguments, not fail)."""
    import subprocess
    pairs = {"spring_b200_cp": capi.CP, "spring_b200_input": capi.Input, "spring_b200_streams": capi.Streams,
             "spring_b200_reorder_out": capi.ReorderOut, "spring_b200_stats": capi.Stats, "spring_b200_blocks": capi.Blocks,
             "spring_b200_packed_reads": capi.PackedReads, "spring_b200_decoded": capi.Decoded, "spring_b200_verify": capi.Verify,
             "spring_b200_exchanged": capi.Exchanged, "spring_b200_shard_layout": capi.ShardLayout, "spring_b200_merged": capi.Merged}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "spring_b200.h"', 'int main(void) {']
    for cname, ct in pairs.items():
        lines.append(f'  printf("{cname} SIZEOF %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)], text=True).split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, field, val = ln.split()
        ct = pairs[cname]
        if field == "SIZEOF":
            assert ctypes.sizeof(ct) == int(val), cname
        else:
            assert getattr(ct, field).offset == int(val), f"{cname}.{field}"
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in pairs.values())


def test_merge_shards_is_the_numpy_merge():
    """spring_b200_merge_shards (host code, no GPU needed) == multigpu.merge_rank_streams on shards finalized the way
    k_finalize_shard does it: positions offset by the lower shards' consensus, local indices mapped to global ids."""
    import numpy as np
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import make_input
    from oracle import pyoracle as po
    from spring_b200 import multigpu
    hp = make_input(num_reads=6000, read_len=100, seed=41, n_frac=0.01)
    n = len(hp.lengths)
    parts, idm = [], []
    for r in range(3):
        lo, hi = r * n // 3, (r + 1) * n // 3
        _, er = po.reorder_encode(hp.packed[lo:hi], hp.lengths[lo:hi], hp.max_readlen, num_chains=2)
        parts.append(er); idm.append(np.arange(lo, hi, dtype=np.uint32))
    want = multigpu.merge_rank_streams(parts, idm)

    class S:
        pass
    fin, base = [], 0
    for er, ids in zip(parts, idm):
        s = S()
        sp, tail = er.packed_seq()
        last = bytes([sum(b"ACGT".index(c) << (2 * j) for j, c in enumerate(tail))]) if tail else b""
        s.seq_packed = np.frombuffer(sp + last, np.uint8); s.seq_len = len(er.seq)
        s.pos = np.asarray(er.pos, np.uint64) + np.uint64(base); s.noise = er.noise; s.noisepos = er.noisepos; s.rc = er.rc
        s.order = ids[np.asarray(er.order, np.int64)]; s.lengths = er.lengths; s.unaligned = er.unaligned
        s.unaligned_len = er.unaligned_len; s.num_aligned = er.num_aligned
        fin.append(s); base += len(er.seq)
    got, shards = capi.merge_shards(fin)
    for f in ("pos", "noise", "noisepos", "rc", "order", "lengths", "unaligned"):
        assert (np.asarray(getattr(got, f)) == np.asarray(getattr(want, f))).all(), f
    assert got.num_aligned == want.num_aligned and got.unaligned_len == want.unaligned_len and got.seq_len == base
    assert [l for _, l in shards] == [len(p.seq) for p in parts]
