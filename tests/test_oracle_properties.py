"""Property tests of the oracle's restatements of the stages around the hot path (hypothesis): they are the
checkers of the GPU kernels, so their own invariants are worth pinning beyond the reference-made vectors."""
from types import SimpleNamespace

import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import pyoracle as po
from spring_b200 import dnaio


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 300), st.integers(0, 2 ** 31))
def test_pe_encode_is_a_pairing_permutation(half, seed):
    n = 2 * half
    order = np.random.default_rng(seed).permutation(n).astype(np.uint32)
    new = po.pe_encode(order)
    assert sorted(new.tolist()) == list(range(n))
    f1 = order < half
    assert (new[f1] == np.arange(half)).all()                       # file-1 reads keep the stream order
    inv = np.empty(n, np.int64); inv[order] = np.arange(n)
    for i in np.nonzero(~f1)[0]:
        assert new[i] == new[inv[order[i] - half]] + half          # every mate follows at + n/2


def random_streams(rng, n, L, seq_len, frac_aligned):
    """Encoder-style streams with arbitrary (not necessarily sorted) positions: exercises the delta / escape coding."""
    na = int(n * frac_aligned)
    lens = rng.integers(0, L + 1, n).astype(np.uint16)
    pos = rng.integers(0, max(seq_len - L, 1), na).astype(np.uint64)
    if rng.random() < 0.5:
        pos = np.sort(pos)
    noise, noisepos = bytearray(), []
    for i in range(na):
        k = int(rng.integers(0, 3)) if lens[i] else 0
        ps = np.sort(rng.choice(int(lens[i]), size=min(k, int(lens[i])), replace=False)) if lens[i] else []
        prev = 0
        for p in ps:
            noise.append(ord("0") + int(rng.integers(0, 3)))          # '3' (-> N) only for N reads in real streams
            noisepos.append(int(p) - prev); prev = int(p)
        noise.append(ord("\n"))
    seqs = ["".join(rng.choice(list("ACGTN"), int(l))).encode() for l in lens[na:]]
    return SimpleNamespace(seq=rng.choice(np.frombuffer(b"ACGT", np.uint8), seq_len), pos=pos, noise=np.frombuffer(bytes(noise), np.uint8),
                           noisepos=np.array(noisepos, np.uint16), rc=rng.choice(np.frombuffer(b"dr", np.uint8), na),
                           order=rng.permutation(n).astype(np.uint32), lengths=lens, unaligned=np.frombuffer(dnaio.write_dnaN_records(seqs), np.uint8),
                           unaligned_len=int(lens[na:].sum()), num_aligned=na)


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31), st.integers(0, 120), st.sampled_from([1, 7, 64, 256000]), st.booleans(), st.booleans())
def test_decode_inverts_reblock_on_arbitrary_streams(seed, half, block, paired, preserve):
    rng = np.random.default_rng(seed)
    n = 2 * half if paired else half
    er = random_streams(rng, n, 40, 100000, float(rng.random()))
    use_order = paired or preserve
    slot = po.pe_encode(er.order) if paired and not preserve else (er.order if use_order else np.arange(n, dtype=np.uint32))
    if not preserve and n:
        # the -r format relies on aligned reads (read 1 of a pair) coming first in every block
        # (reorder_compress_streams.cpp:254-256 vs decompress.cpp:240-253): order the slots accordingly
        key = np.empty(n, np.int64); key[np.asarray(slot, np.int64)] = np.arange(n)      # slot -> stream index
        units = n // 2 if paired else n
        first_slots = np.arange(units)
        aligned_first = np.argsort(~(key[first_slots] < er.num_aligned), kind="stable")  # aligned read-1 units first
        remap = np.empty(units, np.int64); remap[aligned_first] = np.arange(units)
        slot = np.asarray(slot, np.int64)
        slot = np.where(slot < units, remap[slot % max(units, 1)], units + remap[slot % max(units, 1)] if paired else slot).astype(np.uint32)
    blocks = po.reblock(er, paired, preserve, block, order=slot if use_order else None)
    if not use_order and n:
        # SE -r ignores the order file: stream order is slot order, already aligned-first
        pass
    got = po.decode_blocks(blocks, er.seq, n, paired, preserve, block)
    stream_reads = po.decode(er)
    want = [None] * n
    for i, sl in enumerate(np.asarray(slot)):
        want[int(sl)] = stream_reads[i]
    assert got == want
    # sizes are conserved
    assert len(blocks.data["noise"]) == len(er.noise) and len(blocks.data["noisepos"]) == 2 * len(er.noisepos)
    assert len(blocks.data["unaligned"]) == er.unaligned_len and len(blocks.data["lengths"]) == 2 * n
