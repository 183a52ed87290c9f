"""Host-side mirror of the reference's interface for this path.

The reference exposes the path as two calls on a temp_dir (src/call_template_functions.h:9-11):
    call_reorder(temp_dir, cp);  call_encoder(temp_dir, cp);
invoked back to back from spring::compress (src/spring.cpp:153,166).  Same names, same argument
meaning, same error behaviour (RuntimeError where the reference throws std::runtime_error); both
run on the GPU through the C ABI -- there is no CPU path here.

Because the reorder -> encoder hand-off files (temp.dna.<t>, temppos.txt.<t>, ...) are private to
the two stages, the CUDA library keeps that hand-off in HBM: call_reorder does the whole GPU job
and parks the encoder streams, call_encoder writes them out (and BSC-compresses read_seq.bin.<t>
when a compressor is supplied, as pack_compress_seq does, src/encoder.cpp:148-153).
"""
from __future__ import annotations

import os

from . import capi, dnaio

_ctx = {}


def _context(device: int) -> capi.Context:
    if device not in _ctx:
        _ctx[device] = capi.Context(device)
    return _ctx[device]


def _as_cp(cp) -> capi.CP:
    if isinstance(cp, capi.CP):
        return cp
    if isinstance(cp, dnaio.CompressionParams):
        return capi.CP.from_buffer_copy(cp.pack())
    raise TypeError("cp must be a capi.CP or dnaio.CompressionParams")


def call_reorder(temp_dir: str, cp, device: int = 0, num_chains: int = 0) -> None:
    """reorder_main<> (src/reorder.h:732-786) + the encoder's GPU work; consumes
    input_clean_{1,2}.dna, input_N.dna, read_order_N.bin and leaves the encoder streams in temp_dir
    with read_seq.bin.<t> still un-compressed."""
    c = _as_cp(cp)
    if c.max_readlen > 511:
        raise RuntimeError("Wrong bitset size.")          # call_template_functions.cpp:61
    try:
        _context(device).reorder_encode_files(temp_dir, c, num_chains)
    except capi.SpringB200Error as e:
        raise RuntimeError(str(e)) from e


def call_encoder(temp_dir: str, cp, bsc_compress=None) -> None:
    """encoder_main<> (src/encoder.h:572-633): after call_reorder only pack_compress_seq's BSC step is
    left (src/encoder.cpp:148-153), which stays on the host: bsc_compress(infile, outfile) is called
    for every read_seq.bin.<t>, then the packed file is removed, exactly as the reference does."""
    c = _as_cp(cp)
    for t in range(max(1, c.num_thr)):
        base = os.path.join(temp_dir, f"read_seq.bin.{t}")
        if not os.path.exists(base):
            raise RuntimeError(f"call_encoder: {base} missing (call_reorder must run first)")
        if bsc_compress is not None:
            bsc_compress(base, base + ".bsc")
            os.remove(base)


def pe_encode(temp_dir: str, cp) -> None:
    """pe_encode (src/pe_encode.cpp:24-84, called at src/spring.cpp:193).  Nothing is left to do on the
    host: it only rewrites read_order.bin for reorder_compress_streams, and the GPU re-blocking applies
    the same mapping on the device (same contract as csrc/host/reorder_compress_streams_b200.cpp)."""
    _as_cp(cp)


def reorder_compress_streams(temp_dir: str, cp, device: int = 0, bsc_compress=None) -> None:
    """reorder_compress_streams (src/reorder_compress_streams.cpp:31-444): the GPU writes the raw per-block
    streams <name>.<b>; bsc_compress(infile, outfile), when given, is called on every one of them and the
    raw file removed, as :363-428 does."""
    c = _as_cp(cp)
    try:
        _context(device).reblock_files(temp_dir, c)
    except capi.SpringB200Error as e:
        raise RuntimeError(str(e)) from e
    if bsc_compress is None:
        return
    files = ("read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin", "read_rev.txt", "read_unaligned.txt",
             "read_lengths.bin", "read_pos_pair.bin", "read_rev_pair.txt")
    units = c.num_reads // 2 if c.paired_end else c.num_reads
    for b in range((units + c.num_reads_per_block - 1) // c.num_reads_per_block):
        for f in files[: 9 if c.paired_end else 7]:
            raw = os.path.join(temp_dir, f"{f}.{b}")
            bsc_compress(raw, raw + ".bsc")
            os.remove(raw)
