"""ctypes front-end of oracle/liboracle.so (spring_oracle.c) and of oracle/_ref/spring_ref.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from spring_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "spring_ref")
SPLICE_BIN = os.path.join(HERE, "_ref", "spring_b200_ref")
SPLICE2_BIN = os.path.join(HERE, "_ref", "spring_b200_ref2")  # + pe_encode / reorder_compress_streams on the GPU
SPLICE3_BIN = os.path.join(HERE, "_ref", "spring_b200_ref3")  # + preprocess's read path and decompress_short's block decode on the GPU
REFERENCE_SRC = "/root/reference"


def build(force: bool = False) -> None:
    """Compile the restatement; and, where the reference sources exist (this container, not the
    GPU box), the reference itself into oracle/_ref/."""
    srcs = [os.path.join(HERE, f) for f in ("spring_oracle.c", "reblock_oracle.c")]
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")) and (force or not os.path.exists(REF_BIN)):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])
    # the reference host with our library spliced in at call_reorder / call_encoder (end-to-end parity)
    b200 = os.path.join(HERE, "..", "spring_b200", "libspring_b200.so")
    if os.path.isdir(os.path.join(REFERENCE_SRC, "src")) and os.path.exists(b200):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "splice", "splice2", "splice3"])


class _ByteVec(C.Structure):
    _fields_ = [("p", C.POINTER(C.c_uint8)), ("n", C.c_size_t), ("cap", C.c_size_t)]

    def to_numpy(self, dtype=np.uint8) -> np.ndarray:
        if self.n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(self.p, shape=(self.n,)).copy().view(dtype)


class _Counters(C.Structure):
    _fields_ = [("probes", C.c_uint64), ("probe_hits", C.c_uint64), ("compares", C.c_uint64),
                ("passes", C.c_uint64), ("rounds", C.c_uint64), ("lost_proposals", C.c_uint64),
                ("unmatched", C.c_uint32)]


class _ReorderOut(C.Structure):
    _fields_ = [("order", C.POINTER(C.c_uint32)), ("flag", C.POINTER(C.c_uint8)),
                ("pos", C.POINTER(C.c_int64)), ("rc", C.POINTER(C.c_uint8)), ("n", C.c_uint64),
                ("s_order", C.POINTER(C.c_uint32)), ("n_s", C.c_uint64), ("ctr", _Counters)]


class _EncodeOut(C.Structure):
    _fields_ = [("seq", _ByteVec), ("pos", _ByteVec), ("noise", _ByteVec), ("noisepos", _ByteVec),
                ("rc", _ByteVec), ("order", _ByteVec), ("lengths", _ByteVec), ("unaligned", _ByteVec),
                ("unaligned_len", C.c_uint64), ("num_aligned", C.c_uint64),
                ("matched_s", C.c_uint32), ("matched_N", C.c_uint32),
                ("enc_probes", C.c_uint64), ("enc_compares", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        assert _lib.orc_sizeof_reorder_out() == C.sizeof(_ReorderOut)
        assert _lib.orc_sizeof_encode_out() == C.sizeof(_EncodeOut)
        _lib.orc_reorder_dict.restype = C.c_uint32
        _lib.orc_pack_seq.restype = C.c_uint64
    return _lib


def _p(a: np.ndarray, t):
    return a.ctypes.data_as(C.POINTER(t))


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).copy()


@dataclass
class ReorderResult:
    order: np.ndarray      # uint32 clean indices, stream order
    flag: np.ndarray       # uint8 0 = first read of a contig, 1 = matched
    pos: np.ndarray        # int64
    rc: np.ndarray         # uint8 'd' / 'r'
    s_order: np.ndarray    # uint32 singletons
    counters: dict


def reorder(packed: np.ndarray, lengths: np.ndarray, max_readlen: int, num_chains: int = 1) -> ReorderResult:
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
    out = _ReorderOut()
    rc = lib().orc_reorder(_p(packed, C.c_uint64), _p(lengths, C.c_uint16), C.c_uint32(len(lengths)),
                           C.c_int(max_readlen), C.c_int(num_chains), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_reorder failed: {rc}")
    res = ReorderResult(_arr(out.order, out.n, np.uint32), _arr(out.flag, out.n, np.uint8),
                        _arr(out.pos, out.n, np.int64), _arr(out.rc, out.n, np.uint8),
                        _arr(out.s_order, out.n_s, np.uint32),
                        {f: getattr(out.ctr, f) for f, _ in _Counters._fields_})
    lib().orc_reorder_free(C.byref(out))
    return res


@dataclass
class EncodeResult:
    """The encoder's output streams (SURVEY 8b "data contract out"), un-sharded."""
    seq: np.ndarray         # uint8 ASCII consensus
    pos: np.ndarray         # uint64
    noise: np.ndarray       # uint8
    noisepos: np.ndarray    # uint16
    rc: np.ndarray          # uint8
    order: np.ndarray       # uint32 (aligned then unaligned)
    lengths: np.ndarray     # uint16 (same order)
    unaligned: np.ndarray   # uint8 4-bit records
    unaligned_len: int
    num_aligned: int
    matched_s: int
    matched_N: int
    enc_probes: int = 0
    enc_compares: int = 0

    def packed_seq(self) -> tuple[bytes, bytes]:
        """(2-bit packed bytes, ASCII tail) as pack_compress_seq writes them (encoder.cpp:126-147)."""
        n = len(self.seq)
        buf = np.zeros(n // 4, dtype=np.uint8)
        seq = np.ascontiguousarray(self.seq)
        lib().orc_pack_seq(_p(seq, C.c_uint8), C.c_uint64(n), _p(buf, C.c_uint8))
        return buf.tobytes(), seq[n - n % 4:].tobytes() if n % 4 else b""


def encode(packed, lengths, max_readlen, ro: ReorderResult, n_records: bytes, order_n: np.ndarray,
           num_total: int) -> EncodeResult:
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
    nrec = np.frombuffer(n_records, dtype=np.uint8).copy() if n_records else np.zeros(1, np.uint8)
    order_n = np.ascontiguousarray(order_n, dtype=np.uint32)
    onp = order_n if len(order_n) else np.zeros(1, np.uint32)
    out = _EncodeOut()
    so = ro.s_order if len(ro.s_order) else np.zeros(1, np.uint32)
    st = [a if len(a) else np.zeros(1, a.dtype) for a in (ro.order, ro.flag, ro.pos, ro.rc)]
    rc = lib().orc_encode(_p(packed, C.c_uint64), _p(lengths, C.c_uint16), C.c_uint32(len(lengths)), C.c_int(max_readlen),
                          _p(st[0], C.c_uint32), _p(st[1], C.c_uint8), _p(st[2], C.c_int64), _p(st[3], C.c_uint8),
                          C.c_uint64(len(ro.order)), _p(so, C.c_uint32), C.c_uint64(len(ro.s_order)),
                          _p(nrec, C.c_uint8), C.c_uint64(len(n_records)), _p(onp, C.c_uint32), C.c_uint32(len(order_n)),
                          C.c_uint32(num_total), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"orc_encode failed: {rc}")
    res = EncodeResult(out.seq.to_numpy(), out.pos.to_numpy(np.uint64), out.noise.to_numpy(), out.noisepos.to_numpy(np.uint16),
                       out.rc.to_numpy(), out.order.to_numpy(np.uint32), out.lengths.to_numpy(np.uint16),
                       out.unaligned.to_numpy(), out.unaligned_len, out.num_aligned, out.matched_s, out.matched_N,
                       out.enc_probes, out.enc_compares)
    lib().orc_encode_free(C.byref(out))
    return res


def reorder_encode(packed, lengths, max_readlen, n_records=b"", order_n=None, num_total=None, num_chains=1):
    order_n = np.zeros(0, np.uint32) if order_n is None else order_n
    ro = reorder(packed, lengths, max_readlen, num_chains)
    if num_total is None:
        num_total = len(lengths) + len(order_n)
    return ro, encode(packed, lengths, max_readlen, ro, n_records, order_n, num_total)


def decode(er, stride: int | None = None) -> list[bytes]:
    """Rebuild every read of the stream (aligned then unaligned), decompress.cpp:263-283."""
    out = decode_matrix(er, stride)
    return [out[i, : er.lengths[i]].tobytes() for i in range(len(er.lengths))]


def decode_matrix(er, stride: int | None = None) -> np.ndarray:
    """Same, as one uint8[num_reads, stride] ASCII matrix (rows zero-padded): for full-size checks."""
    n = len(er.lengths)
    stride = stride or (int(er.lengths.max()) if n else 1) or 1
    out = np.zeros((max(n, 1), stride), dtype=np.uint8)
    a = [np.ascontiguousarray(x) if len(x) else np.zeros(1, x.dtype) for x in
         (er.seq, er.pos, er.noise, er.noisepos, er.rc, er.lengths, er.unaligned)]
    rc = lib().orc_decode(_p(a[0], C.c_uint8), C.c_uint64(len(er.seq)), _p(a[1], C.c_uint64), _p(a[2], C.c_uint8),
                          C.c_uint64(len(er.noise)), _p(a[3], C.c_uint16), C.c_uint64(len(er.noisepos)),
                          _p(a[4], C.c_uint8), C.c_uint64(er.num_aligned), _p(a[5], C.c_uint16), C.c_uint64(n),
                          _p(a[6], C.c_uint8), C.c_uint64(len(er.unaligned)), _p(out, C.c_uint8), C.c_int(stride))
    if rc != 0:
        raise RuntimeError(f"orc_decode failed: {rc}")
    return out[:n]


def reorder_dict(packed, lengths, max_readlen, which):
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
    n = len(lengths)
    keys = np.zeros(max(n, 1), np.uint64); bs = np.zeros(n + 1, np.uint32); rid = np.zeros(max(n, 1), np.uint32)
    dn = C.c_uint32(0)
    nk = lib().orc_reorder_dict(_p(packed, C.c_uint64), _p(lengths, C.c_uint16), C.c_uint32(n), C.c_int(max_readlen),
                                C.c_int(which), _p(keys, C.c_uint64), _p(bs, C.c_uint32), _p(rid, C.c_uint32), C.byref(dn))
    return keys[:nk].copy(), bs[: nk + 1].copy(), rid[: dn.value].copy()


# ---------------------------------------------------------------------------------------------
# the real reference (oracle/_ref/spring_ref), when it was built
# ---------------------------------------------------------------------------------------------
def have_reference() -> bool:
    return os.path.exists(REF_BIN)


def run_reference_hotpath(temp_dir: str, num_thr: int = 1, unbsc: bool = True) -> tuple[float, float, str]:
    """call_reorder + call_encoder of the unmodified reference on a prepared temp_dir.
    returns (reorder seconds, encode seconds, stdout)."""
    cmd = [REF_BIN, "--hotpath", "--temp", temp_dir, "-t", str(num_thr)] + (["--unbsc"] if unbsc else [])
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=temp_dir)
    if r.returncode != 0:
        raise RuntimeError(f"spring_ref --hotpath failed:\n{r.stdout}\n{r.stderr}")
    line = [l for l in r.stdout.splitlines() if l.startswith("HOTPATH_SECONDS")][0].split()
    return float(line[1]), float(line[2]), r.stdout


def load_reference_streams(temp_dir: str, num_thr: int) -> EncodeResult:
    """Read what the reference's encoder left in temp_dir (after --unbsc) into an EncodeResult."""
    def rd(name, dtype=np.uint8):
        p = os.path.join(temp_dir, name)
        return np.fromfile(p, dtype=dtype) if os.path.getsize(p) else np.zeros(0, dtype)
    seq_parts = []
    code = np.frombuffer(b"ACGT", dtype=np.uint8)
    for t in range(num_thr):
        b = rd(f"read_seq.bin.{t}")
        chars = code[np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).reshape(-1)] if len(b) else np.zeros(0, np.uint8)
        seq_parts += [chars, rd(f"read_seq.bin.{t}.tail")]
    seq = np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
    rc = rd("read_rev.txt")
    ul = int(np.fromfile(os.path.join(temp_dir, "read_unaligned.txt.count"), dtype=np.uint64)[0])
    return EncodeResult(seq, rd("read_pos.bin", np.uint64), rd("read_noise.txt"), rd("read_noisepos.bin", np.uint16), rc,
                        rd("read_order.bin", np.uint32), rd("read_lengths.bin", np.uint16), rd("read_unaligned.txt"),
                        ul, len(rc), -1, -1)


# ---------------------------------------------------------------------------------------------
# pe_encode + the re-blocking of reorder_compress_streams (oracle/reblock_oracle.c), SURVEY 8(f)
# ---------------------------------------------------------------------------------------------
BLOCK_STREAMS = ("flag", "pos", "noise", "noisepos", "rc", "unaligned", "lengths", "pos_pair", "rc_pair")
BLOCK_FILES = ("read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin", "read_rev.txt",
               "read_unaligned.txt", "read_lengths.bin", "read_pos_pair.bin", "read_rev_pair.txt")


class _BlocksOut(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8) * 9), ("size", C.c_uint64 * 9), ("cap", C.c_uint64 * 9),
                ("off", C.POINTER(C.c_uint64) * 9), ("num_blocks", C.c_uint32)]


@dataclass
class BlockStreams:
    """The nine per-block streams reorder_compress_streams hands to BSC, blocks concatenated.
    data[s]: uint8 bytes; off[s]: uint64[num_blocks + 1] byte offsets of the blocks in data[s]."""
    num_blocks: int
    data: dict
    off: dict

    def block(self, stream: str, b: int) -> bytes:
        o = self.off[stream]
        return self.data[stream][int(o[b]): int(o[b + 1])].tobytes()


def pe_encode(order: np.ndarray) -> np.ndarray:
    """pe_encode.cpp:24-84 on read_order.bin's contents."""
    o = np.ascontiguousarray(order, dtype=np.uint32).copy()
    buf = o if len(o) else np.zeros(1, np.uint32)
    if lib().orc_pe_encode(_p(buf, C.c_uint32), C.c_uint32(len(o))) != 0:
        raise RuntimeError("orc_pe_encode failed")
    return o


def reblock(er, paired_end: bool, preserve_order: bool, num_reads_per_block: int = 256000, order=None) -> BlockStreams:
    """reorder_compress_streams.cpp:83-361 on the encoder streams `er` (EncodeResult / StreamsResult);
    `order` overrides er.order (the output of pe_encode for -r paired input)."""
    assert lib().orc_sizeof_blocks_out() == C.sizeof(_BlocksOut)
    order = er.order if order is None else order
    nz = lambda a, dt: np.ascontiguousarray(a, dtype=dt) if len(a) else np.zeros(1, dt)
    pos, noise, noisepos, rc = nz(er.pos, np.uint64), nz(er.noise, np.uint8), nz(er.noisepos, np.uint16), nz(er.rc, np.uint8)
    ordr, lens, unal = nz(order, np.uint32), nz(er.lengths, np.uint16), nz(er.unaligned, np.uint8)
    out = _BlocksOut()
    rcode = lib().orc_reblock(_p(pos, C.c_uint64), _p(noise, C.c_uint8), C.c_uint64(len(er.noise)), _p(noisepos, C.c_uint16),
                              _p(rc, C.c_uint8), C.c_uint64(int(er.num_aligned)), _p(ordr, C.c_uint32), _p(lens, C.c_uint16),
                              C.c_uint64(len(er.lengths)), _p(unal, C.c_uint8), C.c_uint64(len(er.unaligned)),
                              C.c_int(int(paired_end)), C.c_int(int(preserve_order)), C.c_uint32(num_reads_per_block), C.byref(out))
    if rcode != 0:
        raise RuntimeError(f"orc_reblock failed: {rcode}")
    nb = out.num_blocks
    data, off = {}, {}
    for i, name in enumerate(BLOCK_STREAMS):
        n = out.size[i]
        data[name] = np.ctypeslib.as_array(out.data[i], shape=(n,)).copy() if n else np.zeros(0, np.uint8)
        off[name] = np.ctypeslib.as_array(out.off[i], shape=(nb + 1,)).copy()
    lib().orc_blocks_free(C.byref(out))
    return BlockStreams(nb, data, off)


def write_encoder_streams(temp_dir: str, er, cp_bytes: bytes) -> None:
    """Lay the encoder's streams out as files, as encoder.h:386-487 leaves them for the later host stages
    (read_seq.bin.* is not needed by them), plus cp_in.bin for spring_ref --reblock."""
    os.makedirs(temp_dir, exist_ok=True)
    w = lambda name, a, dt: np.ascontiguousarray(a, dtype=dt).tofile(os.path.join(temp_dir, name))
    w("read_pos.bin", er.pos, np.uint64); w("read_noise.txt", er.noise, np.uint8); w("read_noisepos.bin", er.noisepos, np.uint16)
    w("read_rev.txt", er.rc, np.uint8); w("read_order.bin", er.order, np.uint32); w("read_lengths.bin", er.lengths, np.uint16)
    w("read_unaligned.txt", er.unaligned, np.uint8)
    np.array([er.unaligned_len], np.uint64).tofile(os.path.join(temp_dir, "read_unaligned.txt.count"))
    with open(os.path.join(temp_dir, "cp_in.bin"), "wb") as f:
        f.write(cp_bytes)


def run_reference_reblock(temp_dir: str, paired_end: bool, num_blocks: int, num_thr: int = 2) -> BlockStreams:
    """pe_encode + reorder_compress_streams of the unmodified reference on a temp_dir laid out by
    write_encoder_streams; returns the raw (BSC-decoded) per-block files."""
    r = subprocess.run([REF_BIN, "--reblock", "--temp", temp_dir, "-t", str(num_thr)], capture_output=True, text=True, cwd=temp_dir)
    if r.returncode != 0:
        raise RuntimeError(f"spring_ref --reblock failed:\n{r.stdout}\n{r.stderr}")
    data, off = {}, {}
    for name, fn in zip(BLOCK_STREAMS, BLOCK_FILES):
        parts, o = [], [0]
        for b in range(num_blocks):
            p = os.path.join(temp_dir, f"{fn}.{b}")
            if not paired_end and name in ("pos_pair", "rc_pair"):
                a = np.zeros(0, np.uint8)
            else:
                a = np.fromfile(p, dtype=np.uint8) if os.path.getsize(p) else np.zeros(0, np.uint8)
            parts.append(a); o.append(o[-1] + len(a))
        data[name] = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        off[name] = np.array(o, np.uint64)
    return BlockStreams(num_blocks, data, off)


# ---------------------------------------------------------------------------------------------
# preprocess's read path: N split + record packing, as file images
# ---------------------------------------------------------------------------------------------
def pack_reads(seqs: list, num_file1: int | None = None) -> dict:
    """Plain-Python restatement of what preprocess does with the sequence lines (src/preprocess.cpp:196-207,
    :293-304, :364-378) and of the two record packers (src/util.cpp:269-294 write_dna_in_bits,
    :322-348 write_dnaN_in_bits).  Returns the byte images of input_clean_1.dna, input_clean_2.dna,
    input_N.dna, read_order_N.bin and the cp fields preprocess sets (:398-403)."""
    n = len(seqs)
    n1 = n if num_file1 is None else num_file1
    d2 = {ord("A"): 0, ord("C"): 2, ord("G"): 1, ord("T"): 3}
    d4 = dict(d2); d4[ord("N")] = 4
    clean = [bytearray(), bytearray()]
    rec_n, order_n = bytearray(), []
    num_clean = [0, 0]
    max_len = 0
    for i, s in enumerate(seqs):
        j = 0 if i < n1 else 1
        L = len(s)
        if L > 511:
            raise ValueError("Too long read length (please try --long/-l flag).")   # preprocess.cpp:199-206
        max_len = max(max_len, L)
        if b"N" not in s:                                                            # :207-209, :296-298
            out = clean[j]
            out += L.to_bytes(2, "little")
            for k in range(0, L, 4):
                v = 0
                for t, ch in enumerate(s[k:k + 4]):
                    v |= d2[ch] << (2 * t)
                out.append(v)
            num_clean[j] += 1
        else:                                                                        # :299-303
            order_n.append(i)
            rec_n += L.to_bytes(2, "little")
            for k in range(0, L, 2):
                v = 0
                for t, ch in enumerate(s[k:k + 2]):
                    v |= d4[ch] << (4 * t)
                rec_n.append(v)
    return dict(clean_1=bytes(clean[0]), clean_2=bytes(clean[1]), n_records=bytes(rec_n),
                order_n=np.array(order_n, dtype=np.uint32), num_reads_clean=tuple(num_clean), max_readlen=max_len, num_reads=n)


def run_reference_preprocess(temp_dir: str, fastq1: str, fastq2: str | None = None, num_thr: int = 2) -> dict:
    """spring::preprocess of the unmodified reference (-r --no-quality --no-ids) into temp_dir; returns the
    same dict as pack_reads, read back from the files it wrote."""
    ins = [fastq1] + ([fastq2] if fastq2 else [])
    r = subprocess.run([REF_BIN, "--preprocess", "-i", *ins, "--temp", temp_dir, "-r", "--no-quality", "--no-ids", "-t", str(num_thr)],
                       capture_output=True, text=True, cwd=temp_dir)
    if r.returncode != 0:
        raise RuntimeError(f"spring_ref --preprocess failed:\n{r.stdout}\n{r.stderr}")
    rd = lambda f: open(os.path.join(temp_dir, f), "rb").read() if os.path.exists(os.path.join(temp_dir, f)) else b""
    import struct
    cp = rd("cp_in.bin")
    num_reads, c0, c1, max_readlen = struct.unpack_from("<IIII", cp, 28)
    return dict(clean_1=rd("input_clean_1.dna"), clean_2=rd("input_clean_2.dna"), n_records=rd("input_N.dna"),
                order_n=np.frombuffer(rd("read_order_N.bin"), dtype=np.uint32), num_reads_clean=(c0, c1), max_readlen=max_readlen,
                num_reads=num_reads)


# ---------------------------------------------------------------------------------------------
# decompress_short's block decode (oracle/reblock_oracle.c:orc_decode_blocks), SURVEY 8f rank 4
# ---------------------------------------------------------------------------------------------
def decode_blocks(blocks, seq_ascii: np.ndarray, num_reads: int, paired_end: bool, preserve_order: bool,
                  num_reads_per_block: int = 256000) -> list:
    """decompress.cpp:230-320 on the per-block streams (BlockStreams / BlocksResult layout) and the ASCII
    consensus: every read, file 1's in output order, then file 2's."""
    nb = blocks.num_blocks
    datas = [np.ascontiguousarray(blocks.data[s], dtype=np.uint8) if len(blocks.data[s]) else np.zeros(1, np.uint8) for s in BLOCK_STREAMS]
    offs = [np.ascontiguousarray(blocks.off[s], dtype=np.uint64) for s in BLOCK_STREAMS]
    dp = (C.POINTER(C.c_uint8) * 9)(*[_p(d, C.c_uint8) for d in datas])
    op = (C.POINTER(C.c_uint64) * 9)(*[_p(o, C.c_uint64) for o in offs])
    seq = np.ascontiguousarray(seq_ascii, dtype=np.uint8) if len(seq_ascii) else np.zeros(1, np.uint8)
    total = int(np.frombuffer(blocks.data["lengths"].tobytes(), np.uint16).astype(np.int64).sum())
    out = np.zeros(max(total, 1), np.uint8)
    out_off = np.zeros(num_reads + 1, np.uint64)
    out_len = np.zeros(max(num_reads, 1), np.uint16)
    rc = lib().orc_decode_blocks(dp, op, C.c_uint32(nb), _p(seq, C.c_uint8), C.c_uint64(len(seq_ascii)), C.c_uint64(num_reads),
                                 C.c_int(int(paired_end)), C.c_int(int(preserve_order)), C.c_uint32(num_reads_per_block),
                                 _p(out, C.c_uint8), _p(out_off, C.c_uint64), _p(out_len, C.c_uint16))
    if rc != 0:
        raise RuntimeError(f"orc_decode_blocks failed: {rc}")
    return [out[int(out_off[i]): int(out_off[i + 1])].tobytes() for i in range(num_reads)]


def load_archive_blocks(archive: str, work_dir: str):
    """Untar a SPRING archive, BSC-decode its read_* files with the reference's own decoder and return
    (cp bytes, BlockStreams, ASCII consensus): what decompress_short starts from (decompress.cpp:83-229)."""
    import tarfile
    with tarfile.open(archive) as t:
        t.extractall(work_dir, filter="data")
    cpb = open(os.path.join(work_dir, "cp.bin"), "rb").read()
    bsc = sorted(os.path.join(work_dir, f) for f in os.listdir(work_dir) if f.startswith("read_") and f.endswith(".bsc"))
    r = subprocess.run([REF_BIN, "--bsc-decode", "-i", *bsc], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stdout + r.stderr)
    import struct
    paired = bool(cpb[0])
    num_reads = struct.unpack_from("<I", cpb, 28)[0]
    block, = struct.unpack_from("<i", cpb, 48)
    num_thr, = struct.unpack_from("<i", cpb, 56)
    units = num_reads // 2 if paired else num_reads
    nb = (units + block - 1) // block
    data, off = {}, {}
    for name, fn in zip(BLOCK_STREAMS, BLOCK_FILES):
        parts, o = [], [0]
        for b in range(nb):
            p = os.path.join(work_dir, f"{fn}.{b}")
            a = np.fromfile(p, dtype=np.uint8) if os.path.exists(p) and os.path.getsize(p) else np.zeros(0, np.uint8)
            parts.append(a); o.append(o[-1] + len(a))
        data[name] = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
        off[name] = np.array(o, np.uint64)
    code = np.frombuffer(b"ACGT", dtype=np.uint8)
    seq_parts = []
    for t in range(num_thr):  # decompress_unpack_seq, decompress.cpp:615-660
        b = np.fromfile(os.path.join(work_dir, f"read_seq.bin.{t}"), dtype=np.uint8)
        seq_parts.append(code[np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).reshape(-1)] if len(b) else np.zeros(0, np.uint8))
        tp = os.path.join(work_dir, f"read_seq.bin.{t}.tail")
        seq_parts.append(np.fromfile(tp, dtype=np.uint8) if os.path.getsize(tp) else np.zeros(0, np.uint8))
    return cpb, BlockStreams(nb, data, off), np.concatenate(seq_parts) if seq_parts else np.zeros(0, np.uint8)
