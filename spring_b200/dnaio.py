"""Record formats on either side of the hot path (host side, numpy).

Citations are relative to /root/reference/src:
  * ``input_clean_{1,2}.dna`` / ``temp.dna``: ``{u16 len; ceil(len/4) B}``, 2 bits/base,
    A0 G1 C2 T3, base j in bits 2(j%4) of byte j/4            (util.cpp:269-294)
  * ``input_N.dna`` / ``read_unaligned.txt``: ``{u16 len; ceil(len/2) B}``, 4 bits/base,
    A0 G1 C2 T3 N4                                             (util.cpp:322-348)
  * in memory the reference keeps one ``std::bitset<64*W>`` per read, W = (2L-1)/64+1
    (call_template_functions.cpp:10), filled by copying the file bytes (reorder.h:229):
    here a ``uint64[N, W]`` array + ``uint16[N]`` lengths.
  * ``cp.bin``: raw 64-byte ``compression_params``              (util.h:30-51)
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

CODE2CHAR = np.frombuffer(b"AGCT", dtype=np.uint8)          # reorder.h:80
CHAR2CODE = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"AGCT"):
    CHAR2CODE[_c] = _i
CODE4CHAR = np.frombuffer(b"AGCTN", dtype=np.uint8)          # util.cpp:353
CHAR2CODE4 = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"AGCTN"):
    CHAR2CODE4[_c] = _i


def words_per_read(max_readlen: int) -> int:
    """uint64 words of the reorder bitset (call_template_functions.cpp:10)."""
    return (2 * max_readlen - 1) // 64 + 1


def pack_codes(codes: np.ndarray, lengths: np.ndarray, max_readlen: int) -> np.ndarray:
    """codes: uint8[N, Lmax] (values 0..3, garbage beyond each length) -> uint64[N, W]."""
    n, lmax = codes.shape
    w = words_per_read(max_readlen)
    full = np.zeros((n, w * 32), dtype=np.uint64)
    valid = np.arange(lmax)[None, :] < lengths[:, None]
    full[:, :lmax] = np.where(valid, codes, 0)
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    return (full.reshape(n, w, 32) << shifts).sum(axis=2, dtype=np.uint64)


def unpack_codes(packed: np.ndarray, max_len: int) -> np.ndarray:
    """uint64[N, W] -> uint8[N, max_len] 2-bit codes."""
    n, w = packed.shape
    shifts = (2 * np.arange(32, dtype=np.uint64))[None, None, :]
    codes = ((packed[:, :, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(n, w * 32)
    return codes[:, :max_len]


def seqs_to_packed(seqs: list[bytes], max_readlen: int) -> tuple[np.ndarray, np.ndarray]:
    n = len(seqs)
    lengths = np.array([len(s) for s in seqs], dtype=np.uint16)
    lmax = max(int(lengths.max()) if n else 0, 1)
    codes = np.zeros((n, lmax), dtype=np.uint8)
    for i, s in enumerate(seqs):
        codes[i, : len(s)] = CHAR2CODE[np.frombuffer(s, dtype=np.uint8)]
    return pack_codes(codes, lengths, max_readlen), lengths


def packed_to_seqs(packed: np.ndarray, lengths: np.ndarray) -> list[bytes]:
    lmax = int(lengths.max()) if len(lengths) else 0
    chars = CODE2CHAR[unpack_codes(packed, lmax)]
    return [chars[i, : lengths[i]].tobytes() for i in range(len(lengths))]


def write_dna_file(path: str, packed: np.ndarray, lengths: np.ndarray) -> None:
    """write_dna_in_bits records (util.cpp:269-294) from the in-memory layout."""
    n = len(lengths)
    raw = np.ascontiguousarray(packed).view(np.uint8).reshape(n, -1)
    with open(path, "wb") as f:
        if n and (lengths == lengths[0]).all():
            nb = (int(lengths[0]) + 3) // 4
            step = 1 << 22
            for lo in range(0, n, step):  # chunked: a 100 M-read file is 4 GB
                m = min(step, n - lo)
                rec = np.empty((m, 2 + nb), dtype=np.uint8)
                rec[:, 0] = lengths[0] & 0xFF
                rec[:, 1] = lengths[0] >> 8
                rec[:, 2:] = raw[lo:lo + m, :nb]
                f.write(rec.tobytes())
        else:  # variable lengths: records {u16 len; ceil(len/4) B} selected out of a padded matrix, chunk by chunk
            step = 1 << 20
            for lo in range(0, n, step):
                ln = lengths[lo:lo + step].astype(np.int64)
                nb = (ln + 3) // 4
                width = 2 + int(nb.max())
                rec = np.zeros((len(ln), width), dtype=np.uint8)
                rec[:, 0] = ln & 0xFF
                rec[:, 1] = ln >> 8
                rec[:, 2:] = raw[lo:lo + step, : width - 2]
                f.write(rec[np.arange(width)[None, :] < (2 + nb)[:, None]].tobytes())


def read_dna_file(path: str, num_reads: int, max_readlen: int) -> tuple[np.ndarray, np.ndarray]:
    """readDnaFile (reorder.h:222-244): records copied straight into bitset storage."""
    w = words_per_read(max_readlen)
    out = np.zeros((num_reads, w * 8), dtype=np.uint8)
    lengths = np.zeros(num_reads, dtype=np.uint16)
    data = np.fromfile(path, dtype=np.uint8) if num_reads else np.zeros(0, np.uint8)
    off = 0
    for i in range(num_reads):
        ln = int(data[off]) | (int(data[off + 1]) << 8)
        nb = (ln + 3) // 4
        out[i, :nb] = data[off + 2 : off + 2 + nb]
        lengths[i] = ln
        off += 2 + nb
    return out.view(np.uint64).reshape(num_reads, w), lengths


def write_dnaN_records(seqs: list[bytes]) -> bytes:
    """write_dnaN_in_bits (util.cpp:322-348)."""
    parts = []
    for s in seqs:
        codes = CHAR2CODE4[np.frombuffer(s, dtype=np.uint8)]
        if len(codes) % 2:
            codes = np.concatenate([codes, np.zeros(1, np.uint8)])
        parts.append(struct.pack("<H", len(s)))
        parts.append((codes[0::2] | (codes[1::2] << 4)).astype(np.uint8).tobytes())
    return b"".join(parts)


def read_dnaN_records(buf: bytes, num: int | None = None) -> list[bytes]:
    """read_dnaN_from_bits (util.cpp:350-374)."""
    data = np.frombuffer(buf, dtype=np.uint8)
    out, off = [], 0
    while off < len(data) and (num is None or len(out) < num):
        ln = int(data[off]) | (int(data[off + 1]) << 8)
        nb = (ln + 1) // 2
        b = data[off + 2 : off + 2 + nb]
        codes = np.empty(nb * 2, dtype=np.uint8)
        codes[0::2] = b & 15
        codes[1::2] = b >> 4
        out.append(CODE4CHAR[codes[:ln]].tobytes())
        off += 2 + nb
    return out


_CP_FMT = "<8?dIIIIIIIB?2xiii4x"   # util.h:30-51, offsets verified with offsetof


@dataclass
class CompressionParams:
    """Mirror of spring::compression_params (util.h:30-51); 64 bytes on disk (cp.bin)."""
    paired_end: bool = False
    preserve_order: bool = False
    preserve_quality: bool = False
    preserve_id: bool = False
    long_flag: bool = False
    qvz_flag: bool = False
    ill_bin_flag: bool = False
    bin_thr_flag: bool = False
    qvz_ratio: float = 0.0
    bin_thr_thr: int = 0
    bin_thr_high: int = 0
    bin_thr_low: int = 0
    num_reads: int = 0
    num_reads_clean_0: int = 0
    num_reads_clean_1: int = 0
    max_readlen: int = 0
    paired_id_code: int = 0
    paired_id_match: bool = False
    num_reads_per_block: int = 256000
    num_reads_per_block_long: int = 10000
    num_thr: int = 1

    def pack(self) -> bytes:
        b = struct.pack(
            _CP_FMT, self.paired_end, self.preserve_order, self.preserve_quality, self.preserve_id,
            self.long_flag, self.qvz_flag, self.ill_bin_flag, self.bin_thr_flag, self.qvz_ratio,
            self.bin_thr_thr, self.bin_thr_high, self.bin_thr_low, self.num_reads,
            self.num_reads_clean_0, self.num_reads_clean_1, self.max_readlen, self.paired_id_code,
            self.paired_id_match, self.num_reads_per_block, self.num_reads_per_block_long, self.num_thr)
        assert len(b) == 64
        return b

    @classmethod
    def unpack(cls, b: bytes) -> "CompressionParams":
        return cls(*struct.unpack(_CP_FMT, b[:64]))


def write_hotpath_inputs(temp_dir: str, packed: np.ndarray, lengths: np.ndarray, *,
                         max_readlen: int, n_seqs: list[bytes] = (), order_n: np.ndarray | None = None,
                         num_reads: int | None = None, paired_split: int | None = None,
                         num_thr: int = 1) -> CompressionParams:
    """Lay out a temp_dir exactly as preprocess leaves it for call_reorder (preprocess.cpp:296-403):
    input_clean_1.dna [input_clean_2.dna], input_N.dna, read_order_N.bin, plus cp_in.bin for
    oracle/_ref/spring_ref --hotpath."""
    os.makedirs(temp_dir, exist_ok=True)
    n = len(lengths)
    n1 = n if paired_split is None else paired_split
    write_dna_file(os.path.join(temp_dir, "input_clean_1.dna"), packed[:n1], lengths[:n1])
    if paired_split is not None:
        write_dna_file(os.path.join(temp_dir, "input_clean_2.dna"), packed[n1:], lengths[n1:])
    with open(os.path.join(temp_dir, "input_N.dna"), "wb") as f:
        f.write(write_dnaN_records(list(n_seqs)))
    order_n = np.zeros(0, np.uint32) if order_n is None else np.asarray(order_n, dtype=np.uint32)
    order_n.tofile(os.path.join(temp_dir, "read_order_N.bin"))
    cp = CompressionParams(paired_end=paired_split is not None, num_reads=num_reads if num_reads is not None else n + len(n_seqs),
                           num_reads_clean_0=n1, num_reads_clean_1=n - n1, max_readlen=max_readlen, num_thr=num_thr)
    with open(os.path.join(temp_dir, "cp_in.bin"), "wb") as f:
        f.write(cp.pack())
    return cp
