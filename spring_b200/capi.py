"""ctypes binding of libspring_b200.so (include/spring_b200.h).

The CUDA library is the only implementation: if it is missing, or no GPU is visible, every
compute call raises -- there is no CPU fallback on the product path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPRING_B200_LIB") or os.path.join(HERE, "libspring_b200.so")

EXPORTS = [
    "spring_b200_version", "spring_b200_device_count", "spring_b200_create", "spring_b200_destroy",
    "spring_b200_last_error", "spring_b200_get_stats", "spring_b200_reorder_encode",
    "spring_b200_reorder_encode_device", "spring_b200_fetch_streams", "spring_b200_build_dictionary",
    "spring_b200_reorder", "spring_b200_reorder_encode_files", "spring_b200_write_streams",
    "spring_b200_bucket_reads", "spring_b200_set_schedule", "spring_b200_set_chain_stats", "spring_b200_set_stitch", "spring_b200_packed_pending", "spring_b200_reorder_encode_packed", "spring_b200_fetch_reorder", "spring_b200_set_stream",
    "spring_b200_pe_encode", "spring_b200_reblock_streams", "spring_b200_reblock_files", "spring_b200_pack_reads",
    "spring_b200_decode_blocks", "spring_b200_verify_roundtrip",
    "spring_b200_comm_unique_id", "spring_b200_comm_init", "spring_b200_comm_free", "spring_b200_exchange_reads",
    "spring_b200_finalize_shard", "spring_b200_merge_shards", "spring_b200_free_merged", "spring_b200_write_merged",
    "spring_b200_shared_ctx", "spring_b200_reorder_encode_files_multi",
]


class SpringB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"spring_b200 error {code}: {msg}")
        self.code = code


class CP(C.Structure):
    """spring_b200_cp == spring::compression_params (src/util.h:30-51)."""
    _fields_ = [("paired_end", C.c_uint8), ("preserve_order", C.c_uint8), ("preserve_quality", C.c_uint8),
                ("preserve_id", C.c_uint8), ("long_flag", C.c_uint8), ("qvz_flag", C.c_uint8),
                ("ill_bin_flag", C.c_uint8), ("bin_thr_flag", C.c_uint8), ("qvz_ratio", C.c_double),
                ("bin_thr_thr", C.c_uint32), ("bin_thr_high", C.c_uint32), ("bin_thr_low", C.c_uint32),
                ("num_reads", C.c_uint32), ("num_reads_clean", C.c_uint32 * 2), ("max_readlen", C.c_uint32),
                ("paired_id_code", C.c_uint8), ("paired_id_match", C.c_uint8),
                ("num_reads_per_block", C.c_int32), ("num_reads_per_block_long", C.c_int32), ("num_thr", C.c_int32)]


class Input(C.Structure):
    _fields_ = [("reads", C.c_void_p), ("lengths", C.c_void_p), ("num_clean", C.c_uint32), ("max_readlen", C.c_uint32),
                ("n_records", C.c_void_p), ("n_record_bytes", C.c_uint64), ("order_n", C.c_void_p),
                ("num_n", C.c_uint32), ("num_reads", C.c_uint32)]


class Streams(C.Structure):
    _fields_ = [("seq_packed", C.c_void_p), ("seq_len", C.c_uint64), ("pos", C.c_void_p), ("noise", C.c_void_p),
                ("noise_bytes", C.c_uint64), ("noisepos", C.c_void_p), ("num_noise", C.c_uint64), ("rev", C.c_void_p),
                ("order", C.c_void_p), ("lengths", C.c_void_p), ("unaligned", C.c_void_p),
                ("unaligned_bytes", C.c_uint64), ("unaligned_len", C.c_uint64), ("num_aligned", C.c_uint64),
                ("num_reads", C.c_uint64), ("singletons_aligned", C.c_uint32), ("n_reads_aligned", C.c_uint32)]


class ReorderOut(C.Structure):
    _fields_ = [("order", C.c_void_p), ("flag", C.c_void_p), ("pos", C.c_void_p), ("rev", C.c_void_p), ("num", C.c_uint64),
                ("singleton_order", C.c_void_p), ("num_singletons", C.c_uint64)]


NUM_BLOCK_STREAMS = 9
BLOCK_STREAMS = ("flag", "pos", "noise", "noisepos", "rc", "unaligned", "lengths", "pos_pair", "rc_pair")


class Blocks(C.Structure):
    _fields_ = [("num_blocks", C.c_uint32), ("data", C.c_void_p * NUM_BLOCK_STREAMS), ("size", C.c_uint64 * NUM_BLOCK_STREAMS),
                ("off", C.c_void_p * NUM_BLOCK_STREAMS), ("order", C.c_void_p), ("num_reads", C.c_uint64)]


class Decoded(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("offsets", C.c_void_p), ("num_reads", C.c_uint64)]


class PackedReads(C.Structure):
    _fields_ = [("reads", C.c_void_p), ("lengths", C.c_void_p), ("num_clean", C.c_uint32), ("num_clean_file1", C.c_uint32),
                ("max_readlen", C.c_uint32), ("n_records", C.c_void_p), ("n_record_bytes", C.c_uint64), ("order_n", C.c_void_p),
                ("num_n", C.c_uint32), ("num_reads", C.c_uint32)]


class Exchanged(C.Structure):
    _fields_ = [("reads", C.c_void_p), ("lengths", C.c_void_p), ("ids", C.c_void_p), ("num_reads", C.c_uint32),
                ("sent_to_peers", C.c_uint64), ("received_from_peers", C.c_uint64)]


class ShardLayout(C.Structure):
    _fields_ = [("rank", C.c_uint32), ("world", C.c_uint32)] + [(n, C.c_uint64) for n in (
        "seq_base", "aligned_before", "noise_before", "num_noise_before", "unaligned_reads_before", "unaligned_bytes_before",
        "total_seq_len", "total_aligned", "total_reads", "total_noise_bytes", "total_num_noise", "total_unaligned_bytes",
        "total_unaligned_len")]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


class Merged(C.Structure):
    _fields_ = [("streams", Streams), ("num_shards", C.c_int), ("shard_seq", C.POINTER(C.c_void_p)),
                ("shard_seq_len", C.POINTER(C.c_uint64)), ("owner", C.c_void_p)]


class Verify(C.Structure):
    _fields_ = [("ok", C.c_int32), ("num_reads", C.c_uint64), ("reads_checked", C.c_uint64), ("base_mismatch_reads", C.c_uint64),
                ("length_mismatch_reads", C.c_uint64), ("bad_order", C.c_uint64), ("num_blocks", C.c_uint64),
                ("block_stream_bytes", C.c_uint64), ("decoded_bases", C.c_uint64)]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


class Stats(C.Structure):
    _fields_ = [("num_chains", C.c_uint32), ("unmatched", C.c_uint32), ("rounds", C.c_uint64),
                ("lost_proposals", C.c_uint64), ("probes_issued", C.c_uint64), ("probes_seq", C.c_uint64),
                ("compares", C.c_uint64), ("gpu_launches", C.c_uint64), ("ms_h2d", C.c_float), ("ms_dict", C.c_float),
                ("ms_chains", C.c_float), ("ms_scatter", C.c_float), ("ms_encode", C.c_float), ("ms_d2h", C.c_float),
                ("ms_total", C.c_float), ("ms_chain_kernel", C.c_float), ("cyc_search", C.c_uint64),
                ("cyc_wait_a", C.c_uint64), ("cyc_commit", C.c_uint64), ("cyc_wait_b", C.c_uint64),
                ("slot_probes", C.c_uint64), ("ms_reblock", C.c_float), ("ms_exchange", C.c_float),
                ("singletons_aligned", C.c_uint32), ("n_reads_aligned", C.c_uint32),
                ("contigs", C.c_uint32), ("contigs_stitched", C.c_uint32)]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


_lib = None


def load():
    """dlopen the CUDA library; raises if it has not been built (python -m spring_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpringB200Error(-2, f"{LIB_PATH} not built; run `python spring_b200/build.py` (needs nvcc)")
        lib = C.CDLL(LIB_PATH)
        lib.spring_b200_version.restype = C.c_char_p
        lib.spring_b200_last_error.restype = C.c_char_p
        lib.spring_b200_last_error.argtypes = [C.c_void_p]
        lib.spring_b200_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.spring_b200_destroy.argtypes = [C.c_void_p]
        lib.spring_b200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        lib.spring_b200_reorder_encode.argtypes = [C.c_void_p, C.POINTER(Input), C.c_uint32, C.POINTER(Streams)]
        lib.spring_b200_reorder_encode_device.argtypes = [C.c_void_p, C.POINTER(Input), C.c_uint32, C.POINTER(Streams)]
        lib.spring_b200_fetch_streams.argtypes = [C.c_void_p, C.POINTER(Streams)]
        lib.spring_b200_build_dictionary.argtypes = [C.c_void_p, C.POINTER(Input), C.c_int, C.c_void_p, C.c_void_p,
                                                     C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.spring_b200_reorder.argtypes = [C.c_void_p, C.POINTER(Input), C.c_uint32, C.POINTER(ReorderOut)]
        lib.spring_b200_reorder_encode_files.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(CP), C.c_uint32]
        lib.spring_b200_bucket_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        lib.spring_b200_write_streams.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(Streams), C.c_int]
        lib.spring_b200_pe_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        lib.spring_b200_reblock_streams.argtypes = [C.c_void_p, C.POINTER(Streams), C.POINTER(CP), C.POINTER(Blocks)]
        lib.spring_b200_reblock_files.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(CP)]
        lib.spring_b200_decode_blocks.argtypes = [C.c_void_p, C.POINTER(Blocks), C.c_void_p, C.c_uint64, C.POINTER(CP),
                                                  C.POINTER(Decoded)]
        lib.spring_b200_pack_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                               C.POINTER(PackedReads)]
        lib.spring_b200_verify_roundtrip.argtypes = [C.c_void_p, C.POINTER(CP), C.POINTER(Verify)]
        lib.spring_b200_comm_unique_id.argtypes = [C.c_void_p]
        lib.spring_b200_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.spring_b200_comm_free.argtypes = [C.c_void_p]
        lib.spring_b200_exchange_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                                   C.POINTER(Exchanged)]
        lib.spring_b200_finalize_shard.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(ShardLayout)]
        lib.spring_b200_merge_shards.argtypes = [C.POINTER(Streams), C.c_int, C.POINTER(Merged)]
        lib.spring_b200_free_merged.argtypes = [C.POINTER(Merged)]
        lib.spring_b200_write_merged.argtypes = [C.c_char_p, C.POINTER(Merged)]
        _lib = lib
    return _lib


def _view(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


@dataclass
class StreamsResult:
    """Host copy of spring_b200_streams (same fields as the oracle's EncodeResult)."""
    seq_packed: np.ndarray
    seq_len: int
    pos: np.ndarray
    noise: np.ndarray
    noisepos: np.ndarray
    rc: np.ndarray
    order: np.ndarray
    lengths: np.ndarray
    unaligned: np.ndarray
    unaligned_len: int
    num_aligned: int
    matched_s: int
    matched_N: int

    @property
    def seq(self) -> np.ndarray:
        """ASCII consensus (unpacked from 2 bits/base A0 C1 G2 T3)."""
        b = self.seq_packed
        codes = np.stack([(b >> (2 * j)) & 3 for j in range(4)], axis=1).reshape(-1)[: self.seq_len]
        return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


@dataclass
class BlocksResult:
    """Host copy of spring_b200_blocks: data[s] = stream s with its blocks concatenated, off[s] the
    byte offsets of the blocks (same layout as the oracle's BlockStreams)."""
    num_blocks: int
    data: dict
    off: dict
    order: "np.ndarray | None"

    def block(self, stream: str, b: int) -> bytes:
        o = self.off[stream]
        return self.data[stream][int(o[b]): int(o[b + 1])].tobytes()


class Context:
    """One per GPU.  `stream`: raw cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream) or None."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._lib = load()
        self._h = C.c_void_p()
        rc = self._lib.spring_b200_create(device, None, C.byref(self._h))
        if rc != 0:
            raise SpringB200Error(rc, self._lib.spring_b200_last_error(None).decode())
        if stream is not None:  # 0 is a valid handle: the legacy default stream
            self._lib.spring_b200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
            self._check(self._lib.spring_b200_set_stream(self._h, C.c_void_p(stream)))
        self._keep = []

    def close(self):
        if self._h:
            self._lib.spring_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise SpringB200Error(rc, self._lib.spring_b200_last_error(self._h).decode())

    def set_schedule(self, deterministic: bool) -> None:
        """True: round-synchronous chains (reproducible output); False (default): free-running chains."""
        self._lib.spring_b200_set_schedule.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.spring_b200_set_schedule(self._h, 1 if deterministic else 0))

    def set_stitch(self, mode: int) -> None:
        """-1: automatic (default), 0: off (encoder == the reference's encoder on the same reorder stream), 1: on."""
        self._lib.spring_b200_set_stitch.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.spring_b200_set_stitch(self._h, int(mode)))

    def set_chain_stats(self, on: bool) -> None:
        """True: the free-running chain kernel also counts lookups / compares (probes_seq, compares, ... in stats())."""
        self._lib.spring_b200_set_chain_stats.argtypes = [C.c_void_p, C.c_int]
        self._check(self._lib.spring_b200_set_chain_stats(self._h, 1 if on else 0))

    def stats(self) -> dict:
        s = Stats()
        self._check(self._lib.spring_b200_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def make_input(self, reads_ptr: int, lengths_ptr: int, num_clean: int, max_readlen: int, n_records: bytes = b"",
                   order_n: np.ndarray | None = None, num_reads: int | None = None) -> Input:
        order_n = np.zeros(0, np.uint32) if order_n is None else np.ascontiguousarray(order_n, dtype=np.uint32)
        nbuf = np.frombuffer(n_records, dtype=np.uint8).copy() if n_records else np.zeros(0, np.uint8)
        self._keep = [order_n, nbuf]
        inp = Input()
        inp.reads = reads_ptr
        inp.lengths = lengths_ptr
        inp.num_clean = num_clean
        inp.max_readlen = max_readlen
        inp.n_records = nbuf.ctypes.data if len(nbuf) else None
        inp.n_record_bytes = len(nbuf)
        inp.order_n = order_n.ctypes.data if len(order_n) else None
        inp.num_n = len(order_n)
        inp.num_reads = num_reads if num_reads is not None else num_clean + len(order_n)
        return inp

    def _input_from_numpy(self, packed, lengths, max_readlen, n_records, order_n, num_reads):
        packed = np.ascontiguousarray(packed, dtype=np.uint64)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint16)
        inp = self.make_input(packed.ctypes.data if packed.size else None, lengths.ctypes.data if lengths.size else None,
                              len(lengths), max_readlen, n_records, order_n, num_reads)
        self._keep += [packed, lengths]
        return inp

    @staticmethod
    def _streams_to_result(s: Streams) -> StreamsResult:
        return StreamsResult(
            _view(s.seq_packed, (s.seq_len + 3) // 4, np.uint8).copy(), s.seq_len,
            _view(s.pos, s.num_aligned, np.uint64).copy(), _view(s.noise, s.noise_bytes, np.uint8).copy(),
            _view(s.noisepos, s.num_noise, np.uint16).copy(), _view(s.rev, s.num_aligned, np.uint8).copy(),
            _view(s.order, s.num_reads, np.uint32).copy(), _view(s.lengths, s.num_reads, np.uint16).copy(),
            _view(s.unaligned, s.unaligned_bytes, np.uint8).copy(), s.unaligned_len, s.num_aligned,
            s.singletons_aligned, s.n_reads_aligned)

    # ---- the hot path ------------------------------------------------------------------------
    def reorder_encode(self, packed, lengths, max_readlen, n_records=b"", order_n=None, num_reads=None,
                       num_chains: int = 0) -> StreamsResult:
        """Host numpy in, host numpy out (H2D + kernels + D2H)."""
        inp = self._input_from_numpy(packed, lengths, max_readlen, n_records, order_n, num_reads)
        s = Streams()
        self._check(self._lib.spring_b200_reorder_encode(self._h, C.byref(inp), num_chains, C.byref(s)))
        return self._streams_to_result(s)

    def reorder_encode_raw(self, inp: Input, num_chains: int = 0, device: bool = False) -> Streams:
        """No copies of the result: returns the raw struct (pointers owned by the context)."""
        s = Streams()
        fn = self._lib.spring_b200_reorder_encode_device if device else self._lib.spring_b200_reorder_encode
        self._check(fn(self._h, C.byref(inp), num_chains, C.byref(s)))
        return s

    def fetch_streams_raw(self) -> Streams:
        s = Streams()
        self._check(self._lib.spring_b200_fetch_streams(self._h, C.byref(s)))
        return s

    def fetch_streams(self) -> StreamsResult:
        s = Streams()
        self._check(self._lib.spring_b200_fetch_streams(self._h, C.byref(s)))
        return self._streams_to_result(s)

    # ---- stages -------------------------------------------------------------------------------
    def build_dictionary(self, packed, lengths, max_readlen, which: int):
        inp = self._input_from_numpy(packed, lengths, max_readlen, b"", None, None)
        n = len(lengths)
        keys = np.zeros(max(n, 1), np.uint64)
        bs = np.zeros(n + 1, np.uint32)
        rid = np.zeros(max(n, 1), np.uint32)
        nk, dn = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.spring_b200_build_dictionary(self._h, C.byref(inp), which, keys.ctypes.data, bs.ctypes.data,
                                                           rid.ctypes.data, C.byref(nk), C.byref(dn)))
        return keys[: nk.value], bs[: nk.value + 1], rid[: dn.value]

    def reorder(self, packed, lengths, max_readlen, num_chains: int = 0):
        inp = self._input_from_numpy(packed, lengths, max_readlen, b"", None, None)
        o = ReorderOut()
        self._check(self._lib.spring_b200_reorder(self._h, C.byref(inp), num_chains, C.byref(o)))
        return (_view(o.order, o.num, np.uint32).copy(), _view(o.flag, o.num, np.uint8).copy(),
                _view(o.pos, o.num, np.int64).copy(), _view(o.rev, o.num, np.uint8).copy(),
                _view(o.singleton_order, o.num_singletons, np.uint32).copy())

    def bucket_reads(self, reads_ptr: int, lengths_ptr: int, num_reads: int, max_readlen: int, num_buckets: int, out_ptr: int) -> None:
        """Device pointers; out: uint32[num_reads] owner bucket of every read (multi-GPU partitioning)."""
        self._check(self._lib.spring_b200_bucket_reads(self._h, reads_ptr, lengths_ptr, num_reads, max_readlen, num_buckets, out_ptr))

    def fetch_reorder(self):
        """(order, flag, pos, rev, singleton_order) the encoder of the last reorder_encode call consumed."""
        o = ReorderOut()
        self._lib.spring_b200_fetch_reorder.argtypes = [C.c_void_p, C.POINTER(ReorderOut)]
        self._check(self._lib.spring_b200_fetch_reorder(self._h, C.byref(o)))
        return (_view(o.order, o.num, np.uint32).copy(), _view(o.flag, o.num, np.uint8).copy(),
                _view(o.pos, o.num, np.int64).copy(), _view(o.rev, o.num, np.uint8).copy(),
                _view(o.singleton_order, o.num_singletons, np.uint32).copy())

    # ---- the stage before the dictionaries -----------------------------------
    def pack_reads(self, bases: np.ndarray, offsets: np.ndarray, num_reads_file1: int | None = None, keep_on_device: bool = False):
        """preprocess's read path: N split + 2-bit / 4-bit packing.  bases uint8 (sequence lines
        concatenated), offsets uint64[n + 1].  Returns a dict with the fields of spring_b200_input
        (numpy arrays; `reads` / `lengths` are raw device pointers when keep_on_device)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        p = PackedReads()
        self._keep = [bases, offsets]
        self._check(self._lib.spring_b200_pack_reads(self._h, bases.ctypes.data if bases.size else None, offsets.ctypes.data, n,
                                                     n if num_reads_file1 is None else num_reads_file1, 1 if keep_on_device else 0,
                                                     C.byref(p)))
        w = (2 * max(p.max_readlen, 1) - 1) // 64 + 1
        res = dict(num_clean=p.num_clean, num_clean_file1=p.num_clean_file1, max_readlen=p.max_readlen, num_n=p.num_n,
                   num_reads=p.num_reads, n_records=_view(p.n_records, p.n_record_bytes, np.uint8).tobytes(),
                   order_n=_view(p.order_n, p.num_n, np.uint32).copy())
        if keep_on_device:
            res["reads_ptr"], res["lengths_ptr"] = p.reads, p.lengths
        else:
            res["packed"] = _view(p.reads, p.num_clean * w, np.uint64).copy().reshape(p.num_clean, w)
            res["lengths"] = _view(p.lengths, p.num_clean, np.uint16).copy()
        return res

    # ---- the stages after the encoder ----------------------------------------------
    def pe_encode(self, order: np.ndarray) -> np.ndarray:
        """pe_encode (src/pe_encode.cpp:24-84) on read_order.bin's contents."""
        o = np.ascontiguousarray(order, dtype=np.uint32)
        out = np.empty_like(o)
        self._check(self._lib.spring_b200_pe_encode(self._h, o.ctypes.data if len(o) else None, len(o),
                                                    out.ctypes.data if len(o) else None))
        return out

    def reblock_streams(self, cp: CP, streams: "StreamsResult | None" = None) -> "BlocksResult":
        """The re-blocking of reorder_compress_streams (+ pe_encode for -r paired input).  streams=None:
        use the encoder streams the last reorder_encode* call left in HBM."""
        sp = None
        if streams is not None:
            arrs = [np.ascontiguousarray(a, dtype=dt) for a, dt in (
                (streams.pos, np.uint64), (streams.noise, np.uint8), (streams.noisepos, np.uint16), (streams.rc, np.uint8),
                (streams.order, np.uint32), (streams.lengths, np.uint16), (streams.unaligned, np.uint8))]
            self._keep = arrs
            s = Streams()
            ptr = lambda a: a.ctypes.data if a.size else None
            s.pos, s.noise, s.noisepos, s.rev, s.order, s.lengths, s.unaligned = (ptr(a) for a in arrs)
            s.noise_bytes, s.num_noise, s.unaligned_bytes = len(arrs[1]), len(arrs[2]), len(arrs[6])
            s.unaligned_len, s.num_aligned, s.num_reads = int(streams.unaligned_len), int(streams.num_aligned), len(arrs[5])
            sp = C.byref(s)
        b = Blocks()
        self._check(self._lib.spring_b200_reblock_streams(self._h, sp, C.byref(cp), C.byref(b)))
        nb = b.num_blocks
        data = {n: _view(b.data[i], b.size[i], np.uint8).copy() for i, n in enumerate(BLOCK_STREAMS)}
        off = {n: _view(b.off[i], nb + 1, np.uint64).copy() for i, n in enumerate(BLOCK_STREAMS)}
        order = _view(b.order, b.num_reads, np.uint32).copy() if b.order else None
        return BlocksResult(nb, data, off, order)

    def reblock_streams_raw(self, cp: CP) -> Blocks:
        """Device-resident streams of the last reorder_encode* call; no copies of the result (pointers
        owned by the context)."""
        b = Blocks()
        self._check(self._lib.spring_b200_reblock_streams(self._h, None, C.byref(cp), C.byref(b)))
        return b

    def decode_blocks(self, blocks, seq_packed: np.ndarray, seq_len: int, cp: CP):
        """decompress_short's block decode (src/decompress.cpp:230-320): blocks in the BlocksResult /
        oracle BlockStreams layout + packed consensus -> (bases uint8, offsets uint64[n + 1]); file 1's
        reads first, then file 2's."""
        b = Blocks()
        b.num_blocks = blocks.num_blocks
        keep = []
        for i, name in enumerate(BLOCK_STREAMS):
            d = np.ascontiguousarray(blocks.data[name], dtype=np.uint8)
            o = np.ascontiguousarray(blocks.off[name], dtype=np.uint64)
            keep += [d, o]
            b.data[i] = d.ctypes.data if d.size else None
            b.size[i] = d.size
            b.off[i] = o.ctypes.data
        sp = np.ascontiguousarray(seq_packed, dtype=np.uint8)
        keep.append(sp)
        self._keep = keep
        out = Decoded()
        self._check(self._lib.spring_b200_decode_blocks(self._h, C.byref(b), sp.ctypes.data if sp.size else None, seq_len,
                                                        C.byref(cp), C.byref(out)))
        offs = _view(out.offsets, out.num_reads + 1, np.uint64).copy()
        return _view(out.bases, int(offs[-1]) if len(offs) else 0, np.uint8).copy(), offs

    # ---- multi-GPU ------------------------------------------------------------------
    def comm_init(self, comm_id: bytes, rank: int, world: int) -> None:
        """Join the job's NCCL communicator (comm_id from capi.comm_unique_id() on one rank)."""
        buf = (C.c_uint8 * 128).from_buffer_copy(comm_id)
        self._check(self._lib.spring_b200_comm_init(self._h, buf, rank, world))

    def comm_free(self) -> None:
        self._check(self._lib.spring_b200_comm_free(self._h))

    def exchange_reads(self, reads_ptr: int, lengths_ptr: int, ids_ptr: int, num_reads: int, max_readlen: int) -> Exchanged:
        """Device pointers in, device pointers out (owned by the context): the reads this rank owns."""
        x = Exchanged()
        self._check(self._lib.spring_b200_exchange_reads(self._h, reads_ptr, lengths_ptr, ids_ptr, num_reads, max_readlen, C.byref(x)))
        return x

    def finalize_shard(self, ids_ptr: int, num_owned: int, n_ids: np.ndarray | None = None) -> dict:
        """Absolute positions + global ids for the streams of the last reorder_encode_raw(device=True) call."""
        n_ids = np.zeros(0, np.uint32) if n_ids is None else np.ascontiguousarray(n_ids, dtype=np.uint32)
        lay = ShardLayout()
        self._check(self._lib.spring_b200_finalize_shard(self._h, ids_ptr, num_owned, n_ids.ctypes.data if len(n_ids) else None,
                                                         len(n_ids), C.byref(lay)))
        return lay.as_dict()

    def verify_roundtrip(self, cp: CP) -> dict:
        """Re-block -> block decode -> exact compare with the input of the last reorder_encode* call, all in HBM."""
        v = Verify()
        self._check(self._lib.spring_b200_verify_roundtrip(self._h, C.byref(cp), C.byref(v)))
        return v.as_dict()

    def reblock_files(self, temp_dir: str, cp: CP) -> None:
        self._check(self._lib.spring_b200_reblock_files(self._h, temp_dir.encode(), C.byref(cp)))

    # ---- files ---------------------------------------------------------------------------------
    def reorder_encode_files(self, temp_dir: str, cp: CP, num_chains: int = 0) -> None:
        self._check(self._lib.spring_b200_reorder_encode_files(self._h, temp_dir.encode(), C.byref(cp), num_chains))


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (one rank makes it, the host hands it to the others)."""
    buf = (C.c_uint8 * 128)()
    rc = load().spring_b200_comm_unique_id(buf)
    if rc != 0:
        raise SpringB200Error(rc, load().spring_b200_last_error(None).decode())
    return bytes(buf)


def merge_shards(parts: list) -> tuple["StreamsResult", list]:
    """spring_b200_merge_shards on finalized host shards (StreamsResult-like objects, rank order): the whole job's
    streams (aligned pieces of every shard, then the unaligned ones) and the per-shard (packed consensus, length)."""
    lib = load()
    n = len(parts)
    arr = (Streams * n)()
    keep = []
    for i, s in enumerate(parts):
        cols = [np.ascontiguousarray(a, dtype=dt) for a, dt in (
            (s.seq_packed, np.uint8), (s.pos, np.uint64), (s.noise, np.uint8), (s.noisepos, np.uint16), (s.rc, np.uint8),
            (s.order, np.uint32), (s.lengths, np.uint16), (s.unaligned, np.uint8))]
        keep.append(cols)
        ptr = lambda a: a.ctypes.data if a.size else None
        t = arr[i]
        t.seq_packed, t.pos, t.noise, t.noisepos, t.rev, t.order, t.lengths, t.unaligned = (ptr(a) for a in cols)
        t.seq_len, t.noise_bytes, t.num_noise = int(s.seq_len), len(cols[2]), len(cols[3])
        t.unaligned_bytes, t.unaligned_len = len(cols[7]), int(s.unaligned_len)
        t.num_aligned, t.num_reads = int(s.num_aligned), len(cols[5])
        t.singletons_aligned, t.n_reads_aligned = int(getattr(s, "matched_s", 0)), int(getattr(s, "matched_N", 0))
    m = Merged()
    rc = lib.spring_b200_merge_shards(arr, n, C.byref(m))
    if rc != 0:
        raise SpringB200Error(rc, lib.spring_b200_last_error(None).decode())
    try:
        res = Context._streams_to_result(m.streams)
        shards = [(keep[i][0], int(m.shard_seq_len[i])) for i in range(n)]
    finally:
        lib.spring_b200_free_merged(C.byref(m))
    return res, shards
