"""Seeded synthetic read sets for the BASELINE.json configs.

Genome i.i.d. uniform ACGT, read starts uniform, strand RC with p = 0.5, substitutions either
uniform (``sub_rate``) or rising linearly along the read ("illumina": 0.1 % -> 1 %), optional
variable lengths U[lo, hi], optional reads with 1-3 ``N``s, optional pairing (insert ~U[200, 500],
mate 2 = reverse strand of the fragment end).

Everything is torch so the same code makes a 20 k-read CPU test input and a 10 M-read bench input
on the GPU; this is harness plumbing (there is no network for real FASTQ), not the product.
Base codes follow the reference's 2-bit layout A0 G1 C2 T3 (reorder.h:97-106); 4 = N.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import dnaio


@dataclass
class ReadSet:
    """codes: uint8[N, L] (0..3, 4 = N; undefined beyond lengths[i]); file 2 mates follow file 1."""
    codes: torch.Tensor
    lengths: torch.Tensor
    max_readlen: int
    paired: bool

    @property
    def num_reads(self) -> int:
        return int(self.codes.shape[0])


def generate(num_reads: int, read_len: int = 150, genome_len: int | None = None, seed: int = 3,
             sub_rate: float = 0.005, error_model: str = "uniform", var_len: tuple[int, int] | None = None,
             paired: bool = False, n_frac: float = 0.0, device: str | torch.device = "cpu",
             chunk: int = 1 << 20, read_seed: int | None = None, repeats: bool = False) -> ReadSet:
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = read_len if var_len is None else var_len[1]
    if genome_len is None:
        genome_len = max(num_reads * L // 30, 4 * L + 600)       # 30x coverage
    genome = torch.randint(0, 4, (genome_len,), dtype=torch.uint8, device=dev, generator=g)
    hot = None
    if repeats:
        # repeat-rich genome: interspersed copies of 300-3000 bp segments (1 % divergence), tandem arrays of short
        # units, one low-complexity (poly-A / poly-G) stretch -- and coverage spikes: 5 % of the reads start inside a
        # few 200 bp hot spots (amplicon-like), so single columns stack tens of thousands of reads
        n_rep = max(genome_len // 20000, 4)
        for _ in range(n_rep):
            ln = int(torch.randint(300, 3001, (1,), generator=g, device=dev))
            src = int(torch.randint(0, genome_len - ln, (1,), generator=g, device=dev))
            dst = int(torch.randint(0, genome_len - ln, (1,), generator=g, device=dev))
            seg = genome[src:src + ln].clone()
            mut = torch.rand((ln,), device=dev, generator=g) < 0.01
            seg = torch.where(mut, (seg + 1) & 3, seg)
            genome[dst:dst + ln] = seg
        for _ in range(max(genome_len // 200000, 2)):
            unit = int(torch.randint(2, 40, (1,), generator=g, device=dev))
            reps = int(torch.randint(10, 200, (1,), generator=g, device=dev))
            at = int(torch.randint(0, max(genome_len - unit * reps, 1), (1,), generator=g, device=dev))
            u = torch.randint(0, 4, (unit,), dtype=torch.uint8, device=dev, generator=g)
            ln = min(unit * reps, genome_len - at)
            genome[at:at + ln] = u.repeat(reps)[:ln]
        at = int(torch.randint(0, max(genome_len - 600, 1), (1,), generator=g, device=dev))
        genome[at:at + 300] = 0
        genome[at + 300:at + 600] = 1
        hot = torch.randint(0, max(genome_len - 3000, 1), (4,), device=dev, generator=g)
    if read_seed is not None:  # same genome, different reads (one block of a larger read set)
        g.manual_seed(read_seed)
    n_frag = num_reads // 2 if paired else num_reads
    out_codes = torch.empty((num_reads, L), dtype=torch.uint8, device=dev)
    out_len = torch.empty((num_reads,), dtype=torch.int32, device=dev)
    ar = torch.arange(L, device=dev)
    if error_model == "illumina":
        perr = (0.001 + (0.01 - 0.001) * ar.float() / max(L - 1, 1))[None, :]
    else:
        perr = torch.full((1, L), float(sub_rate), device=dev)

    def sample(starts: torch.Tensor, rc: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
        """reads of length lens starting at genome[starts] on the forward strand, or (rc) the
        reverse complement of genome[starts : starts+len]."""
        m = starts.shape[0]
        idx_f = starts[:, None] + ar[None, :]
        idx_r = starts[:, None] + (lens[:, None] - 1 - ar[None, :])
        idx = torch.where(rc[:, None], idx_r, idx_f).clamp_(0, genome_len - 1)
        c = genome[idx]
        c = torch.where(rc[:, None], 3 - c, c)
        err = torch.rand((m, L), device=dev, generator=g) < perr
        delta = torch.randint(1, 4, (m, L), dtype=torch.uint8, device=dev, generator=g)
        c = torch.where(err, (c + delta) & 3, c)
        if n_frac > 0:
            has_n = torch.rand((m,), device=dev, generator=g) < n_frac
            for _ in range(3):
                p = (torch.rand((m,), device=dev, generator=g) * lens).long().clamp_(max=L - 1)
                hit = has_n & (torch.rand((m,), device=dev, generator=g) < 0.67)
                rows = torch.nonzero(hit).squeeze(1)
                c[rows, p[rows]] = 4
        return c

    for lo in range(0, n_frag, chunk):
        hi = min(lo + chunk, n_frag)
        m = hi - lo
        if var_len is None:
            lens = torch.full((m,), L, dtype=torch.long, device=dev)
        else:
            lens = torch.randint(var_len[0], var_len[1] + 1, (m,), device=dev, generator=g)
        rc = torch.rand((m,), device=dev, generator=g) < 0.5
        if not paired:
            starts = torch.randint(0, genome_len - L + 1, (m,), device=dev, generator=g)
            if hot is not None:
                spike = torch.rand((m,), device=dev, generator=g) < 0.05
                hs = hot[torch.randint(0, len(hot), (m,), device=dev, generator=g)] + torch.randint(0, 200, (m,), device=dev, generator=g)
                starts = torch.where(spike, hs.clamp_(max=genome_len - L), starts)
            out_codes[lo:hi] = sample(starts, rc, lens)
            out_len[lo:hi] = lens.int()
        else:
            insert = torch.randint(200, 501, (m,), device=dev, generator=g).clamp_(min=L)
            # exact integer starts: a float32 uniform has 2^24 distinct values, a 30 bp grid on a 500 Mbp genome
            fstart = torch.randint(0, max(genome_len - 500 - L, 1), (m,), device=dev, generator=g)
            lens2 = lens if var_len is None else torch.randint(var_len[0], var_len[1] + 1, (m,), device=dev, generator=g)
            # fragment on strand rc: mate 1 reads the fragment start, mate 2 the opposite strand of its end
            s1 = torch.where(rc, fstart + insert - lens, fstart)
            s2 = torch.where(rc, fstart, fstart + insert - lens2)
            out_codes[lo:hi] = sample(s1, rc, lens)
            out_codes[n_frag + lo:n_frag + hi] = sample(s2, ~rc, lens2)
            out_len[lo:hi] = lens.int()
            out_len[n_frag + lo:n_frag + hi] = lens2.int()
    return ReadSet(out_codes, out_len, L, paired)


def pack_reads(codes: torch.Tensor, lengths: torch.Tensor, max_readlen: int) -> torch.Tensor:
    """uint8[N, L] codes (0..3) -> int64[N, W] in the reference's bitset layout (bit pattern of
    the uint64 words; base j at bits 2j, 2j+1; zero beyond the read's length)."""
    n, L = codes.shape
    w = dnaio.words_per_read(max_readlen)
    dev = codes.device
    out = torch.empty((n, w), dtype=torch.int64, device=dev)
    shifts = (2 * torch.arange(32, device=dev, dtype=torch.int64))[None, None, :]
    step = 1 << 18
    for lo in range(0, n, step):
        hi = min(lo + step, n)
        full = torch.zeros((hi - lo, w * 32), dtype=torch.int64, device=dev)
        valid = torch.arange(L, device=dev)[None, :] < lengths[lo:hi, None]
        full[:, :L] = torch.where(valid, codes[lo:hi].long() & 3, 0)
        out[lo:hi] = (full.view(hi - lo, w, 32) << shifts).sum(dim=2)
    return out


@dataclass
class HotpathInput:
    """What preprocess hands to the hot path (preprocess.cpp:296-403), in memory."""
    packed: np.ndarray          # uint64[N_clean, W]
    lengths: np.ndarray         # uint16[N_clean]
    n_seqs: list                # N reads as ASCII
    order_n: np.ndarray         # uint32 original indices of N reads (read_order_N.bin)
    num_reads: int              # cp.num_reads
    num_clean: tuple            # cp.num_reads_clean[2]
    max_readlen: int
    paired: bool

    @property
    def n_records(self) -> bytes:
        return dnaio.write_dnaN_records(self.n_seqs)


def to_hotpath_input(rs: ReadSet) -> HotpathInput:
    """Split clean / N reads exactly as preprocess does: clean reads keep file order (file 1 then
    file 2), N reads go to input_N.dna with their original index (file-2 indices offset by the
    file-1 read count, preprocess.cpp:300-301, :373-378)."""
    codes, lens = rs.codes, rs.lengths
    valid = torch.arange(codes.shape[1], device=codes.device)[None, :] < lens[:, None]
    has_n = ((codes == 4) & valid).any(dim=1)
    clean_idx = torch.nonzero(~has_n).squeeze(1)
    packed = pack_reads(codes[clean_idx], lens[clean_idx], rs.max_readlen).cpu().numpy().view(np.uint64)
    lengths = lens[clean_idx].cpu().numpy().astype(np.uint16)
    n_idx = torch.nonzero(has_n).squeeze(1).cpu().numpy()
    ccpu = codes[has_n].cpu().numpy()
    lcpu = lens[has_n].cpu().numpy()
    n_seqs = [dnaio.CODE4CHAR[ccpu[i, : lcpu[i]]].tobytes() for i in range(len(n_idx))]
    half = rs.num_reads // 2 if rs.paired else rs.num_reads
    c0 = int((~has_n[:half]).sum())
    return HotpathInput(packed, lengths, n_seqs, n_idx.astype(np.uint32), rs.num_reads,
                        (c0, len(lengths) - c0), rs.max_readlen, rs.paired)


def write_fastq(rs: ReadSet, path1: str, path2: str | None = None, quality: bool = True) -> None:
    """FASTQ text with ids ``@r<k>/<mate>``; constant-ish qualities."""
    codes = rs.codes.cpu().numpy()
    lens = rs.lengths.cpu().numpy()
    half = rs.num_reads // 2 if rs.paired else rs.num_reads
    for mate, path, lo, hi in ((1, path1, 0, half), (2, path2, half, rs.num_reads)):
        if path is None or lo == hi:
            continue
        with open(path, "wb") as f:
            for i in range(lo, hi):
                s = dnaio.CODE4CHAR[codes[i, : lens[i]]].tobytes()
                f.write(b"@r%d/%d\n" % (i - lo, mate))
                f.write(s + b"\n+\n")
                f.write(b"I" * len(s) + b"\n")


@dataclass
class DeviceInput:
    """The hot path's input with the clean reads resident on the GPU (what spring_b200_pack_reads with
    keep_on_device leaves), built in chunks so that a 100 M-read set never needs more than its own size."""
    reads: torch.Tensor         # int64[N_clean, W] (bit pattern of the uint64 rows)
    lengths: torch.Tensor       # int16[N_clean]
    n_seqs: list                # reads with N, ASCII
    order_n: np.ndarray         # uint32 original index of every N read, ascending
    num_reads: int
    num_clean: tuple            # cp.num_reads_clean[2]
    max_readlen: int
    paired: bool

    @property
    def n_records(self) -> bytes:
        return dnaio.write_dnaN_records(self.n_seqs)


def to_device_input(rs: ReadSet, chunk: int = 1 << 20) -> DeviceInput:
    """Same split as to_hotpath_input (preprocess.cpp:293-304, :364-378), on rs.codes' device."""
    codes, lens = rs.codes, rs.lengths
    dev, n, L = codes.device, rs.num_reads, codes.shape[1]
    w = dnaio.words_per_read(rs.max_readlen)
    ar = torch.arange(L, device=dev)[None, :]
    has_n = torch.empty((n,), dtype=torch.bool, device=dev)
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        has_n[lo:hi] = ((codes[lo:hi] == 4) & (ar < lens[lo:hi, None])).any(dim=1)
    n_clean = int((~has_n).sum())
    reads = torch.empty((n_clean, w), dtype=torch.int64, device=dev)
    lengths = torch.empty((n_clean,), dtype=torch.int16, device=dev)
    at = 0
    for lo in range(0, n, chunk):
        hi = min(lo + chunk, n)
        keep = ~has_n[lo:hi]
        c, l = codes[lo:hi][keep], lens[lo:hi][keep]
        m = int(c.shape[0])
        reads[at:at + m] = pack_reads(c, l, rs.max_readlen)
        lengths[at:at + m] = l.to(torch.int16)
        at += m
    n_idx = torch.nonzero(has_n).squeeze(1)
    ccpu, lcpu = codes[n_idx].cpu().numpy(), lens[n_idx].cpu().numpy()
    n_seqs = [dnaio.CODE4CHAR[ccpu[i, : lcpu[i]]].tobytes() for i in range(len(lcpu))]
    half = n // 2 if rs.paired else n
    c0 = int((~has_n[:half]).sum())
    return DeviceInput(reads, lengths, n_seqs, n_idx.cpu().numpy().astype(np.uint32), n, (c0, n_clean - c0), rs.max_readlen, rs.paired)
