// decode.cu -- decompress_short's block decode on the GPU: the nine per-block streams
// + the consensus -> every read as ASCII.  Reference: src/decompress.cpp:230-320 (dec_noise :664-685,
// reverse_complement util.cpp:376-381).  The mirror image of reblock.cu; the reference decompressor itself
// stays the parity oracle of the compressor (oracle/reblock_oracle.c:orc_decode_blocks restates this
// stage and is pinned against `spring -d`).
//
// The reference walks a block sequentially, pulling from nine streams.  Here:
//   1. k_dec_units   : thread per unit (read or pair): what it consumes from every stream, from its flag
//   2. one scan      : where each unit's orientation chars, noise lines, pair entries and unaligned text start
//   3. k_dec_pos     : WARP PER BLOCK over the position stream, the one truly sequential stream of -r mode
//                      (u16 deltas, 65535 = escape to an absolute u64): 32 units per step, offsets by warp
//                      prefix sum, escapes resolved left to right with ballots, deltas turned into positions
//                      by a segmented warp scan.  Order-preserving mode (all u64) needs no such pass.
//   4. newline positions of the noise stream (stream compaction)
//   5. k_dec_reads   : thread per read: consensus window -> ASCII, noise substitutions, reverse complement
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {
namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;
static inline uint32_t grid_for(uint64_t n, int block) { return (uint32_t)((n + block - 1) / block); }

struct DecOff {  // per unit: entries consumed before it (exclusive scan)
  unsigned long long unal, pos;  // bytes of unaligned text; bytes of the position stream (order-preserving mode only)
  uint32_t rc, lines, pair, pad;
};
struct DecAdd {
  __device__ DecOff operator()(const DecOff &a, const DecOff &b) const {
    DecOff r;
    r.unal = a.unal + b.unal; r.pos = a.pos + b.pos; r.rc = a.rc + b.rc; r.lines = a.lines + b.lines; r.pair = a.pair + b.pair; r.pad = 0;
    return r;
  }
};
struct IsNewline {
  const uint8_t *noise;
  __device__ bool operator()(uint32_t p) const { return noise[p] == '\n'; }
};

struct DecArgs {
  const uint8_t *flag; const uint8_t *pos; const uint8_t *noise; const uint8_t *noisepos; const uint8_t *rc;
  const uint8_t *unal; const uint8_t *len; const uint8_t *pos_pair; const uint8_t *rc_pair;
  const unsigned long long *boff;  // [9][nb + 1] block byte offsets
  uint32_t nb, units, block;
  int paired, preserve;
  const uint8_t *seq_packed; unsigned long long seq_len;
  uint64_t sizes[RB_NSTREAMS];
  DecOff *off;                     // [units + 1]
  unsigned long long *pos1, *pos2; // [units] decoded positions of read 1 / of an independently coded read 2
  const uint32_t *nl_pos; uint32_t num_lines;
  const unsigned long long *out_off;  // [num_reads + 1]
  uint8_t *out; int *err;
};

__device__ __forceinline__ uint16_t ld_u16(const uint8_t *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
__device__ __forceinline__ unsigned long long ld_u64(const uint8_t *p) {
  unsigned long long v = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) v |= (unsigned long long)p[k] << (8 * k);
  return v;
}
__device__ __forceinline__ bool aligned1(uint8_t f) { return !(f == '2' || f == '4'); }               // decompress.cpp:234
__device__ __forceinline__ bool aligned2(uint8_t f) { return !(f == '2' || f == '3'); }               // :286
__device__ __forceinline__ bool own_pos2(uint8_t f) { return f == '1' || f == '4'; }                  // :290

__global__ void k_dec_lengths(DecArgs a, unsigned long long *len64) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.paired ? 2 * a.units : a.units;
  if (i > n) return;
  if (i == n) { len64[i] = 0; return; }
  // slot i: file 1 is units [0, U), file 2 follows; read_lengths holds the two mates interleaved (:231,:287)
  const uint32_t u = i < a.units ? i : i - a.units;
  const size_t e = a.paired ? 2 * (size_t)u + (i < a.units ? 0 : 1) : u;
  len64[i] = ld_u16(a.len + 2 * e);
}

__global__ void k_dec_units(DecArgs a, const unsigned long long *__restrict__ out_off) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.units) return;
  DecOff u{};
  if (i < a.units) {
    const uint8_t f = a.flag[i];
    if (f < '0' || f > '4' || (!a.paired && f != '0' && f != '2')) *a.err = 1;
    const bool a1 = aligned1(f), a2 = a.paired && aligned2(f);
    const unsigned len1 = (unsigned)(out_off[i + 1] - out_off[i]);
    const unsigned len2 = a.paired ? (unsigned)(out_off[a.units + i + 1] - out_off[a.units + i]) : 0;
    u.rc = (a1 ? 1u : 0u) + ((a2 && own_pos2(f)) ? 1u : 0u);
    u.lines = (a1 ? 1u : 0u) + (a2 ? 1u : 0u);
    u.pair = (a.paired && f == '0') ? 1u : 0u;
    u.unal = (a1 ? 0u : len1) + ((a.paired && !a2) ? len2 : 0u);
    u.pos = a.preserve ? 8ull * u.rc : 0ull;
  }
  a.off[i] = u;
}

// every block must start where the scan says it does: otherwise the streams do not belong together
__global__ void k_dec_check_blocks(DecArgs a) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > a.nb) return;
  const unsigned long long i = (unsigned long long)b * a.block < a.units ? (unsigned long long)b * a.block : a.units;
  const DecOff o = a.off[i];
  const size_t st = (size_t)a.nb + 1;
  bool ok = a.boff[RB_FLAG * st + b] == i && a.boff[RB_LENGTHS * st + b] == (a.paired ? 4ull : 2ull) * i &&
            a.boff[RB_RC * st + b] == o.rc && a.boff[RB_UNALIGNED * st + b] == o.unal &&
            a.boff[RB_POS_PAIR * st + b] == 2ull * o.pair && a.boff[RB_RC_PAIR * st + b] == o.pair;
  if (a.preserve) ok = ok && a.boff[RB_POS * st + b] == o.pos;
  if (b == a.nb) ok = ok && o.rc == a.sizes[RB_RC] && o.unal == a.sizes[RB_UNALIGNED] && o.lines == a.num_lines &&
                      2ull * o.pair == a.sizes[RB_POS_PAIR] && o.pair == a.sizes[RB_RC_PAIR];
  if (!ok) *a.err = 2;
}

// Position stream of one block, -r mode (decompress.cpp:236-254, :290-293): the first aligned read 1 of the
// block is an absolute u64; every later one a u16 delta to the previous, 65535 announcing an absolute u64;
// an independently coded read 2 (flags 1, 4) adds a u64 right behind its mate's entry.
__global__ void k_dec_pos(DecArgs a) {
  const uint32_t b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= a.nb) return;
  const size_t st = (size_t)a.nb + 1;
  const uint32_t start = b * a.block, end = min(a.units, start + a.block);
  unsigned long long p = a.boff[RB_POS * st + b];
  const unsigned long long pend = a.boff[RB_POS * st + b + 1];
  unsigned long long prevpos = 0;
  bool seen_first = false;
  for (uint32_t base = start; base < end; base += 32) {
    const uint32_t i = base + lane;
    uint8_t f = '2';
    if (i < end) f = a.flag[i];
    const bool a1 = i < end && aligned1(f), p2 = i < end && a.paired && aligned2(f) && own_pos2(f);
    // the block's first aligned read 1 (an absolute entry)
    const unsigned m1 = __ballot_sync(FULL, a1);
    const bool is_first = a1 && !seen_first && lane == __ffs(m1) - 1;
    // offsets assuming no escape, then escapes from left to right
    uint32_t sz = (a1 ? (is_first ? 8u : 2u) : 0u) + (p2 ? 8u : 0u);
    uint32_t pre = sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(FULL, pre, d); if (lane >= d) pre += t; }
    uint32_t off = pre - sz;  // exclusive
    bool esc = false;
    int resolved = -1;
    for (;;) {
      bool cand = false;
      if (a1 && !is_first && lane > resolved && p + off + 2 <= pend) cand = ld_u16(a.pos + p + off) == 65535;
      const unsigned em = __ballot_sync(FULL, cand);
      if (!em) break;
      const int e = __ffs(em) - 1;  // every lane before e is final, so e's offset is: a true escape
      if (lane == e) esc = true;
      if (lane > e) off += 8;
      resolved = e;
    }
    const uint32_t my = (a1 ? (is_first ? 8u : (esc ? 10u : 2u)) : 0u) + (p2 ? 8u : 0u);
    const uint32_t total = __shfl_sync(FULL, off + my, 31);
    if (p + total > pend) { if (lane == 0) *a.err = 3; return; }
    // positions: absolute entries restart the running sum
    bool abs_ = false;
    unsigned long long v = 0;
    if (a1) {
      if (is_first) { abs_ = true; v = ld_u64(a.pos + p + off); }
      else if (esc) { abs_ = true; v = ld_u64(a.pos + p + off + 2); }
      else v = ld_u16(a.pos + p + off);
    }
    if (lane == 0 && !abs_) { v += prevpos; }  // carry in (lane 0 without an entry carries prevpos through)
    bool sa = abs_ || lane == 0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long tv = __shfl_up_sync(FULL, v, d);
      const bool ta = __shfl_up_sync(FULL, sa, d);
      if (lane >= d && !sa) { v += tv; sa = ta; }
    }
    if (a1) a.pos1[i] = v;
    if (p2) a.pos2[i] = ld_u64(a.pos + p + off + (my - 8));
    prevpos = __shfl_sync(FULL, v, 31);
    seen_first = seen_first || m1 != 0;
    p += total;
  }
  if (p != pend && lane == 0) *a.err = 4;
}

// order-preserving mode: every entry is an absolute u64 (decompress.cpp:236-237)
__global__ void k_dec_pos_preserve(DecArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.units) return;
  const uint8_t f = a.flag[i];
  unsigned long long p = a.off[i].pos;
  if (aligned1(f)) { a.pos1[i] = ld_u64(a.pos + p); p += 8; }
  if (a.paired && aligned2(f) && own_pos2(f)) a.pos2[i] = ld_u64(a.pos + p);
}

// file coding of read_seq.bin: A0 C1 G2 T3, 4 bases per byte, LSB first (encoder.cpp:126-141)
__device__ __forceinline__ uint8_t seq_base(const uint8_t *sp, unsigned long long x) { return "ACGT"[(sp[x >> 2] >> (2 * (x & 3))) & 3]; }
__device__ __forceinline__ uint8_t dec_noise(uint8_t ref, uint8_t sym) {  // decompress.cpp:664-685
  const int s = sym - '0';
  switch (ref) {
    case 'A': return "CGTN"[s];
    case 'C': return "AGTN"[s];
    case 'G': return "TACN"[s];
    case 'T': return "GCAN"[s];
    default: return "AGCT"[s];
  }
}
__device__ __forceinline__ uint8_t comp(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

__global__ void k_dec_reads(DecArgs a) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.paired ? 2 * a.units : a.units;
  if (t >= n) return;
  const int mate = t >= a.units ? 1 : 0;
  const uint32_t i = mate ? t - a.units : t;
  const uint8_t f = a.flag[i];
  const DecOff o = a.off[i];
  const bool a1 = aligned1(f), a2 = a.paired && aligned2(f);
  const unsigned long long ob = a.out_off[t];
  const int len = (int)(a.out_off[t + 1] - ob);
  uint8_t *dst = a.out + ob;
  if (mate == 0 ? !a1 : !a2) {  // :275-278, :313-317: raw text, read 1's before read 2's
    unsigned long long u = o.unal;
    if (mate == 1 && !a1) u += a.out_off[i + 1] - a.out_off[i];
    for (int k = 0; k < len; k++) dst[k] = a.unal[u + k];
    return;
  }
  unsigned long long pos;
  uint8_t rc;
  uint32_t line = o.lines;
  if (mate == 0) { pos = a.pos1[i]; rc = a.rc[o.rc]; }
  else {
    line += a1 ? 1u : 0u;
    if (own_pos2(f)) { pos = a.pos2[i]; rc = a.rc[o.rc + (a1 ? 1u : 0u)]; }
    else {  // :295-305: relative to read 1
      const int16_t d = (int16_t)ld_u16(a.pos_pair + 2ull * o.pair);
      pos = a.pos1[i] + (long long)d;
      const uint8_t rc1 = a.rc[o.rc];
      rc = a.rc_pair[o.pair] == '0' ? (rc1 == 'd' ? 'r' : 'd') : (rc1 == 'd' ? 'd' : 'r');
    }
  }
  if (pos > a.seq_len || (unsigned long long)len > a.seq_len - pos) { *a.err = 5; return; }  // no wrap for pos near 2^64
  const bool rev = rc != 'd';
  // consensus window, written in the read's final orientation; noise positions refer to the forward window
  for (int k = 0; k < len; k++) {
    const uint8_t c = seq_base(a.seq_packed, pos + k);
    if (!rev) dst[k] = c; else dst[len - 1 - k] = comp(c);
  }
  const uint32_t nb0 = line ? a.nl_pos[line - 1] + 1 : 0, nb1 = a.nl_pos[line];
  int np = 0;
  for (uint32_t k = nb0; k < nb1; k++) {  // :258-266
    np += ld_u16(a.noisepos + 2ull * (k - line));
    if (np >= len) { *a.err = 6; return; }
    const uint8_t sym = a.noise[k];
    if (sym < '0' || sym > '3') { *a.err = 7; return; }
    const uint8_t c = dec_noise(seq_base(a.seq_packed, pos + np), sym);
    if (!rev) dst[np] = c; else dst[len - 1 - np] = comp(c);
  }
}

}  // namespace

void run_decode_blocks(Ctx &c, const ReblockDev &b, const uint64_t *sizes, const uint8_t *d_seq_packed, uint64_t seq_len,
                       uint64_t num_reads, bool paired, bool preserve, uint32_t block, DecodeDev &out) {
  cudaStream_t st = c.stream;
  out = DecodeDev{};
  if (!block) throw LimitError("decode: num_reads_per_block is 0");
  if (paired && (num_reads & 1)) throw LimitError("decode: odd number of reads in paired-end mode");
  const uint32_t n = (uint32_t)num_reads, units = paired ? n / 2 : n;
  const uint32_t nb = (uint32_t)(((uint64_t)units + block - 1) / block);
  if (nb != b.num_blocks) throw LimitError("decode: block count does not match cp.num_reads / num_reads_per_block");
  if (sizes[RB_FLAG] != units || sizes[RB_LENGTHS] != 2ull * n) throw LimitError("decode: flag / length streams do not match cp.num_reads");
  if (sizes[RB_NOISE] >= 0x7FFFFFFFull) throw LimitError("decode: noise stream of >= 2 GiB");
  size_t cub_bytes = 1 << 20;
  void *cub_tmp = c.pool.device("dc.cubtmp", cub_bytes);
  auto cub_need = [&](size_t need) { if (need > cub_bytes) { cub_bytes = need; cub_tmp = c.pool.device("dc.cubtmp", cub_bytes); } };

  DecArgs a{};
  a.flag = b.data[RB_FLAG]; a.pos = b.data[RB_POS]; a.noise = b.data[RB_NOISE]; a.noisepos = b.data[RB_NOISEPOS]; a.rc = b.data[RB_RC];
  a.unal = b.data[RB_UNALIGNED]; a.len = b.data[RB_LENGTHS]; a.pos_pair = b.data[RB_POS_PAIR]; a.rc_pair = b.data[RB_RC_PAIR];
  a.boff = b.block_off; a.nb = nb; a.units = units; a.block = block; a.paired = paired; a.preserve = preserve;
  a.seq_packed = d_seq_packed; a.seq_len = seq_len;
  for (int s = 0; s < RB_NSTREAMS; s++) a.sizes[s] = sizes[s];
  int *d_err = c.pool.dev<int>("dc.err", 4);
  SB_CUDA(cudaMemsetAsync(d_err, 0, 4 * sizeof(int), st));
  a.err = d_err;
  // output offsets: file 1's reads, then file 2's
  unsigned long long *len64 = c.pool.dev<unsigned long long>("dc.len64", (size_t)n + 1);
  unsigned long long *out_off = c.pool.dev<unsigned long long>("dc.out_off", (size_t)n + 1);
  k_dec_lengths<<<grid_for((uint64_t)n + 1, 256), 256, 0, st>>>(a, len64);
  {
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, len64, out_off, (int)n + 1, st); cub_need(need);
    need = cub_bytes; cub::DeviceScan::ExclusiveSum(cub_tmp, need, len64, out_off, (int)n + 1, st);
  }
  c.launches += 3;
  a.out_off = out_off;
  // newline positions
  uint32_t *nl_pos = c.pool.dev<uint32_t>("dc.nl_pos", sizes[RB_NOISE] + 1);
  uint32_t *d_cnt = c.pool.dev<uint32_t>("dc.cnt", 4);
  uint32_t num_lines = 0;
  if (sizes[RB_NOISE]) {
    cub::CountingInputIterator<uint32_t> it(0);
    IsNewline pred{a.noise};
    size_t need = 0;
    cub::DeviceSelect::If(nullptr, need, it, nl_pos, d_cnt, (int)sizes[RB_NOISE], pred, st); cub_need(need);
    need = cub_bytes; cub::DeviceSelect::If(cub_tmp, need, it, nl_pos, d_cnt, (int)sizes[RB_NOISE], pred, st);
    c.launches += 2;
    SB_CUDA(cudaMemcpyAsync(&num_lines, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (2ull * (sizes[RB_NOISE] - num_lines) != sizes[RB_NOISEPOS]) throw LimitError("decode: noise and noisepos streams do not match");
  } else if (sizes[RB_NOISEPOS]) throw LimitError("decode: noise and noisepos streams do not match");
  a.nl_pos = nl_pos; a.num_lines = num_lines;
  // per-unit consumption -> offsets
  DecOff *off = c.pool.dev<DecOff>("dc.off", (size_t)units + 1);
  a.off = off;
  k_dec_units<<<grid_for((uint64_t)units + 1, 256), 256, 0, st>>>(a, out_off);
  {
    size_t need = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, need, off, off, DecAdd(), DecOff{}, (int)units + 1, st); cub_need(need);
    need = cub_bytes; cub::DeviceScan::ExclusiveScan(cub_tmp, need, off, off, DecAdd(), DecOff{}, (int)units + 1, st);
  }
  k_dec_check_blocks<<<grid_for((uint64_t)nb + 1, 128), 128, 0, st>>>(a);
  c.launches += 4;
  a.pos1 = c.pool.dev<unsigned long long>("dc.pos1", (size_t)units + 1);
  a.pos2 = c.pool.dev<unsigned long long>("dc.pos2", (size_t)units + 1);
  if (units) {
    if (preserve) k_dec_pos_preserve<<<grid_for(units, 256), 256, 0, st>>>(a);
    else k_dec_pos<<<grid_for(32ull * nb, 128), 128, 0, st>>>(a);
    c.launches++;
  }
  unsigned long long total = 0;
  int h_err[2] = {0, 0};
  SB_CUDA(cudaMemcpyAsync(&total, out_off + n, sizeof(total), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h_err[0]) throw LimitError("decode: inconsistent block streams (check " + std::to_string(h_err[0]) + ")");
  out.bases = c.pool.dev<uint8_t>("dc.out", total + 1);
  a.out = out.bases;
  if (n) { k_dec_reads<<<grid_for(n, 128), 128, 0, st>>>(a); c.launches++; }
  SB_CUDA(cudaMemcpyAsync(h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h_err[0]) throw LimitError("decode: inconsistent block streams (check " + std::to_string(h_err[0]) + ")");
  out.offsets = out_off; out.total = total; out.num_reads = n;
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
