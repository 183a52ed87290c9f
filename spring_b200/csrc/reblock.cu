// reblock.cu -- the two host stages that follow the hot path in `spring -c`, on the GPU:
//
//   pe_encode                (reference src/pe_encode.cpp:24-84)        read_order.bin -> output slots
//   reorder_compress_streams (src/reorder_compress_streams.cpp:83-361)  encoder streams -> per-block streams
//                            (its BSC calls, :363-428, stay on the host)
//
// The reference loads every stream into RAM arrays indexed by output slot and then writes the blocks
// of 256 000 reads (pairs) one after the other, one thread per block.  Blocks are contiguous ranges
// of output slots, so the concatenation of all blocks of a stream is ONE stream in slot order; here
//   1. inv[slot] = stream index of the read that lands in the slot          (scatter)
//   2. where each aligned read's noise starts = position of the newline before it (stream compaction)
//   3. previous aligned read-1 of the same block, for the delta-coded positions  (max-scan)
//   4. bytes every unit (read, or pair) adds to each variable-rate stream       (one thread per unit)
//   5. one exclusive scan of those size tuples = the unit's offset in every stream; the scan values at
//      multiples of num_reads_per_block are the block boundaries
//   6. every unit writes its bytes                                             (one thread per unit)
// Bit-exact against oracle/reblock_oracle.c, which is pinned against the reference's own files.
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {
namespace {

static inline uint32_t grid_for(uint64_t n, int block) { return (uint32_t)((n + block - 1) / block); }

// ---- pe_encode ----------------------------------------------------------------------------------------
// inverse[] starts as all ones; an index out of range or taken twice sets *err instead of being used
// (order comes from a file or a caller when the streams are not the library's own)
__global__ void k_invert(const uint32_t *__restrict__ order, uint32_t n, uint32_t *inverse, int *err) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = order[i];
  if (o >= n) { *err = 1; return; }
  if (atomicExch(inverse + o, i) != 0xFFFFFFFFu) *err = 1;
}
__global__ void k_is_file1(const uint32_t *__restrict__ order, uint32_t n, uint32_t half, uint32_t *f1) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f1[i] = order[i] < half ? 1u : 0u;
}
// pe_encode.cpp:53-69: file-1 reads are numbered in stream order; a file-2 read takes its mate's number + half
__global__ void k_pe_slots(const uint32_t *__restrict__ order, const uint32_t *__restrict__ inverse,
                           const uint32_t *__restrict__ rank1, uint32_t n, uint32_t half, uint32_t *slot) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = order[i];
  slot[i] = o < half ? rank1[i] : rank1[inverse[o - half]] + half;
}

// ---- re-blocking -----------------------------------------------------------------------------------------
struct IsNewline {
  const uint8_t *noise;
  __device__ bool operator()(uint32_t p) const { return noise[p] == '\n'; }
};
struct MaxOp {
  __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};
// bytes (entries) one unit adds to the variable-rate streams; summed by the scan
struct UnitOff {
  unsigned long long pos, noise, unal;
  uint32_t rc, pair, nal, pad;  // nal: aligned reads (= newlines in the noise stream)
};
struct UnitAdd {
  __device__ UnitOff operator()(const UnitOff &a, const UnitOff &b) const {
    UnitOff r;
    r.pos = a.pos + b.pos; r.noise = a.noise + b.noise; r.unal = a.unal + b.unal;
    r.rc = a.rc + b.rc; r.pair = a.pair + b.pair; r.nal = a.nal + b.nal; r.pad = 0;
    return r;
  }
};

struct ReblockArgs {
  // encoder streams (device)
  const uint64_t *pos; const uint8_t *noise; const uint16_t *noisepos; const uint8_t *rev;
  const uint16_t *lengths; const uint8_t *unaligned;
  uint32_t n, ma;                  // reads, aligned reads
  const uint32_t *inv;             // slot -> stream index (nullptr: identity, SE without order)
  const uint32_t *nl_pos;          // [ma] position of the newline that ends aligned read i's noise
  const unsigned long long *urec;  // [n - ma] byte offset of unaligned read u's record
  const uint32_t *last1;           // [units] 1 + last unit <= i whose read 1 is aligned, 0 if none (inclusive max-scan)
  uint32_t units, half, block;
  int paired, preserve;
  UnitOff *off;                    // [units + 1] sizes, scanned in place to offsets
  // outputs (device)
  uint8_t *o_flag; uint8_t *o_pos; uint8_t *o_noise; uint8_t *o_noisepos; uint8_t *o_rc; uint8_t *o_unal;
  uint16_t *o_len; uint8_t *o_pos_pair; uint8_t *o_rc_pair;
};

__device__ __forceinline__ uint32_t stream_of(const ReblockArgs &a, uint32_t slot) { return a.inv ? a.inv[slot] : slot; }
__device__ __forceinline__ uint32_t noise_begin(const ReblockArgs &a, uint32_t s) { return s ? a.nl_pos[s - 1] + 1 : 0; }

// bytes the position of read 1 of unit i takes (reorder_compress_streams.cpp:254-271 / :301-318), and the
// position itself: absolute u64 in order-preserving mode and for the first unit of a block, else a u16
// delta to the previous aligned read 1 OF THE BLOCK (0 if there is none), 65535 + u64 when it does not fit
__device__ __forceinline__ int pos1_bytes(const ReblockArgs &a, uint32_t i, uint32_t s1, uint64_t &p, uint64_t &diff) {
  p = a.pos[s1];
  diff = 0;
  if (a.preserve) return 8;
  const uint32_t bstart = (i / a.block) * a.block;
  if (i == bstart) return 8;
  uint64_t prevpos = 0;
  const uint32_t l = a.last1[i - 1];  // 1 + index
  if (l > bstart) prevpos = a.pos[stream_of(a, l - 1)];
  diff = p - prevpos;
  return diff < 65535 ? 2 : 10;
}
__device__ __forceinline__ void put_u16(uint8_t *o, uint16_t v) { o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); }
__device__ __forceinline__ void put_u64(uint8_t *o, uint64_t v) {
#pragma unroll
  for (int k = 0; k < 8; k++) o[k] = (uint8_t)(v >> (8 * k));
}

__global__ void k_aligned1(ReblockArgs a, uint32_t *key) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.units) key[i] = stream_of(a, i) < a.ma ? i + 1 : 0u;
}

// pair flag, reorder_compress_streams.cpp:283-299 (0: both aligned and |distance| < 32767, 1: both aligned,
// 2: none, 3: only read 1, 4: only read 2); single end: 0 aligned / 2 unaligned (:252-280)
__device__ __forceinline__ int unit_flag(const ReblockArgs &a, uint32_t s1, uint32_t s2) {
  const bool a1 = s1 < a.ma;
  if (!a.paired) return a1 ? 0 : 2;
  const bool a2 = s2 < a.ma;
  if (a1 && a2) {
    const long long d = (long long)a.pos[s2] - (long long)a.pos[s1];
    return (d < 0 ? -d : d) < 32767 ? 0 : 1;
  }
  if (!a1 && !a2) return 2;
  return a1 ? 3 : 4;
}

__global__ void k_unit_sizes(ReblockArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.units) return;
  UnitOff u{};
  if (i < a.units) {
    const uint32_t s1 = stream_of(a, i), s2 = a.paired ? stream_of(a, a.half + i) : 0;
    const int flag = unit_flag(a, s1, s2);
    a.o_flag[i] = (uint8_t)('0' + flag);
    if (!a.paired) a.o_len[i] = a.lengths[s1];
    else { a.o_len[2 * (size_t)i] = a.lengths[s1]; a.o_len[2 * (size_t)i + 1] = a.lengths[s2]; }
    if (flag == 0 || flag == 1 || flag == 3) {  // read 1 aligned
      uint64_t p, d;
      u.pos += pos1_bytes(a, i, s1, p, d);
      u.noise += a.nl_pos[s1] - noise_begin(a, s1) + 1;
      u.nal += 1; u.rc += 1;
    } else {
      u.unal += a.lengths[s1];
    }
    if (a.paired) {
      if (flag == 0) u.pair = 1;
      if (flag == 0 || flag == 1 || flag == 4) {  // read 2 aligned
        u.noise += a.nl_pos[s2] - noise_begin(a, s2) + 1;
        u.nal += 1;
        if (flag != 0) { u.pos += 8; u.rc += 1; }
      } else {
        u.unal += a.lengths[s2];
      }
    }
  }
  a.off[i] = u;  // entry `units` stays zero: after the exclusive scan it holds the totals
}

__device__ void copy_noise(const ReblockArgs &a, uint32_t s, unsigned long long &onoise, uint32_t &onal) {
  const uint32_t b = noise_begin(a, s), e = a.nl_pos[s];
  const unsigned long long np = onoise - onal;  // noise symbols written so far = index in noisepos
  for (uint32_t k = b; k < e; k++) {
    a.o_noise[onoise + (k - b)] = a.noise[k];
    put_u16(a.o_noisepos + 2 * (np + (k - b)), a.noisepos[(k - s) ]);  // k - s: newlines before k removed
  }
  a.o_noise[onoise + (e - b)] = '\n';
  onoise += e - b + 1;
  onal += 1;
}
__device__ void copy_unaligned(const ReblockArgs &a, uint32_t s, unsigned long long &ounal) {
  const uint8_t *rec = a.unaligned + a.urec[s - a.ma] + 2;  // {u16 len; 4 bits/base A0 G1 C2 T3 N4} (util.cpp:322-374)
  const int len = a.lengths[s];
  for (int j = 0; j < len; j++) a.o_unal[ounal + j] = (uint8_t)"AGCTN"[(rec[j >> 1] >> (4 * (j & 1))) & 15];
  ounal += len;
}

__global__ void k_unit_write(ReblockArgs a) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.units) return;
  UnitOff o = a.off[i];
  const uint32_t s1 = stream_of(a, i), s2 = a.paired ? stream_of(a, a.half + i) : 0;
  const int flag = a.o_flag[i] - '0';
  if (flag == 0 && a.paired) {  // :293-299
    put_u16(a.o_pos_pair + 2 * (size_t)o.pair, (uint16_t)(int16_t)((long long)a.pos[s2] - (long long)a.pos[s1]));
    a.o_rc_pair[o.pair] = a.rev[s1] != a.rev[s2] ? '0' : '1';
  }
  if (flag == 0 || flag == 1 || flag == 3) {
    uint64_t p, d;
    const int nb = pos1_bytes(a, i, s1, p, d);
    if (nb == 8) put_u64(a.o_pos + o.pos, p);
    else {
      put_u16(a.o_pos + o.pos, nb == 2 ? (uint16_t)d : (uint16_t)65535);
      if (nb == 10) put_u64(a.o_pos + o.pos + 2, p);
    }
    o.pos += nb;
    copy_noise(a, s1, o.noise, o.nal);
    a.o_rc[o.rc++] = a.rev[s1];
  } else {
    copy_unaligned(a, s1, o.unal);
  }
  if (a.paired) {
    if (flag == 0 || flag == 1 || flag == 4) {
      copy_noise(a, s2, o.noise, o.nal);
      if (flag != 0) { put_u64(a.o_pos + o.pos, a.pos[s2]); a.o_rc[o.rc++] = a.rev[s2]; }
    } else {
      copy_unaligned(a, s2, o.unal);
    }
  }
}

__global__ void k_urec_sizes(const uint16_t *__restrict__ lengths, uint32_t ma, uint32_t nu, unsigned long long *sz) {
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < nu) sz[u] = 2ull + ((unsigned long long)lengths[ma + u] + 1) / 2;
  if (u == nu) sz[u] = 0;
}

// block boundaries: byte offsets of block b in every stream = the scan value at unit b * block
__global__ void k_block_offsets(const UnitOff *__restrict__ off, uint32_t units, uint32_t block, uint32_t nb, int paired,
                                unsigned long long *out /* [9][nb + 1] */) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > nb) return;
  const unsigned long long i = (unsigned long long)b * block < units ? (unsigned long long)b * block : units;
  const UnitOff o = off[i];
  const size_t st = (size_t)nb + 1;
  out[0 * st + b] = i;                               // flag: one char per unit
  out[1 * st + b] = o.pos;
  out[2 * st + b] = o.noise;
  out[3 * st + b] = 2ull * (o.noise - o.nal);        // noisepos: u16 per noise symbol
  out[4 * st + b] = o.rc;
  out[5 * st + b] = o.unal;
  out[6 * st + b] = (paired ? 4ull : 2ull) * i;      // lengths: u16 per read
  out[7 * st + b] = 2ull * o.pair;
  out[8 * st + b] = o.pair;
}

}  // namespace

void run_pe_encode(Ctx &c, const uint32_t *order, uint32_t n, uint32_t *slot) {
  cudaStream_t st = c.stream;
  if (!n) return;
  const uint32_t half = n / 2;
  uint32_t *inverse = c.pool.dev<uint32_t>("rb.inverse", n), *f1 = c.pool.dev<uint32_t>("rb.f1", n), *rank1 = c.pool.dev<uint32_t>("rb.rank1", n);
  int *err = c.pool.dev<int>("rb.perm_err", 1);
  SB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
  SB_CUDA(cudaMemsetAsync(inverse, 0xFF, sizeof(uint32_t) * n, st));
  k_invert<<<grid_for(n, 256), 256, 0, st>>>(order, n, inverse, err);
  {
    int bad = 0;
    SB_CUDA(cudaMemcpyAsync(&bad, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (bad) throw LimitError("pe_encode: order is not a permutation of 0..num_reads-1");
  }
  k_is_file1<<<grid_for(n, 256), 256, 0, st>>>(order, n, half, f1);
  size_t need = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, f1, rank1, (int)n, st);
  void *tmp = c.pool.device("pe.cubtmp", need);  // its own buffer: run_reblock holds a pointer into "rb.cubtmp" across this call
  cub::DeviceScan::ExclusiveSum(tmp, need, f1, rank1, (int)n, st);
  k_pe_slots<<<grid_for(n, 256), 256, 0, st>>>(order, inverse, rank1, n, half, slot);
  c.launches += 5;
  SB_CUDA(cudaGetLastError());
}

void run_reblock(Ctx &c, const EncodeDev &e, bool paired, bool preserve, uint32_t block, ReblockDev &out) {
  cudaStream_t st = c.stream;
  out = ReblockDev{};
  const uint32_t n = (uint32_t)e.num_reads, ma = (uint32_t)e.num_aligned, nu = n - ma;
  if (!block) throw LimitError("reblock: num_reads_per_block is 0");
  if (paired && (n & 1)) throw LimitError("reblock: odd number of reads in paired-end mode");
  if (e.noise_bytes >= 0x7FFFFFFFull) throw LimitError("reblock: noise stream of >= 2 GiB per shard");
  const uint32_t half = n / 2, units = paired ? half : n;
  const uint32_t nb = (uint32_t)(((uint64_t)units + block - 1) / block);
  out.num_blocks = nb;
  size_t cub_bytes = 1 << 20;
  void *cub_tmp = c.pool.device("rb.cubtmp", cub_bytes);
  auto cub_need = [&](size_t need) { if (need > cub_bytes) { cub_bytes = need; cub_tmp = c.pool.device("rb.cubtmp", cub_bytes); } };

  ReblockArgs a{};
  a.pos = e.pos; a.noise = e.noise; a.noisepos = e.noisepos; a.rev = e.rev; a.lengths = e.lengths; a.unaligned = e.unaligned;
  a.n = n; a.ma = ma; a.units = units; a.half = half; a.block = block; a.paired = paired; a.preserve = preserve;

  // 1. output slot of every stream read (reorder_compress_streams.cpp:114-115,:139: the order file is only
  //    read for paired-end or order-preserving runs; pe_encode rewrote it for -r paired input, spring.cpp:193)
  uint32_t *slot = nullptr;
  out.order = nullptr;
  if ((paired || preserve) && n) {
    slot = c.pool.dev<uint32_t>("rb.slot", n);
    if (paired && !preserve) run_pe_encode(c, e.order, n, slot);
    else SB_CUDA(cudaMemcpyAsync(slot, e.order, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, st));
    uint32_t *inv = c.pool.dev<uint32_t>("rb.inv", n);
    int *err = c.pool.dev<int>("rb.perm_err", 1);
    SB_CUDA(cudaMemsetAsync(err, 0, sizeof(int), st));
    SB_CUDA(cudaMemsetAsync(inv, 0xFF, sizeof(uint32_t) * n, st));
    k_invert<<<grid_for(n, 256), 256, 0, st>>>(slot, n, inv, err);
    c.launches++;
    int bad = 0;
    SB_CUDA(cudaMemcpyAsync(&bad, err, sizeof(int), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (bad) throw LimitError("reblock: read_order.bin is not a permutation of 0..num_reads-1");
    a.inv = inv;
    out.order = slot;
  }
  // 2. newline positions of the noise stream
  uint32_t *nl_pos = c.pool.dev<uint32_t>("rb.nl_pos", (size_t)ma + 1);
  uint32_t *d_cnt = c.pool.dev<uint32_t>("rb.cnt", 4);
  if (ma) {
    cub::CountingInputIterator<uint32_t> it(0);
    IsNewline pred{e.noise};
    size_t need = 0;
    cub::DeviceSelect::If(nullptr, need, it, nl_pos, d_cnt, (int)e.noise_bytes, pred, st); cub_need(need);
    need = cub_bytes; cub::DeviceSelect::If(cub_tmp, need, it, nl_pos, d_cnt, (int)e.noise_bytes, pred, st);
    c.launches += 2;
    uint32_t got = 0;
    SB_CUDA(cudaMemcpyAsync(&got, d_cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (got != ma) throw LimitError("reblock: read_noise.txt does not hold one line per aligned read");
  }
  a.nl_pos = nl_pos;
  // unaligned records
  unsigned long long *usz = c.pool.dev<unsigned long long>("rb.usz", (size_t)nu + 1), *urec = c.pool.dev<unsigned long long>("rb.urec", (size_t)nu + 1);
  if (nu) {
    k_urec_sizes<<<grid_for((uint64_t)nu + 1, 256), 256, 0, st>>>(e.lengths, ma, nu, usz);
    size_t need = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, need, usz, urec, (int)nu + 1, st); cub_need(need);
    need = cub_bytes; cub::DeviceScan::ExclusiveSum(cub_tmp, need, usz, urec, (int)nu + 1, st);
    c.launches += 3;
    unsigned long long total = 0;
    SB_CUDA(cudaMemcpyAsync(&total, urec + nu, sizeof(total), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (total != e.unaligned_bytes) throw LimitError("reblock: read_unaligned.txt does not match read_lengths.bin");
  }
  a.urec = urec;
  // 3. previous aligned read 1
  uint32_t *key1 = c.pool.dev<uint32_t>("rb.key1", (size_t)units + 1), *last1 = c.pool.dev<uint32_t>("rb.last1", (size_t)units + 1);
  if (units) {
    k_aligned1<<<grid_for(units, 256), 256, 0, st>>>(a, key1);
    size_t need = 0;
    cub::DeviceScan::InclusiveScan(nullptr, need, key1, last1, MaxOp(), (int)units, st); cub_need(need);
    need = cub_bytes; cub::DeviceScan::InclusiveScan(cub_tmp, need, key1, last1, MaxOp(), (int)units, st);
    c.launches += 3;
  }
  a.last1 = last1;
  // 4./5. sizes -> offsets; the fixed-rate streams (flag, lengths) are written by the size kernel
  UnitOff *off = c.pool.dev<UnitOff>("rb.off", (size_t)units + 1);
  a.off = off;
  out.data[RB_FLAG] = c.pool.dev<uint8_t>("rb.o_flag", (size_t)units + 1);
  out.data[RB_LENGTHS] = reinterpret_cast<uint8_t *>(c.pool.dev<uint16_t>("rb.o_len", (size_t)n + 1));
  a.o_flag = out.data[RB_FLAG]; a.o_len = reinterpret_cast<uint16_t *>(out.data[RB_LENGTHS]);
  k_unit_sizes<<<grid_for((uint64_t)units + 1, 256), 256, 0, st>>>(a);
  {
    size_t need = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, need, off, off, UnitAdd(), UnitOff{}, (int)units + 1, st); cub_need(need);
    need = cub_bytes; cub::DeviceScan::ExclusiveScan(cub_tmp, need, off, off, UnitAdd(), UnitOff{}, (int)units + 1, st);
  }
  c.launches += 3;
  UnitOff tot{};
  SB_CUDA(cudaMemcpyAsync(&tot, off + units, sizeof(UnitOff), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (tot.noise != e.noise_bytes || tot.nal != ma || tot.unal != e.unaligned_len)
    throw LimitError("reblock: stream sizes are not conserved (inconsistent input streams)");
  out.size[RB_FLAG] = units; out.size[RB_POS] = tot.pos; out.size[RB_NOISE] = tot.noise; out.size[RB_NOISEPOS] = 2 * (tot.noise - tot.nal);
  out.size[RB_RC] = tot.rc; out.size[RB_UNALIGNED] = tot.unal; out.size[RB_LENGTHS] = 2ull * n;
  out.size[RB_POS_PAIR] = 2ull * tot.pair; out.size[RB_RC_PAIR] = tot.pair;
  out.data[RB_POS] = c.pool.dev<uint8_t>("rb.o_pos", out.size[RB_POS] + 1);
  out.data[RB_NOISE] = c.pool.dev<uint8_t>("rb.o_noise", out.size[RB_NOISE] + 1);
  out.data[RB_NOISEPOS] = c.pool.dev<uint8_t>("rb.o_noisepos", out.size[RB_NOISEPOS] + 2);
  out.data[RB_RC] = c.pool.dev<uint8_t>("rb.o_rc", out.size[RB_RC] + 1);
  out.data[RB_UNALIGNED] = c.pool.dev<uint8_t>("rb.o_unal", out.size[RB_UNALIGNED] + 1);
  out.data[RB_POS_PAIR] = c.pool.dev<uint8_t>("rb.o_pos_pair", out.size[RB_POS_PAIR] + 2);
  out.data[RB_RC_PAIR] = c.pool.dev<uint8_t>("rb.o_rc_pair", out.size[RB_RC_PAIR] + 1);
  a.o_pos = out.data[RB_POS]; a.o_noise = out.data[RB_NOISE]; a.o_noisepos = out.data[RB_NOISEPOS]; a.o_rc = out.data[RB_RC];
  a.o_unal = out.data[RB_UNALIGNED]; a.o_pos_pair = out.data[RB_POS_PAIR]; a.o_rc_pair = out.data[RB_RC_PAIR];
  // 6. write
  if (units) { k_unit_write<<<grid_for(units, 256), 256, 0, st>>>(a); c.launches++; }
  out.block_off = c.pool.dev<unsigned long long>("rb.block_off", (size_t)RB_NSTREAMS * (nb + 1));
  k_block_offsets<<<grid_for((uint64_t)nb + 1, 128), 128, 0, st>>>(off, units, block, nb, paired ? 1 : 0, out.block_off);
  c.launches++;
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
