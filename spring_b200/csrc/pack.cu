// pack.cu -- the read path of preprocess on the GPU: split reads with N from clean
// reads and pack both, producing in HBM exactly the arrays the hot path starts from.
//
// Reference: src/preprocess.cpp:196-207 (length limits, "mark reads with N"), :293-304 (clean reads to
// input_clean_<j>.dna in input order, N reads to input_N.dna with their original index in
// read_order_N.bin, file-2 indices offset by the file-1 read count, :364-378), and the record packers
// src/util.cpp:269-294 (2 bits/base A0 G1 C2 T3) and :322-348 (4 bits/base, N = 4).  The reference's
// tables are only defined for A, C, G, T (and N): any other character is refused here
// (SPRING_B200_EINVAL) instead of packing an undefined value.
//
//   1. k_scan_reads    : warp per read: length, has-N flag, bad-character flag
//   2. scans           : clean rank / N rank of every read, byte offset of every N record, max length
//   3. k_pack_clean    : thread per 64-bit word of a clean read's bitset row (32 bases), the layout
//                        readDnaFile builds in RAM (reorder.h:222-244)
//   4. k_pack_n        : thread per N read: {u16 len; 4-bit codes} record + read_order_N entry
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {
namespace {

static inline uint32_t grid_for(uint64_t n, int block) { return (uint32_t)((n + block - 1) / block); }

__global__ void k_scan_reads(const uint8_t *__restrict__ bases, const unsigned long long *__restrict__ offsets, uint32_t n,
                             uint32_t *len, uint32_t *is_clean, uint32_t *is_n, unsigned long long *nrec_bytes, int *bad) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i > n) return;
  if (i == n) {  // closing entries of the exclusive scans
    if (lane == 0) { is_clean[i] = 0; is_n[i] = 0; nrec_bytes[i] = 0; len[i] = 0; }
    return;
  }
  const unsigned long long b = offsets[i], e = offsets[i + 1];
  const uint32_t l = (uint32_t)(e - b);
  bool hn = false, hb = false;
  for (unsigned long long p = b + lane; p < e; p += 32) {
    const uint8_t ch = bases[p];
    hn |= ch == 'N';
    hb |= !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'N');
  }
  hn = __any_sync(0xFFFFFFFFu, hn);
  hb = __any_sync(0xFFFFFFFFu, hb);
  if (lane == 0) {
    len[i] = l;
    is_clean[i] = hn ? 0u : 1u;
    is_n[i] = hn ? 1u : 0u;
    nrec_bytes[i] = hn ? 2ull + (l + 1) / 2 : 0ull;
    if (hb) *bad = 1;
    if (e < b) *bad = 2;
  }
}

// A0 G1 C2 T3 from ASCII: (c >> 1) & 3 is A0 C1 T2 G3 -> table {0, 2, 3, 1}
__device__ __forceinline__ uint32_t code2(uint8_t c) { return (0x78u >> (2 * ((c >> 1) & 3))) & 3u; }

__global__ void k_pack_clean(const uint8_t *__restrict__ bases, const unsigned long long *__restrict__ offsets,
                             const uint32_t *__restrict__ is_clean, const uint32_t *__restrict__ clean_rank, uint32_t n, int W,
                             uint64_t *reads, uint16_t *lens) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t i = (uint32_t)(t / W);
  const int w = (int)(t - (uint64_t)i * W);
  if (i >= n || !is_clean[i]) return;
  const unsigned long long b = offsets[i];
  const int l = (int)(offsets[i + 1] - b);
  const uint32_t r = clean_rank[i];
  uint64_t v = 0;
  const int j0 = 32 * w, j1 = min(l, j0 + 32);
  for (int j = j0; j < j1; j++) v |= (uint64_t)code2(bases[b + j]) << (2 * (j - j0));
  reads[(size_t)r * W + w] = v;
  if (w == 0) lens[r] = (uint16_t)l;
}

__global__ void k_pack_n(const uint8_t *__restrict__ bases, const unsigned long long *__restrict__ offsets,
                         const uint32_t *__restrict__ is_n, const uint32_t *__restrict__ n_rank,
                         const unsigned long long *__restrict__ nrec_off, uint32_t n, uint8_t *records, uint32_t *order_n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !is_n[i]) return;
  const unsigned long long b = offsets[i];
  const int l = (int)(offsets[i + 1] - b);
  uint8_t *o = records + nrec_off[i];
  o[0] = (uint8_t)(l & 0xFF); o[1] = (uint8_t)(l >> 8);
  for (int k = 0; k < (l + 1) / 2; k++) {
    uint8_t v = 0;
    for (int h = 0; h < 2; h++) {
      const int j = 2 * k + h;
      if (j < l) { const uint8_t ch = bases[b + j]; v |= (uint8_t)((ch == 'N' ? 4u : code2(ch)) << (4 * h)); }
    }
    o[2 + k] = v;
  }
  order_n[n_rank[i]] = i;  // preprocess.cpp:300-301: original index, file 2 after file 1
}


// input_N.dna records {u16 len; ceil(len / 2) bytes, 4 bits per base A0 G1 C2 T3 N4} (util.cpp:322-374) -> the 2-bit
// rows (N stored as 00) + N bit-plane the encoder's pool keeps: thread per (read, word of 32 bases).  The record
// offsets come from the host (one hop per record over the length fields; the per-base work is here).
__global__ void k_unpack_n(const uint8_t *__restrict__ rec, const unsigned long long *__restrict__ off, uint32_t nn, int W,
                           uint64_t *__restrict__ codes, uint64_t *__restrict__ nflag, uint16_t *__restrict__ lens) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (uint64_t)nn * W) return;
  const uint32_t i = (uint32_t)(t / W);
  const int w = (int)(t - (uint64_t)i * W);
  const uint8_t *r = rec + off[i];
  const int len = (int)r[0] | ((int)r[1] << 8);
  if (w == 0) lens[i] = (uint16_t)len;
  uint64_t c = 0, f = 0;
  const int j0 = 32 * w, j1 = min(len, j0 + 32);
  for (int j = j0; j < j1; j += 2) {  // j0 is even: one byte = bases j, j + 1
    const uint32_t b = r[2 + (j >> 1)];
    const uint32_t v0 = b & 15u, v1 = b >> 4;
    const int sh = 2 * (j - j0);
    if (v0 >= 4) f |= 1ull << sh; else c |= (uint64_t)v0 << sh;
    if (j + 1 < j1) { if (v1 >= 4) f |= 1ull << (sh + 2); else c |= (uint64_t)v1 << (sh + 2); }
  }
  codes[t] = c; nflag[t] = f;
}
}  // namespace

void run_pack_reads(Ctx &c, const uint8_t *d_bases, const unsigned long long *d_offsets, uint32_t n, uint32_t n_file1, PackDev &out) {
  cudaStream_t st = c.stream;
  out = PackDev{};
  out.num_reads = n;
  const uint32_t nn = n + 1;
  uint32_t *len = c.pool.dev<uint32_t>("pk.len", nn), *is_clean = c.pool.dev<uint32_t>("pk.is_clean", nn), *is_n = c.pool.dev<uint32_t>("pk.is_n", nn);
  uint32_t *clean_rank = c.pool.dev<uint32_t>("pk.clean_rank", nn), *n_rank = c.pool.dev<uint32_t>("pk.n_rank", nn);
  unsigned long long *nrec_bytes = c.pool.dev<unsigned long long>("pk.nrec_bytes", nn), *nrec_off = c.pool.dev<unsigned long long>("pk.nrec_off", nn);
  uint32_t *d_misc = c.pool.dev<uint32_t>("pk.misc", 4);  // [0] bad flag, [1] max length
  SB_CUDA(cudaMemsetAsync(d_misc, 0, 4 * sizeof(uint32_t), st));
  k_scan_reads<<<grid_for(32ull * nn, 256), 256, 0, st>>>(d_bases, d_offsets, n, len, is_clean, is_n, nrec_bytes, reinterpret_cast<int *>(d_misc));
  size_t need = 0, tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, is_clean, clean_rank, (int)nn, st); tmp_bytes = need;
  cub::DeviceScan::ExclusiveSum(nullptr, need, nrec_bytes, nrec_off, (int)nn, st); if (need > tmp_bytes) tmp_bytes = need;
  cub::DeviceReduce::Max(nullptr, need, len, d_misc + 1, (int)nn, st); if (need > tmp_bytes) tmp_bytes = need;
  void *tmp = c.pool.device("pk.cubtmp", tmp_bytes);
  need = tmp_bytes; cub::DeviceScan::ExclusiveSum(tmp, need, is_clean, clean_rank, (int)nn, st);
  need = tmp_bytes; cub::DeviceScan::ExclusiveSum(tmp, need, is_n, n_rank, (int)nn, st);
  need = tmp_bytes; cub::DeviceScan::ExclusiveSum(tmp, need, nrec_bytes, nrec_off, (int)nn, st);
  need = tmp_bytes; cub::DeviceReduce::Max(tmp, need, len, d_misc + 1, (int)nn, st);
  c.launches += 9;
  struct { uint32_t misc[2]; uint32_t num_clean, clean_file1; unsigned long long nbytes; } h{};
  SB_CUDA(cudaMemcpyAsync(h.misc, d_misc, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&h.num_clean, clean_rank + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&h.clean_file1, clean_rank + n_file1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(&h.nbytes, nrec_off + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  if (h.misc[0] == 2) throw LimitError("pack_reads: offsets are not ascending");
  if (h.misc[0]) throw LimitError("pack_reads: a read holds a character other than A, C, G, T, N");
  if (h.misc[1] > (uint32_t)kMaxReadLen)  // preprocess.cpp:199-206
    throw LimitError("Too long read length (please try --long/-l flag).");
  out.max_readlen = h.misc[1];
  out.num_clean = h.num_clean; out.num_clean_file1 = h.clean_file1; out.num_n = n - h.num_clean; out.n_record_bytes = h.nbytes;
  const int W = words_for(out.max_readlen ? (int)out.max_readlen : 1);
  out.W = W;
  out.reads = c.pool.dev<uint64_t>("pk.reads", (size_t)(out.num_clean ? out.num_clean : 1) * W);
  out.lengths = c.pool.dev<uint16_t>("pk.lengths", out.num_clean ? out.num_clean : 1);
  out.n_records = c.pool.dev<uint8_t>("pk.n_records", out.n_record_bytes + 1);
  out.order_n = c.pool.dev<uint32_t>("pk.order_n", out.num_n + 1);
  if (n) {
    k_pack_clean<<<grid_for((uint64_t)n * W, 256), 256, 0, st>>>(d_bases, d_offsets, is_clean, clean_rank, n, W, out.reads, out.lengths);
    k_pack_n<<<grid_for(n, 128), 128, 0, st>>>(d_bases, d_offsets, is_n, n_rank, nrec_off, n, out.n_records, out.order_n);
    c.launches += 2;
  }
  SB_CUDA(cudaGetLastError());
}

void run_unpack_n(Ctx &c, const uint8_t *d_records, const unsigned long long *d_offsets, uint32_t nn, int W, uint64_t *codes,
                  uint64_t *nflag, uint16_t *lens) {
  if (!nn) return;
  k_unpack_n<<<grid_for((uint64_t)nn * W, 256), 256, 0, c.stream>>>(d_records, d_offsets, nn, W, codes, nflag, lens);
  c.launches++;
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
