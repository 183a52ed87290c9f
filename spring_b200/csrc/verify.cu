// verify.cu -- full-size round trip without leaving HBM: the encoder streams of the last
// spring_b200_reorder_encode* call -> re-blocking (reblock.cu, what reorder_compress_streams writes) ->
// block decode (decode.cu, what decompress_short reads) -> every decoded read compared base by base with
// the input read that read_order.bin says it is.  This is the reference's own -r check
// (util/test_script.sh:78-82: decompress, sort, compare) made exact -- the order stream names the original
// of every decoded read, so no sort is needed -- and it runs at the bench's full sizes, where a host-side
// comparison of 15 GB of bases would not.
#include "kernels.cuh"

namespace sb {
namespace {

constexpr unsigned FULL = 0xFFFFFFFFu;

struct VerifyArgs {
  // decoded reads (decode.cu): ASCII, slot s at [off[s], off[s + 1])
  const uint8_t *bases; const unsigned long long *off;
  const uint32_t *slot;    // stream index -> output slot (nullptr: identity)
  const uint32_t *order;   // stream index -> original index (read_order.bin as the encoder wrote it)
  uint32_t n;              // reads
  // the input: clean reads (2 bits/base rows), reads with N (codes + N bit-plane), their original indices
  const uint64_t *reads; const uint16_t *lens; uint32_t num_clean; int W;
  NReads nr;
  uint32_t *seen;          // bitmap over original indices
  unsigned long long *bad; // [4]: base mismatches, length mismatches, duplicate / out-of-range originals, reads checked
};

// One warp per stream read.
__global__ void k_verify(VerifyArgs a) {
  const uint32_t i = (uint32_t)(((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= a.n) return;
  const uint32_t s = a.slot ? a.slot[i] : i;
  const uint32_t o = a.order[i];
  if (o >= a.n || s >= a.n) { if (lane == 0) atomicAdd(a.bad + 2, 1ull); return; }
  if (lane == 0) {
    const uint32_t old = atomicOr(a.seen + (o >> 5), 1u << (o & 31));
    if ((old >> (o & 31)) & 1u) atomicAdd(a.bad + 2, 1ull);
  }
  // original index -> N read j or clean read o - (number of N reads before o); order_n ascending
  uint32_t lo = 0, hi = a.nr.num;
  while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.nr.order[mid] < o) lo = mid + 1; else hi = mid; }
  const bool is_n = lo < a.nr.num && a.nr.order[lo] == o;
  const uint64_t *codes; const uint64_t *nflag = nullptr; int len;
  if (is_n) { codes = a.nr.codes + (size_t)lo * a.W; nflag = a.nr.nflag + (size_t)lo * a.W; len = a.nr.lens[lo]; }
  else {
    const uint32_t k = o - lo;
    if (k >= a.num_clean) { if (lane == 0) atomicAdd(a.bad + 2, 1ull); return; }
    codes = a.reads + (size_t)k * a.W; len = a.lens[k];
  }
  const unsigned long long b0 = a.off[s], b1 = a.off[s + 1];
  if ((unsigned long long)len != b1 - b0) { if (lane == 0) atomicAdd(a.bad + 1, 1ull); return; }
  unsigned mism = 0;
  for (int j = lane; j < len; j += 32) {
    const int c = base_code(codes, j);
    char want = "AGCT"[c];
    if (nflag && ((nflag[j >> 5] >> (2 * (j & 31))) & 1ull)) want = 'N';
    if (a.bases[b0 + j] != (uint8_t)want) mism++;
  }
  mism = __reduce_add_sync(FULL, mism);
  if (lane == 0) {
    if (mism) atomicAdd(a.bad + 0, 1ull);
    atomicAdd(a.bad + 3, 1ull);
  }
}

}  // namespace

void run_verify(Ctx &c, const EncodeDev &e, const uint64_t *reads, const uint16_t *lens, uint32_t num_clean, int W,
                const NReads &nr, bool paired, bool preserve, uint32_t block, VerifyReport &out) {
  cudaStream_t st = c.stream;
  out = VerifyReport{};
  const uint32_t n = (uint32_t)e.num_reads;
  ReblockDev rb;
  run_reblock(c, e, paired, preserve, block, rb);
  DecodeDev dd;
  run_decode_blocks(c, rb, rb.size, e.seq_packed, e.seq_len, n, paired, preserve, block, dd);
  if (dd.num_reads != n) throw LimitError("verify: the block decode returned a different number of reads");
  VerifyArgs a{};
  a.bases = dd.bases; a.off = dd.offsets; a.slot = rb.order; a.order = e.order; a.n = n;
  a.reads = reads; a.lens = lens; a.num_clean = num_clean; a.W = W; a.nr = nr;
  const size_t bm = ((size_t)n + 31) / 32 + 1;
  a.seen = c.pool.dev<uint32_t>("vf.seen", bm);
  a.bad = c.pool.dev<unsigned long long>("vf.bad", 4);
  SB_CUDA(cudaMemsetAsync(a.seen, 0, bm * sizeof(uint32_t), st));
  SB_CUDA(cudaMemsetAsync(a.bad, 0, 4 * sizeof(unsigned long long), st));
  if (n) {
    k_verify<<<(uint32_t)(((uint64_t)n * 32 + 255) / 256), 256, 0, st>>>(a);
    c.launches++;
  }
  unsigned long long h[4];
  SB_CUDA(cudaMemcpyAsync(h, a.bad, sizeof(h), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_CUDA(cudaGetLastError());
  out.base_mismatch_reads = h[0]; out.length_mismatch_reads = h[1]; out.bad_order = h[2]; out.reads_checked = h[3];
  out.num_reads = n; out.num_blocks = rb.num_blocks; out.decoded_bases = dd.total;
  uint64_t blk = 0;
  for (int s = 0; s < RB_NSTREAMS; s++) blk += rb.size[s];
  out.block_stream_bytes = blk;
}

}  // namespace sb
