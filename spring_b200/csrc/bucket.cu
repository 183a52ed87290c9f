// bucket.cu -- multi-GPU partitioning key: a strand-canonical minimizer bucket per read.
//
// The reference has no distributed path; reads can only ever match reads
// that share a dictionary window up to a shift and a strand flip (reorder.h:246-318), so the
// multi-GPU path routes every read to the GPU that owns hash(min over its k-mers of the canonical
// k-mer) mod G: a read, its reverse complement and its shifted neighbours mostly agree on that
// minimizer, so overlapping reads meet on one GPU after a single all-to-all.  The bucket only
// decides which reads can meet (compression ratio), never correctness.
#include "kernels.cuh"

namespace sb {
namespace {
constexpr int kK = 16;  // k-mer length (32-bit k-mers)

__global__ void k_bucket(const uint64_t *__restrict__ reads, const uint16_t *__restrict__ lens, uint32_t n, int W,
                         uint32_t num_buckets, uint32_t *__restrict__ bucket) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t *r = reads + (size_t)i * W;
  const int len = lens[i];
  uint64_t best = ~0ull;
  uint32_t fwd = 0, rc = 0;
  uint64_t w = 0;
  for (int j = 0; j < len; j++) {
    if ((j & 31) == 0) w = r[j >> 5];
    const uint32_t c = (uint32_t)(w & 3ull);
    w >>= 2;
    fwd = (fwd << 2) | c;                      // kK = 16 bases fill the 32-bit word exactly
    rc = (rc >> 2) | ((3u - c) << (2 * (kK - 1)));
    if (j >= kK - 1) {
      const uint64_t h = mix64((uint64_t)(fwd < rc ? fwd : rc));
      best = h < best ? h : best;
    }
  }
  if (len < kK) best = mix64((uint64_t)len);
  bucket[i] = (uint32_t)((best >> 16) % num_buckets);
}
}  // namespace

void bucket_reads(Ctx &c, const uint64_t *reads, const uint16_t *lens, uint32_t n, int L, uint32_t num_buckets, uint32_t *bucket) {
  if (!n) return;
  k_bucket<<<(n + 255) / 256, 256, 0, c.stream>>>(reads, lens, n, words_for(L), num_buckets, bucket);
  c.launches++;
  SB_CUDA(cudaGetLastError());
}
}  // namespace sb
