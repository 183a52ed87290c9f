// exchange.cu -- the one exchange of the multi-GPU path, inside the library.
//
// The reference is single-process (OpenMP only); its threads share one pool of reads.  Here a GPU plays the role of
// one reference thread, and reads can only ever match reads that share a dictionary window up to a shift and a strand
// flip (reorder.h:246-318), so every read is routed to the GPU that owns hash(strand-canonical 16-mer minimizer) mod G
// (bucket.cu) with ONE all-to-all(v) of {8W-byte row, u16 length, u32 global id} over NCCL / NVLink:
//
//   k_bucket_hist   thread per read: minimizer bucket (kept as one byte) + per-block histogram of the owners
//   one CUB scan    over the owner-major histogram -> where each block's reads of every owner go in the send regions
//   k_scatter_send  block per 256 reads: rank inside the block by __match_any_sync (stable: input order is kept inside
//                   every owner's region), rows / lengths / ids written straight into the per-destination send regions
//   ncclAllGather of the G send counts (so every rank knows what it receives), then ONE ncclGroup of
//   3 x (G - 1) ncclSend / ncclRecv pairs that land in the final arrays (the rank's own region moves with a D2D copy)
//
// and after the single-GPU path has run on the reads a rank owns:
//
//   k_finalize_shard  positions made absolute in the concatenation of all ranks' consensus (what the reference does for
//                     its thread shards, encoder.h:473-487), read_order.bin's local indices replaced by the global ids
//                     that travelled with the reads
//
// NCCL is loaded with dlopen on first use: single-GPU users of libspring_b200.so do not need it, and under PyTorch the
// already loaded libnccl.so.2 is shared.
#include <dlfcn.h>
#include <nccl.h>
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {

namespace {

struct NcclApi {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi &nccl() {
  static NcclApi api;
  if (api.h) return api;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.h) break;
  }
  if (!api.h) throw CudaError(std::string("multi-GPU path needs NCCL: dlopen(libnccl.so.2) failed: ") + dlerror());
  auto sym = [&](const char *s) {
    void *p = dlsym(api.h, s);
    if (!p) throw CudaError(std::string("NCCL symbol missing: ") + s);
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
  api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  return api;
}

#define SB_NCCL(expr)                                                                                                   \
  do {                                                                                                                  \
    ncclResult_t _r = (expr);                                                                                           \
    if (_r != ncclSuccess) throw sb::CudaError(std::string(#expr) + ": " + nccl().GetErrorString(_r));                   \
  } while (0)

constexpr int kBlock = 256;   // reads per block of the histogram / scatter passes
constexpr int kMaxWorld = 256;

// owner of every read (one byte) + histogram of the owners per block of 256 reads
__global__ void __launch_bounds__(kBlock) k_bucket_hist(const uint64_t *__restrict__ reads, const uint16_t *__restrict__ lens, uint32_t n,
                                                        int W, uint32_t world, uint8_t *__restrict__ owner, uint32_t *__restrict__ block_hist) {
  __shared__ uint32_t s_hist[kMaxWorld];
  for (uint32_t t = threadIdx.x; t < world; t += kBlock) s_hist[t] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i < n) {
    const uint32_t b = minimizer_bucket(reads + (size_t)i * W, lens[i], world);
    owner[i] = (uint8_t)b;
    atomicAdd(&s_hist[b], 1u);
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < world; t += kBlock) block_hist[(size_t)t * gridDim.x + blockIdx.x] = s_hist[t];  // owner-major
}

// The histogram is owner-major ([owner][block]), so ONE exclusive scan over it gives, for (owner b, block k), the reads
// of lower owners plus the reads of owner b in earlier blocks: the first destination of block k's reads of owner b.
__global__ void k_send_counts(const uint32_t *__restrict__ dst_base, uint32_t nblocks, uint32_t world, uint32_t n, uint32_t *send_count) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= world) return;
  const uint32_t lo = dst_base[(size_t)b * nblocks];
  const uint32_t hi = b + 1 < world ? dst_base[(size_t)(b + 1) * nblocks] : n;
  send_count[b] = hi - lo;
}

// block per 256 reads: destination = region start of the owner + reads of that owner in earlier blocks + earlier reads
// of that owner in this block (warp: __match_any_sync, block: per-warp counts in shared memory)
__global__ void __launch_bounds__(kBlock) k_scatter_send(const uint64_t *__restrict__ reads, const uint16_t *__restrict__ lens,
                                                         const uint32_t *__restrict__ ids, uint32_t n, int W, uint32_t world,
                                                         const uint8_t *__restrict__ owner, const uint32_t *__restrict__ dst_base,
                                                         uint64_t *__restrict__ s_rows,
                                                         uint16_t *__restrict__ s_lens, uint32_t *__restrict__ s_ids) {
  __shared__ uint32_t s_cnt[kBlock / 32][kMaxWorld];
  __shared__ uint32_t s_dst[kBlock];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t t = threadIdx.x; t < (kBlock / 32) * kMaxWorld; t += kBlock) (&s_cnt[0][0])[t] = 0;
  __syncthreads();
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  const bool in = i < n;
  const uint32_t b = in ? owner[i] : 0xFFFFFFFFu;
  const unsigned peers = __match_any_sync(0xFFFFFFFFu, b);
  const uint32_t rank_in_warp = (uint32_t)__popc(peers & ((1u << lane) - 1u));
  if (in && rank_in_warp == 0) s_cnt[warp][b] = (uint32_t)__popc(peers);
  __syncthreads();
  uint32_t before = 0;
  if (in) {
    for (int w2 = 0; w2 < warp; w2++) before += s_cnt[w2][b];
    s_dst[threadIdx.x] = dst_base[(size_t)b * gridDim.x + blockIdx.x] + before + rank_in_warp;
  }
  __syncthreads();
  if (in) {
    const uint32_t d = s_dst[threadIdx.x];
    s_lens[d] = lens[i];
    s_ids[d] = ids[i];
  }
  // rows: the block's 256 x W words are read coalesced; thread t moves words t, t + 256, ...
  const uint32_t first = blockIdx.x * kBlock, cnt = min((uint32_t)kBlock, n - first);
  for (uint32_t t = threadIdx.x; t < cnt * (uint32_t)W; t += kBlock) {
    const uint32_t r = t / (uint32_t)W, w = t - r * (uint32_t)W;
    s_rows[(size_t)s_dst[r] * W + w] = reads[(size_t)first * W + t];
  }
}

__global__ void k_finalize_shard(uint64_t *pos, uint64_t num_aligned, uint64_t seq_base, uint32_t *order, uint64_t num_reads,
                                 const uint32_t *__restrict__ ids, uint32_t n_owned, const uint32_t *__restrict__ n_ids, uint32_t n_n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < num_aligned) pos[i] += seq_base;
  if (i < num_reads) {
    const uint32_t o = order[i];
    order[i] = o < n_owned ? ids[o] : (o - n_owned < n_n ? n_ids[o - n_owned] : 0xFFFFFFFFu);
  }
}

// original FASTQ index of clean read k0 + i: clean reads keep their order, the reads with N (order_n, ascending) sit in
// between (encoder.cpp:177-222, the same map as corrected_order in encode.cu)
__global__ void k_original_ids(uint32_t k0, uint32_t n, const uint32_t *__restrict__ order_n, uint32_t nn, uint32_t *ids) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = k0 + i;
  uint32_t lo = 0, hi = nn;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (order_n[mid] - mid <= k) lo = mid + 1; else hi = mid;
  }
  ids[i] = k + lo;
}

}  // namespace

void run_original_ids(Ctx &c, uint32_t k0, uint32_t n, const uint32_t *d_order_n, uint32_t nn, uint32_t *ids) {
  if (!n) return;
  k_original_ids<<<(n + 255) / 256, 256, 0, c.stream>>>(k0, n, d_order_n, nn, ids);
  c.launches++;
  SB_CUDA(cudaGetLastError());
}

struct Comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

void comm_unique_id(uint8_t id[128]) {
  ncclUniqueId u;
  SB_NCCL(nccl().GetUniqueId(&u));
  memcpy(id, u.internal, 128);
}

Comm *comm_create(const uint8_t id[128], int rank, int world) {
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) throw CudaError("comm_init: bad rank / world (1..256 ranks)");
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  Comm *c = new Comm();
  c->rank = rank; c->world = world;
  ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
  if (r != ncclSuccess) { delete c; throw CudaError(std::string("ncclCommInitRank: ") + nccl().GetErrorString(r)); }
  return c;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int comm_rank(const Comm *c) { return c->rank; }
int comm_world(const Comm *c) { return c->world; }

void run_exchange(Ctx &c, Comm *cm, const uint64_t *reads, const uint16_t *lens, const uint32_t *ids, uint32_t n, int L, ExchangeDev &out) {
  if (!cm) throw CudaError("exchange: no communicator (spring_b200_comm_init)");
  cudaStream_t st = c.stream;
  out = ExchangeDev{};
  const int W = words_for(L);
  const uint32_t world = (uint32_t)cm->world, rank = (uint32_t)cm->rank;
  const uint32_t nn = n ? n : 1, nblocks = (n + kBlock - 1) / kBlock;
  uint8_t *owner = c.pool.dev<uint8_t>("xg.owner", nn);
  const size_t nh = (size_t)(nblocks ? nblocks : 1) * world;
  uint32_t *block_hist = c.pool.dev<uint32_t>("xg.block_hist", nh), *dst_base = c.pool.dev<uint32_t>("xg.dst_base", nh);
  uint32_t *send_count = c.pool.dev<uint32_t>("xg.counts", world);
  uint32_t *d_all = c.pool.dev<uint32_t>("xg.all_counts", (size_t)world * world);     // [src][dst]
  uint32_t *h_all = c.pool.pin<uint32_t>("xg.h_all_counts", (size_t)world * world);
  uint64_t *s_rows = c.pool.dev<uint64_t>("xg.s_rows", (size_t)nn * W);
  uint16_t *s_lens = c.pool.dev<uint16_t>("xg.s_lens", nn);
  uint32_t *s_ids = c.pool.dev<uint32_t>("xg.s_ids", nn);
  size_t cub_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, block_hist, dst_base, (int)nh, st);
  void *cub_tmp = c.pool.device("xg.cubtmp", cub_bytes);
  if (!c.ev_x0) { SB_CUDA(cudaEventCreate(&c.ev_x0)); SB_CUDA(cudaEventCreate(&c.ev_x1)); }
  SB_CUDA(cudaEventRecord(c.ev_x0, st));
  if (n) {
    k_bucket_hist<<<nblocks, kBlock, 0, st>>>(reads, lens, n, W, world, owner, block_hist);
    cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, block_hist, dst_base, (int)nh, st);
    k_send_counts<<<(world + 63) / 64, 64, 0, st>>>(dst_base, nblocks, world, n, send_count);
    c.launches += 3;
  } else {
    SB_CUDA(cudaMemsetAsync(send_count, 0, world * sizeof(uint32_t), st));
  }
  // everybody's send counts: all[src][dst]
  SB_NCCL(nccl().AllGather(send_count, d_all, world, ncclUint32, cm->comm, st));
  SB_CUDA(cudaMemcpyAsync(h_all, d_all, sizeof(uint32_t) * world * world, cudaMemcpyDeviceToHost, st));
  if (n) {  // the scatter runs while the counts travel
    k_scatter_send<<<nblocks, kBlock, 0, st>>>(reads, lens, ids, n, W, world, owner, dst_base, s_rows, s_lens, s_ids);
    c.launches++;
  }
  SB_CUDA(cudaStreamSynchronize(st));  // the counts are host arguments of ncclSend / ncclRecv
  std::vector<uint64_t> soff(world + 1, 0), roff(world + 1, 0);
  for (uint32_t p = 0; p < world; p++) {
    soff[p + 1] = soff[p] + h_all[(size_t)rank * world + p];
    roff[p + 1] = roff[p] + h_all[(size_t)p * world + rank];
  }
  if (soff[world] != n) throw CudaError("exchange: send counts do not add up");
  const uint64_t nr = roff[world];
  if (nr >= 0x7FFFFFF0ull) throw LimitError("exchange: a rank would own >= 2^31 reads");
  const size_t nrn = nr ? nr : 1;
  uint64_t *r_rows = c.pool.dev<uint64_t>("xg.r_rows", nrn * W);
  uint16_t *r_lens = c.pool.dev<uint16_t>("xg.r_lens", nrn);
  uint32_t *r_ids = c.pool.dev<uint32_t>("xg.r_ids", nrn);
  SB_NCCL(nccl().GroupStart());
  for (uint32_t p = 0; p < world; p++) {
    if (p == rank) continue;
    const uint64_t sc = soff[p + 1] - soff[p], rc = roff[p + 1] - roff[p];
    if (sc) {
      SB_NCCL(nccl().Send(s_rows + soff[p] * W, sc * W, ncclUint64, (int)p, cm->comm, st));
      SB_NCCL(nccl().Send(s_lens + soff[p], sc * 2, ncclUint8, (int)p, cm->comm, st));
      SB_NCCL(nccl().Send(s_ids + soff[p], sc, ncclUint32, (int)p, cm->comm, st));
    }
    if (rc) {
      SB_NCCL(nccl().Recv(r_rows + roff[p] * W, rc * W, ncclUint64, (int)p, cm->comm, st));
      SB_NCCL(nccl().Recv(r_lens + roff[p], rc * 2, ncclUint8, (int)p, cm->comm, st));
      SB_NCCL(nccl().Recv(r_ids + roff[p], rc, ncclUint32, (int)p, cm->comm, st));
    }
  }
  SB_NCCL(nccl().GroupEnd());
  const uint64_t self = soff[rank + 1] - soff[rank];
  if (self) {
    SB_CUDA(cudaMemcpyAsync(r_rows + roff[rank] * W, s_rows + soff[rank] * W, self * W * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    SB_CUDA(cudaMemcpyAsync(r_lens + roff[rank], s_lens + soff[rank], self * sizeof(uint16_t), cudaMemcpyDeviceToDevice, st));
    SB_CUDA(cudaMemcpyAsync(r_ids + roff[rank], s_ids + soff[rank], self * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  SB_CUDA(cudaEventRecord(c.ev_x1, st));
  SB_CUDA(cudaGetLastError());
  out.reads = r_rows; out.lens = r_lens; out.ids = r_ids; out.n = (uint32_t)nr;
  out.sent_to_peers = n - self; out.received_from_peers = nr - self;
}

float exchange_ms(Ctx &c) {
  float t = 0;
  if (c.ev_x0 && cudaEventSynchronize(c.ev_x1) == cudaSuccess) cudaEventElapsedTime(&t, c.ev_x0, c.ev_x1);
  return t;
}

void run_finalize_shard(Ctx &c, Comm *cm, EncodeDev &e, const uint32_t *ids, uint32_t n_owned, const uint32_t *h_n_ids, uint32_t n_n,
                        ShardLayout &out) {
  if (!cm) throw CudaError("finalize: no communicator (spring_b200_comm_init)");
  cudaStream_t st = c.stream;
  const uint32_t world = (uint32_t)cm->world, rank = (uint32_t)cm->rank;
  if ((uint64_t)n_owned + n_n != e.num_reads) throw CudaError("finalize: ids do not cover the shard's reads");
  // sizes of every shard: {seq_len, num_aligned, num_reads, noise_bytes, num_noise, unaligned_bytes, unaligned_len, 0}
  unsigned long long *d_sz = c.pool.dev<unsigned long long>("xg.sizes", 8 * ((size_t)world + 1));
  unsigned long long *h_sz = c.pool.pin<unsigned long long>("xg.h_sizes", 8 * ((size_t)world + 1));
  unsigned long long mine[8] = {e.seq_len, e.num_aligned, e.num_reads, e.noise_bytes, e.num_noise, e.unaligned_bytes, e.unaligned_len, 0};
  SB_CUDA(cudaMemcpyAsync(d_sz + 8 * (size_t)world, mine, sizeof(mine), cudaMemcpyHostToDevice, st));
  SB_NCCL(nccl().AllGather(d_sz + 8 * (size_t)world, d_sz, 8, ncclUint64, cm->comm, st));
  SB_CUDA(cudaMemcpyAsync(h_sz, d_sz, sizeof(unsigned long long) * 8 * world, cudaMemcpyDeviceToHost, st));
  uint32_t *d_nids = c.pool.dev<uint32_t>("xg.n_ids", (size_t)n_n + 1);
  if (n_n) SB_CUDA(cudaMemcpyAsync(d_nids, h_n_ids, sizeof(uint32_t) * n_n, cudaMemcpyHostToDevice, st));
  SB_CUDA(cudaStreamSynchronize(st));
  out = ShardLayout{};
  out.rank = rank; out.world = world;
  for (uint32_t p = 0; p < world; p++) {
    const unsigned long long *s = h_sz + 8 * (size_t)p;
    if (p < rank) {
      out.seq_base += s[0]; out.aligned_before += s[1]; out.noise_before += s[3]; out.num_noise_before += s[4];
      out.unaligned_reads_before += s[2] - s[1]; out.unaligned_bytes_before += s[5];
    }
    out.total_seq_len += s[0]; out.total_aligned += s[1]; out.total_reads += s[2]; out.total_noise_bytes += s[3];
    out.total_num_noise += s[4]; out.total_unaligned_bytes += s[5]; out.total_unaligned_len += s[6];
  }
  const uint64_t m = e.num_reads > e.num_aligned ? e.num_reads : e.num_aligned;
  if (m) {
    k_finalize_shard<<<(uint32_t)((m + 255) / 256), 256, 0, st>>>(e.pos, e.num_aligned, out.seq_base, e.order, e.num_reads, ids, n_owned, d_nids, n_n);
    c.launches++;
  }
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
