// dict.cu -- dictionary construction in HBM.
//
// Replaces constructdictionary<> (reference src/bitset_util.h:74-221): window key per read -> (drop reads shorter
// than the window) -> sort -> unique keys -> bins with ascending read ids.  The reference maps key -> bin through
// BooPHF and a CSR startpos[]; here the unique keys go into an open-addressing table of 32-byte slots {hashed key, bin
// start, live count, bin size, three highest ids} (one 32 B sector per probe), and read_id[] is the value array of a
// stable radix sort of (hashed key, read id) pairs, so ids are ascending inside every bin exactly as
// bitset_util.h:188-206 leaves them.
//
// The sort key is hk = mix64(key), a bijection of the 64-bit key, and a slot's home is the top bits of hk: sorted order
// IS table order.  Slot of the k-th unique key = k + max_{j <= k}(home_j - j) -- the first free slot at or after its
// home when the keys go in ascending -- is one running-max scan; the table, the bins and the key filter (word = top
// bits of hk as well) are then written front to back with plain coalesced stores, no atomics, no probing.
//
// The radix sort runs on the TOP 32 bits of hk only (4 passes instead of 8), with (low 32 bits of hk, read id) as its
// 64-bit value.  Two different keys share their top 32 bits about once per 2^32 / n keys (1 % of the keys at 100 M); such a
// run comes out of the stable sort ordered by read id with the keys interleaved, and one thread re-orders it by the
// low 32 bits (stable insertion sort; k_fix_runs) before the bins are cut.  The value 0xFFFFFFFF of the top bits is
// reserved for reads that are not indexed.
//
// No host synchronisation: reads that are not indexed get hk = ~0 and sort behind everything, the number of unique keys
// stays on the device (kernels are launched over the upper bound n and read it there), the table is sized from n.
#include <climits>
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {

namespace {

struct MaxOp {
  __device__ int operator()(int a, int b) const { return a > b ? a : b; }
};

// hk = mix64((read & mask1) >> 2*start) (bitset_util.h:93-94) as sort key k32 = hk >> 32 and value (hk << 32) | read id;
// k32 = 0xFFFFFFFF marks a read that is not indexed (too short for the window, N inside it -- or, once per 2^32 keys, a
// real key whose top bits are all ones: its reads simply stay unindexed).  One thread per read; a warp touches 32
// consecutive rows of W words (coalesced across the warp's combined footprint).
__global__ void k_extract_keys(const uint64_t *__restrict__ reads, const uint16_t *__restrict__ lens,
                               const uint64_t *__restrict__ nflag, uint32_t n, int W, int start, int end,
                               uint32_t *__restrict__ k32, uint64_t *__restrict__ val, uint32_t *__restrict__ num_valid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (i < n) {
    const uint64_t *r = reads + (size_t)i * W;
    const int nbits = 2 * (end - start + 1);
    const uint64_t hk = mix64(extract_bits(r, W, 2 * start, nbits));
    ok = lens[i] > end;  // bitset_util.h:99-105
    if (ok && nflag) ok = extract_bits(nflag + (size_t)i * W, W, 2 * start, nbits) == 0;
    if ((uint32_t)(hk >> 32) == 0xFFFFFFFFu) ok = false;
    k32[i] = ok ? (uint32_t)(hk >> 32) : 0xFFFFFFFFu;
    val[i] = (ok ? hk << 32 : 0xFFFFFFFF00000000ull) | i;
  }
  const int c = __syncthreads_count(ok);
  if (threadIdx.x == 0 && c) atomicAdd(num_valid, (uint32_t)c);
}

// The sort looked at the top `32 - rshift` bits only (rshift = 0 unless SPRING_B200_DICT_SORT_BITS says otherwise: a
// test knob that makes multi-key runs common at any input size).  A run of entries equal in those bits is in read-id
// order; if it holds several keys out of order -- thread i sits on such an inversion -- the first thread to take the
// run's lock (at the run's head) re-orders the whole run by the full key, stably, so read ids stay ascending inside
// every key.  Runs that are already in order -- all but ~n / 2^32 of them at 32 bits -- cost two loads per entry.
__device__ __forceinline__ uint64_t full_key(uint32_t k, uint64_t v) { return ((uint64_t)k << 32) | (v >> 32); }
__global__ void k_fix_runs(uint32_t *k32, uint64_t *val, uint32_t n, int rshift, uint8_t *lock) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 || i >= n) return;
  const uint32_t k = k32[i], run = k >> rshift;
  if (rshift == 0 && k == 0xFFFFFFFFu) return;  // the unindexed reads: one homogeneous run, possibly huge
  if (run != (k32[i - 1] >> rshift) || full_key(k, val[i]) >= full_key(k32[i - 1], val[i - 1])) return;
  uint32_t h = i - 1;
  while (h > 0 && (k32[h - 1] >> rshift) == run) h--;
  // one byte per entry, four entries per word: set the byte of the run's head
  unsigned int *w = reinterpret_cast<unsigned int *>(lock + (h & ~3u));
  const unsigned int bit = 1u << (8 * (h & 3u));
  if (atomicOr(w, bit) & bit) return;  // another thread of this run has it
  uint32_t e = i + 1;
  while (e < n && (k32[e] >> rshift) == run) e++;
  for (uint32_t a = h + 1; a < e; a++) {  // stable insertion sort of the pairs [h, e) by their full key
    const uint32_t xk = k32[a];
    const uint64_t xv = val[a], x = full_key(xk, xv);
    uint32_t b = a;
    while (b > h && full_key(k32[b - 1], val[b - 1]) > x) { k32[b] = k32[b - 1]; val[b] = val[b - 1]; b--; }
    k32[b] = xk; val[b] = xv;
  }
}

// sorted (k32, value) -> hashed keys hk (unindexed: ~0), read ids, and the heads of the runs of equal hk = the bins
__global__ void k_unpack_heads(const uint32_t *__restrict__ k32, const uint64_t *__restrict__ val, uint32_t n,
                               uint64_t *__restrict__ hkeys, uint32_t *__restrict__ rids, uint8_t *__restrict__ head,
                               uint32_t *__restrict__ head32) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t k = k32[i];
  const uint64_t v = val[i];
  const uint64_t hk = k == 0xFFFFFFFFu ? kInvalidKey : ((uint64_t)k << 32) | (v >> 32);
  uint32_t h = 0;
  if (k != 0xFFFFFFFFu) h = (i == 0 || k != k32[i - 1] || (uint32_t)(v >> 32) != (uint32_t)(val[i - 1] >> 32)) ? 1u : 0u;
  hkeys[i] = hk;
  rids[i] = (uint32_t)v;
  head[i] = (uint8_t)h;
  head32[i] = h;
}

// home_k - k for the running max (INT_MIN beyond the last key)
__global__ void k_home_minus_rank(const uint64_t *__restrict__ hkeys, const uint32_t *__restrict__ bin_start_idx,
                                  const uint32_t *__restrict__ numkeys, uint32_t n, int slot_shift, int *__restrict__ v) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  v[k] = k < *numkeys ? (int)slot_home(hkeys[bin_start_idx[k]], slot_shift) - (int)k : INT_MIN;
}

// bin k (sorted entries [s, e)) gets header index s + k in bins[], slot k + m[k] of the table and its filter bits
__global__ void k_insert_slots(const uint64_t *__restrict__ hkeys, const uint32_t *__restrict__ rid_sorted,
                               const uint32_t *__restrict__ bin_start_idx, const uint32_t *__restrict__ numkeys_p,
                               const uint32_t *__restrict__ num_valid, const int *__restrict__ m, uint32_t n, uint32_t slot_limit,
                               uint32_t filter_words, DictSlot *slots, uint32_t *bins, uint32_t *slot_of_bin, uint32_t *filter,
                               uint32_t *dropped) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t numkeys = *numkeys_p;
  if (k >= n || k >= numkeys) return;
  const uint32_t s = bin_start_idx[k];
  const uint32_t e = (k + 1 < numkeys) ? bin_start_idx[k + 1] : *num_valid;
  const uint64_t hk = hkeys[s];
  const long long slot = (long long)k + m[k];
  bins[s + k] = e - s;
  if (slot >= (long long)slot_limit) {  // a probe chain longer than the spare slots: the bin stays unindexed (ratio only)
    slot_of_bin[k] = 0xFFFFFFFFu;
    atomicAdd(dropped, 1u);
    return;
  }
  DictSlot sl;
  sl.key = hk; sl.start1 = s + k + 1; sl.live = e - s; sl.count = e - s;
  sl.rid[0] = rid_sorted[e - 1];
  sl.rid[1] = e - s > 1 ? rid_sorted[e - 2] : 0;
  sl.rid[2] = e - s > 2 ? rid_sorted[e - 3] : 0;
  uint4 *dst = reinterpret_cast<uint4 *>(slots + slot);
  dst[0] = make_uint4((uint32_t)sl.key, (uint32_t)(sl.key >> 32), sl.start1, sl.live);
  dst[1] = make_uint4(sl.count, sl.rid[0], sl.rid[1], sl.rid[2]);
  slot_of_bin[k] = (uint32_t)slot;
  // keys arrive in ascending hk and the filter word is the top bits of hk: neighbouring threads hit neighbouring words
  atomicOr(filter + filter_word(hk, filter_words), filter_bits(hk));
}

// sorted entry i (ascending id inside its bin) -> descending position behind the bin header
__global__ void k_fill_bins(const uint64_t *__restrict__ hkeys, const uint32_t *__restrict__ rid_sorted, const uint32_t *__restrict__ kidx1,
                            const uint32_t *__restrict__ bin_start_idx, const uint32_t *__restrict__ slot_of_bin,
                            const uint32_t *__restrict__ numkeys_p, const uint32_t *__restrict__ num_valid, uint32_t n, uint32_t *bins,
                            uint32_t *slot_of_read) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t rid = rid_sorted[i];
  if (hkeys[i] == kInvalidKey) { slot_of_read[rid] = 0xFFFFFFFFu; return; }
  const uint32_t numkeys = *numkeys_p;
  const uint32_t k = kidx1[i] - 1;
  const uint32_t s = bin_start_idx[k];
  const uint32_t e = (k + 1 < numkeys) ? bin_start_idx[k + 1] : *num_valid;
  bins[s + k + 1 + (e - 1 - i)] = rid;
  slot_of_read[rid] = slot_of_bin[k];
}

inline uint32_t grid_for(uint64_t n, int block) { return (uint32_t)((n + block - 1) / block); }
inline int log2_pow2(uint64_t v) { int b = 0; while ((1ull << b) < v) b++; return b; }

}  // namespace

void build_dictionary(Ctx &c, const uint64_t *reads, const uint16_t *lens, const uint64_t *nflag, uint32_t n, int W,
                      int start, int end, const char *tag, DictBuild &out) {
  std::string t(tag);
  auto nm = [&](const char *s) { return t + s; };
  cudaStream_t st = c.stream;
  out = DictBuild{};
  out.view.start = start;
  out.view.end = end;
  out.view.key_bits = 2 * (end - start + 1);
  const uint32_t nn = n ? n : 1;
  uint64_t *val_a = c.pool.dev<uint64_t>(nm(".val_a").c_str(), nn), *val_b = c.pool.dev<uint64_t>(nm(".val_b").c_str(), nn);
  uint32_t *k32_a = c.pool.dev<uint32_t>(nm(".k32_a").c_str(), nn), *k32_b = c.pool.dev<uint32_t>(nm(".k32_b").c_str(), nn);
  uint64_t *keys_b = val_a;   // the sort's input buffers are free again when the sorted pairs are unpacked
  uint32_t *rid_c = k32_a;
  uint32_t *bin_start_idx = c.pool.dev<uint32_t>(nm(".bin_start").c_str(), nn);
  uint8_t *flag = c.pool.dev<uint8_t>(nm(".flag").c_str(), (size_t)nn + 4);
  uint32_t *head32 = c.pool.dev<uint32_t>(nm(".head32").c_str(), nn);
  uint32_t *kidx1 = c.pool.dev<uint32_t>(nm(".kidx1").c_str(), nn);
  int *hm = c.pool.dev<int>(nm(".home_minus").c_str(), nn), *hmax = c.pool.dev<int>(nm(".home_max").c_str(), nn);
  uint32_t *slot_of_bin = c.pool.dev<uint32_t>(nm(".slot_of_bin").c_str(), nn);
  uint32_t *slot_of_read = c.pool.dev<uint32_t>(nm(".slot_of_read").c_str(), nn);
  uint32_t *bins = c.pool.dev<uint32_t>(nm(".bins").c_str(), 2 * (size_t)nn);
  uint32_t *skip = c.pool.dev<uint32_t>(nm(".skip").c_str(), 2 * (size_t)nn);
  uint32_t *d_count = c.pool.dev<uint32_t>(nm(".count").c_str(), 4);  // [0] indexed reads, [1] unique keys, [2] dropped bins
  SB_CUDA(cudaMemsetAsync(skip, 0, 2 * (size_t)nn * sizeof(uint32_t), st));
  SB_CUDA(cudaMemsetAsync(d_count, 0, 4 * sizeof(uint32_t), st));
  // sized from the upper bound n: load factor <= 1/2 whatever the number of unique keys turns out to be
  if (n > (1u << 30)) throw LimitError("dictionary: more than 2^30 reads in one GPU shard");
  uint64_t cap = 16;
  while (cap < 2ull * n) cap <<= 1;
  out.capacity = (uint32_t)cap;
  out.view.slot_shift = 64 - log2_pow2(cap);
  DictSlot *slots = c.pool.dev<DictSlot>(nm(".slots").c_str(), (size_t)cap + kSlotPad);
  SB_CUDA(cudaMemsetAsync(slots, 0, sizeof(DictSlot) * ((size_t)cap + kSlotPad), st));
  // 32-bit words, 2 bits set per key; SPRING_B200_FILTER_BITS bits per read (default 8: ~5 % false positives).  The size is
  // not rounded to a power of two: two filters of 100 M keys are 200 MB then, not 256 -- every MB counts against a 126 MB L2.
  static const double kFilterBitsPerKey = getenv("SPRING_B200_FILTER_BITS") ? atof(getenv("SPRING_B200_FILTER_BITS")) : 8.0;
  uint64_t fwords = (uint64_t)(kFilterBitsPerKey * (double)n / 32.0) + 1;
  if (fwords < 2048) fwords = 2048;
  if (fwords > 0x7FFFFFFFull) fwords = 0x7FFFFFFFull;
  uint32_t *filter = c.pool.dev<uint32_t>(nm(".filter").c_str(), fwords);
  SB_CUDA(cudaMemsetAsync(filter, 0, fwords * sizeof(uint32_t), st));
  out.view.filter = filter;
  out.view.filter_words = (uint32_t)fwords;
  out.view.slots = slots;
  out.view.bins = bins;
  out.view.slot_of_read = slot_of_read;
  out.view.skip = skip;
  out.sorted_keys = keys_b;
  out.bin_start_idx = bin_start_idx;
  out.sorted_rids = rid_c;
  out.d_counts = d_count;
  if (n == 0) return;

  size_t tmp_bytes = 0, need = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, need, k32_a, k32_b, val_a, val_b, (int)n, 0, 32, st); tmp_bytes = need;
  cub::DeviceScan::InclusiveSum(nullptr, need, head32, kidx1, (int)n, st); if (need > tmp_bytes) tmp_bytes = need;
  cub::DeviceScan::InclusiveScan(nullptr, need, hm, hmax, MaxOp(), (int)n, st); if (need > tmp_bytes) tmp_bytes = need;
  {
    cub::CountingInputIterator<uint32_t> cnt(0);
    cub::DeviceSelect::Flagged(nullptr, need, cnt, flag, bin_start_idx, d_count + 1, (int)n, st);
    if (need > tmp_bytes) tmp_bytes = need;
  }
  void *tmp = c.pool.device(nm(".cubtmp").c_str(), tmp_bytes);

  k_extract_keys<<<grid_for(n, 256), 256, 0, st>>>(reads, lens, nflag, n, W, start, end, k32_a, val_a, d_count);
  // stable LSD radix sort on the top 32 bits: read ids stay ascending inside equal keys; unindexed reads end up last
  need = tmp_bytes;
  static const int kSortBits = [] {
    const char *e = getenv("SPRING_B200_DICT_SORT_BITS");
    const int b = e ? atoi(e) : 32;
    return b < 1 ? 1 : (b > 32 ? 32 : b);
  }();
  cub::DeviceRadixSort::SortPairs(tmp, need, k32_a, k32_b, val_a, val_b, (int)n, 32 - kSortBits, 32, st);
  SB_CUDA(cudaMemsetAsync(flag, 0, ((size_t)nn + 3) & ~(size_t)3, st));  // run locks of k_fix_runs
  k_fix_runs<<<grid_for(n, 256), 256, 0, st>>>(k32_b, val_b, n, 32 - kSortBits, flag);
  k_unpack_heads<<<grid_for(n, 256), 256, 0, st>>>(k32_b, val_b, n, keys_b, rid_c, flag, head32);
  need = tmp_bytes;
  cub::DeviceScan::InclusiveSum(tmp, need, head32, kidx1, (int)n, st);
  cub::CountingInputIterator<uint32_t> cnt(0);
  need = tmp_bytes;
  cub::DeviceSelect::Flagged(tmp, need, cnt, flag, bin_start_idx, d_count + 1, (int)n, st);
  k_home_minus_rank<<<grid_for(n, 256), 256, 0, st>>>(keys_b, bin_start_idx, d_count + 1, n, out.view.slot_shift, hm);
  need = tmp_bytes;
  cub::DeviceScan::InclusiveScan(tmp, need, hm, hmax, MaxOp(), (int)n, st);
  k_insert_slots<<<grid_for(n, 256), 256, 0, st>>>(keys_b, rid_c, bin_start_idx, d_count + 1, d_count, hmax, n, (uint32_t)cap + kSlotPad - 1,
                                                  out.view.filter_words, slots, bins, slot_of_bin, filter, d_count + 2);
  k_fill_bins<<<grid_for(n, 256), 256, 0, st>>>(keys_b, rid_c, kidx1, bin_start_idx, slot_of_bin, d_count + 1, d_count, n, bins, slot_of_read);
  c.launches += 6 + (2 + 4) + 2 + 3 + 2;  // ours + CUB (sort: histogram + 4 x onesweep, scans, select)
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
