// dict.cu -- dictionary construction in HBM.
//
// Replaces constructdictionary<> (reference src/bitset_util.h:74-221): window key per read ->
// (drop reads shorter than the window) -> sort -> unique keys -> bins with ascending read ids.
// The reference maps key -> bin through BooPHF and a CSR startpos[]; here the unique keys go into
// an open-addressing table of 32-byte slots {key, bin start, live count, bin size, three highest ids}
// (one 32 B sector per probe),
// and read_id[] is the value array of a stable radix sort of (key, read id) pairs, so ids are
// ascending inside every bin exactly as bitset_util.h:188-206 leaves them.
#include <cub/cub.cuh>
#include "kernels.cuh"

namespace sb {

// key = ((read & mask1) >> 2*start).to_ullong()  (bitset_util.h:93-94).  One thread per read; a warp
// touches 32 consecutive rows of W words (coalesced across the warp's combined footprint).
__global__ void k_extract_keys(const uint64_t *__restrict__ reads, const uint16_t *__restrict__ lens,
                               const uint64_t *__restrict__ nflag, uint32_t n, int W, int start, int end,
                               uint64_t *__restrict__ keys, uint32_t *__restrict__ rids, uint8_t *__restrict__ valid) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t *r = reads + (size_t)i * W;
  int nbits = 2 * (end - start + 1);
  uint64_t key = extract_bits(r, W, 2 * start, nbits);
  bool ok = lens[i] > end;  // bitset_util.h:99-105
  if (ok && nflag) ok = extract_bits(nflag + (size_t)i * W, W, 2 * start, nbits) == 0;
  keys[i] = key;
  rids[i] = i;
  valid[i] = ok ? 1 : 0;
}

__global__ void k_mark_heads(const uint64_t *__restrict__ keys, uint32_t n, uint8_t *__restrict__ head,
                             uint32_t *__restrict__ head32) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
  head[i] = (uint8_t)h;
  head32[i] = h;
}

// bin k (sorted entries [s, e)) gets header index s + k in bins[] and one slot in the table
__global__ void k_insert_slots(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ bin_start_idx,
                               uint32_t numkeys, uint32_t n_valid, DictSlot *slots, uint32_t mask, uint32_t *bins,
                               uint32_t *slot_of_bin) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= numkeys) return;
  uint32_t s = bin_start_idx[k];
  uint32_t e = (k + 1 < numkeys) ? bin_start_idx[k + 1] : n_valid;
  uint64_t key = keys[s];
  uint32_t h = (uint32_t)mix64(key) & mask;
  bins[s + k] = e - s;
  for (;;) {
    if (atomicCAS(&slots[h].start1, 0u, s + k + 1) == 0u) {  // keys are unique: no key compare needed
      slots[h].key = key;
      slots[h].live = e - s;
      slots[h].count = e - s;
      slot_of_bin[k] = h;
      return;
    }
    h = (h + 1) & mask;
  }
}

// key filter: 2 bits in one word per key, >= 8 bits per key => ~5 % false positives in half the
// footprint of a 1-hash bitmap (32 MB for both dictionaries of 10 M reads: fits one L2 partition)
__global__ void k_set_filter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ bin_start_idx, uint32_t numkeys,
                             uint32_t *filter, uint32_t filter_mask) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= numkeys) return;
  const uint64_t hk = mix64(keys[bin_start_idx[k]]);
  atomicOr(filter + filter_word(hk, filter_mask), filter_bits(hk));
}

// sorted entry i (ascending id inside its bin) -> descending position behind the bin header
__global__ void k_fill_bins(const uint32_t *__restrict__ rid_sorted, const uint32_t *__restrict__ kidx1,
                            const uint32_t *__restrict__ bin_start_idx, const uint32_t *__restrict__ slot_of_bin,
                            uint32_t numkeys, uint32_t n_valid, uint32_t *bins, uint32_t *slot_of_read, DictSlot *slots) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_valid) return;
  const uint32_t k = kidx1[i] - 1;
  const uint32_t s = bin_start_idx[k];
  const uint32_t e = (k + 1 < numkeys) ? bin_start_idx[k + 1] : n_valid;
  const uint32_t rid = rid_sorted[i];
  const uint32_t dpos = e - 1 - i, sl = slot_of_bin[k];
  bins[s + k + 1 + dpos] = rid;
  slot_of_read[rid] = sl;
  if (dpos < 3) slots[sl].rid[dpos] = rid;
}

static inline uint32_t grid_for(uint64_t n, int block) { return (uint32_t)((n + block - 1) / block); }

void build_dictionary(Ctx &c, const uint64_t *reads, const uint16_t *lens, const uint64_t *nflag, uint32_t n, int W,
                      int start, int end, const char *tag, DictBuild &out) {
  std::string t(tag);
  auto nm = [&](const char *s) { return t + s; };
  cudaStream_t st = c.stream;
  out = DictBuild{};
  out.view.start = start;
  out.view.end = end;
  out.view.key_bits = 2 * (end - start + 1);
  uint32_t nn = n ? n : 1;
  uint64_t *keys_a = c.pool.dev<uint64_t>(nm(".keys_a").c_str(), nn);
  uint64_t *keys_b = c.pool.dev<uint64_t>(nm(".keys_b").c_str(), nn);
  uint32_t *rid_a = c.pool.dev<uint32_t>(nm(".rid_a").c_str(), nn);
  uint32_t *rid_b = c.pool.dev<uint32_t>(nm(".rid_b").c_str(), nn);
  uint32_t *rid_c = c.pool.dev<uint32_t>(nm(".rid_c").c_str(), nn);
  uint8_t *flag = c.pool.dev<uint8_t>(nm(".flag").c_str(), nn);
  uint32_t *head32 = c.pool.dev<uint32_t>(nm(".head32").c_str(), nn);
  uint32_t *kidx1 = c.pool.dev<uint32_t>(nm(".kidx1").c_str(), nn);
  uint32_t *slot_of_bin = c.pool.dev<uint32_t>(nm(".slot_of_bin").c_str(), nn);
  uint32_t *slot_of_read = c.pool.dev<uint32_t>(nm(".slot_of_read").c_str(), nn);
  uint32_t *bins = c.pool.dev<uint32_t>(nm(".bins").c_str(), 2 * (size_t)nn);
  uint32_t *skip = c.pool.dev<uint32_t>(nm(".skip").c_str(), 2 * (size_t)nn);
  SB_CUDA(cudaMemsetAsync(skip, 0, 2 * (size_t)nn * sizeof(uint32_t), st));
  out.view.skip = skip;
  uint32_t *d_count = c.pool.dev<uint32_t>(nm(".count").c_str(), 4);
  uint32_t *h_count = c.pool.pin<uint32_t>(nm(".hcount").c_str(), 4);
  if (n == 0) {
    uint32_t *filter0 = c.pool.dev<uint32_t>(nm(".filter").c_str(), 2048);
    SB_CUDA(cudaMemsetAsync(filter0, 0, 2048 * sizeof(uint32_t), st));
    out.view.filter = filter0; out.view.filter_mask = 2047;
    out.capacity = 16;
    DictSlot *slots = c.pool.dev<DictSlot>(nm(".slots").c_str(), out.capacity);
    SB_CUDA(cudaMemsetAsync(slots, 0, sizeof(DictSlot) * out.capacity, st));
    out.view.slots = slots; out.view.slot_mask = out.capacity - 1; out.view.bins = bins; out.view.slot_of_read = slot_of_read;
    out.sorted_keys = keys_b; out.bin_start_idx = rid_a; out.sorted_rids = rid_c;
    return;
  }
  k_extract_keys<<<grid_for(n, 256), 256, 0, st>>>(reads, lens, nflag, n, W, start, end, keys_a, rid_a, flag);
  c.launches++;
  // compaction of reads shorter than the window (variable-length input only)
  size_t tmp_bytes = 0, need = 0;
  cub::DeviceSelect::Flagged(nullptr, need, keys_a, flag, keys_b, d_count, (int)n, st); tmp_bytes = need;
  cub::DeviceSelect::Flagged(nullptr, need, rid_a, flag, rid_b, d_count, (int)n, st); if (need > tmp_bytes) tmp_bytes = need;
  cub::DeviceRadixSort::SortPairs(nullptr, need, keys_b, keys_a, rid_b, rid_c, (int)n, 0, out.view.key_bits, st);
  if (need > tmp_bytes) tmp_bytes = need;
  cub::DeviceScan::InclusiveSum(nullptr, need, head32, kidx1, (int)n, st);
  if (need > tmp_bytes) tmp_bytes = need;
  {
    cub::CountingInputIterator<uint32_t> cnt(0);
    cub::DeviceSelect::Flagged(nullptr, need, cnt, flag, rid_a, d_count, (int)n, st);
    if (need > tmp_bytes) tmp_bytes = need;
  }
  void *tmp = c.pool.device(nm(".cubtmp").c_str(), tmp_bytes);
  need = tmp_bytes;
  cub::DeviceSelect::Flagged(tmp, need, keys_a, flag, keys_b, d_count, (int)n, st);
  need = tmp_bytes;
  cub::DeviceSelect::Flagged(tmp, need, rid_a, flag, rid_b, d_count + 1, (int)n, st);
  c.launches += 4;
  SB_CUDA(cudaMemcpyAsync(h_count, d_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  uint32_t nv = h_count[0];
  out.dict_numreads = nv;
  // stable LSD radix sort: read ids stay ascending inside equal keys
  if (nv) {
    need = tmp_bytes;
    cub::DeviceRadixSort::SortPairs(tmp, need, keys_b, keys_a, rid_b, rid_c, (int)nv, 0, out.view.key_bits, st);
    c.launches += 2 + (out.view.key_bits + 7) / 8 * 2;
    k_mark_heads<<<grid_for(nv, 256), 256, 0, st>>>(keys_a, nv, flag, head32);
    need = tmp_bytes;
    cub::DeviceScan::InclusiveSum(tmp, need, head32, kidx1, (int)nv, st);
    c.launches += 2;
    cub::CountingInputIterator<uint32_t> cnt(0);
    need = tmp_bytes;
    cub::DeviceSelect::Flagged(tmp, need, cnt, flag, rid_a, d_count + 2, (int)nv, st);
    c.launches += 3;
    SB_CUDA(cudaMemcpyAsync(h_count + 2, d_count + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    out.numkeys = h_count[2];
  }
  // slot indices are 32 bits: at most 2^30 unique keys per dictionary (capacity 2^31)
  if (out.numkeys > (1u << 30)) throw LimitError("dictionary: more than 2^30 unique keys in one GPU shard");
  uint32_t cap = 16;
  while (cap < 2ull * out.numkeys) cap <<= 1;
  out.capacity = cap;
  DictSlot *slots = c.pool.dev<DictSlot>(nm(".slots").c_str(), cap);
  SB_CUDA(cudaMemsetAsync(slots, 0, sizeof(DictSlot) * (size_t)cap, st));
  uint64_t fbits = 65536;
  static const unsigned long long kFilterBitsPerKey = getenv("SPRING_B200_FILTER_BITS") ? strtoull(getenv("SPRING_B200_FILTER_BITS"), nullptr, 10) : 8ull;
  while (fbits < kFilterBitsPerKey * out.numkeys && fbits < (1ull << 36)) fbits <<= 1;
  uint32_t *filter = c.pool.dev<uint32_t>(nm(".filter").c_str(), fbits / 32);
  SB_CUDA(cudaMemsetAsync(filter, 0, fbits / 8, st));
  out.view.filter = filter;
  out.view.filter_mask = (uint32_t)(fbits / 32 - 1);
  SB_CUDA(cudaMemsetAsync(slot_of_read, 0xFF, sizeof(uint32_t) * (size_t)n, st));
  if (out.numkeys) {
    k_insert_slots<<<grid_for(out.numkeys, 256), 256, 0, st>>>(keys_a, rid_a, out.numkeys, nv, slots, cap - 1, bins, slot_of_bin);
    k_fill_bins<<<grid_for(nv, 256), 256, 0, st>>>(rid_c, kidx1, rid_a, slot_of_bin, out.numkeys, nv, bins, slot_of_read, slots);
    k_set_filter<<<grid_for(out.numkeys, 256), 256, 0, st>>>(keys_a, rid_a, out.numkeys, filter, out.view.filter_mask);
    c.launches += 3;
  }
  out.view.slots = slots;
  out.view.slot_mask = cap - 1;
  out.view.bins = bins;
  out.view.slot_of_read = slot_of_read;
  out.sorted_keys = keys_a;
  out.bin_start_idx = rid_a;
  out.sorted_rids = rid_c;
  SB_CUDA(cudaGetLastError());
}

}  // namespace sb
