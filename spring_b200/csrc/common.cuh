// common.cuh -- shared definitions of the spring_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

namespace sb {

// ---- constants of the path (reference: src/params.h:22-37) ---------------------------------
constexpr int kMaxReadLen = 511;       // MAX_READ_LEN
constexpr int kMaxWords = 16;          // ceil(2*511/64)
constexpr int kNumDict = 2;            // NUM_DICT_REORDER / NUM_DICT_ENCODER
constexpr int kMaxSearch = 1000;       // MAX_SEARCH_REORDER / MAX_SEARCH_ENCODER
constexpr int kThreshReorder = 4;      // THRESH_REORDER
constexpr int kThreshEncoder = 24;     // THRESH_ENCODER
constexpr uint32_t kStopWindow = 1000000u;  // reorder.h:433
constexpr uint32_t kStopUnmatched = 500000u;  // STOP_CRITERIA_REORDER * 1e6
constexpr uint32_t kNoWinner = 0xFFFFFFFFu;

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string &m) : std::runtime_error(m) {}
};
struct LimitError : std::runtime_error {
  explicit LimitError(const std::string &m) : std::runtime_error(m) {}
};

#define SB_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      throw sb::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                          std::to_string(__LINE__) + ")");                                   \
  } while (0)

// ---- dictionary in HBM ------------------------------------------------------------------------
// One open-addressing slot per unique key, 32 B = one sector per probe.  Replaces BooPHF + startpos[] + empty_bin[]
// (bitset_util.h:22-41): the (hashed) key is stored, so the reference's "verify against the first read of the bin"
// step (reorder.h:282-285) is folded into the probe, and `live` -- decremented when a read of the bin is claimed --
// plays the role of bbhashdict::remove / empty_bin (bitset_util.cpp:37-63): a bin whose reads are all gone is skipped
// by the probe itself, without touching the bin.  The bin's size and its three highest read ids ride along, so a hit
// on a bin of <= 3 reads (almost all of them) needs no access to bins[] at all.
//
// Placement is ORDERED linear probing: hk = mix64(key) is a bijection of the 64-bit key, the home slot is its top
// bits, and the keys are inserted in ascending hk (the order the build's radix sort leaves them in), each at the first
// free slot at or after its home.  So (a) the build writes the table front to back -- coalesced stores, no atomics --
// and (b) every key between home(x) and x's slot is smaller than x: a lookup stops at the first larger key or empty
// slot.  The table has kSlotPad spare slots behind the last home instead of wrapping around.
struct __align__(32) DictSlot {
  uint64_t key;     // hk = mix64(window key)
  uint32_t start1;  // 1 + index of the bin header in bins[]; 0 = empty slot
  uint32_t live;    // reads of the bin not yet claimed
  uint32_t count;   // reads in the bin
  uint32_t rid[3];  // the bin's highest ids, descending (bins[start1 .. start1+2])
};
constexpr uint32_t kSlotPad = 4096;

// bins[]: per unique key {count, read ids in DESCENDING order}: the reference scans a bin from its
// highest id down (reorder.h:287-288), so a scan is a forward walk from the header.
struct DictView {
  DictSlot *slots;               // [capacity + kSlotPad]
  int slot_shift;                // home slot of hk = hk >> slot_shift (capacity = 2^(64 - slot_shift))
  const uint32_t *bins;
  const uint32_t *slot_of_read;  // [n] slot index holding read i's key, 0xFFFFFFFF if the read is not indexed
  uint32_t *skip;                // [like bins] at a bin's header index: entries before this offset are all claimed
                                 // (monotone hint: a scan of a big bin starts there; plays bbhashdict::remove's compaction)
  const uint32_t *filter;        // blocked Bloom filter over the keys: 2 bits in one 32-bit word, >= 8 bits per key
  uint32_t filter_words;         // filter word of hk = (top 32 bits of hk) * filter_words >> 32: any size, monotone in hk
                                 // (the build sets it front to back too)
  int start, end;                // base window [start, end]
  int key_bits;                  // bits per base * (end - start + 1)
};

// multiply-fold hash of a window key (internal: slot placement and filter bits only)
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 31;
  x *= 0x9E3779B97F4A7C15ULL;
  x ^= x >> 29;
  return x;
}

// key filter (kept L2-resident by the chain kernel as far as it fits): word = top bits of hk scaled to the filter's
// size (no power-of-two rounding: at 100 M keys that rounding alone was 128 MB instead of 100), bits hk[0:5) and hk[5:10)
__host__ __device__ inline uint32_t filter_word(uint64_t hk, uint32_t nwords) { return (uint32_t)(((hk >> 32) * (uint64_t)nwords) >> 32); }
__host__ __device__ inline uint32_t filter_bits(uint64_t hk) { return (1u << (hk & 31)) | (1u << ((hk >> 5) & 31)); }
__host__ __device__ inline uint32_t slot_home(uint64_t hk, int sshift) { return (uint32_t)(hk >> sshift); }
constexpr uint64_t kInvalidKey = ~0ull;  // hk of a read that is not indexed (too short for the window, N inside it)

__host__ __device__ inline int words_for(int max_readlen) { return (2 * max_readlen - 1) / 64 + 1; }

// ---- multi-GPU owner of a read: strand-canonical 16-mer minimizer bucket (bucket.cu, exchange.cu) -----------------
// One pass over the read: forward and reverse-complement 16-mer in two 32-bit registers, per position a 32-bit
// multiply-xorshift hash of the smaller one (a handful of integer instructions: the pass is ALU work, 150 positions per
// read), minimum over the positions; the minimum is spread by mix64 once per read.  A read, its reverse complement and
// its shifted neighbours mostly agree on the minimizer.
constexpr int kMinimizerK = 16;
__host__ __device__ inline uint32_t kmer_hash32(uint32_t x) {
  x *= 0x9E3779B1u;
  return x ^ (x >> 15);
}
__host__ __device__ inline uint32_t minimizer_bucket(const uint64_t *r, int len, uint32_t num_buckets) {
  uint32_t best = 0xFFFFFFFFu, fwd = 0, rc = 0;
  uint64_t w = 0;
  for (int j = 0; j < len; j++) {
    if ((j & 31) == 0) w = r[j >> 5];
    const uint32_t c = (uint32_t)(w & 3ull);
    w >>= 2;
    fwd = (fwd << 2) | c;                      // 16 bases fill the 32-bit word exactly
    rc = (rc >> 2) | ((3u - c) << (2 * (kMinimizerK - 1)));
    if (j >= kMinimizerK - 1) {
      const uint32_t h = kmer_hash32(fwd < rc ? fwd : rc);
      best = h < best ? h : best;
    }
  }
  if (len < kMinimizerK) best = kmer_hash32((uint32_t)len);
  return (uint32_t)((mix64((uint64_t)best) >> 16) % num_buckets);
}

// reorder dictionary windows (reorder.h:752-759)
inline void reorder_windows(int L, int start[2], int end[2]) {
  start[0] = L > 100 ? L / 2 - 32 : L / 2 - L * 32 / 100;
  end[0] = L / 2 - 1;
  start[1] = L / 2;
  end[1] = L > 100 ? L / 2 - 1 + 32 : L / 2 - 1 + L * 32 / 100;
}
// encoder dictionary windows (encoder.h:609-620)
inline void encoder_windows(int L, int start[2], int end[2]) {
  if (L > 50) { start[0] = 0; end[0] = 20; start[1] = 21; end[1] = 41; }
  else { start[0] = 0; end[0] = 20 * L / 50; start[1] = 20 * L / 50 + 1; end[1] = 41 * L / 50; }
}

#ifdef __CUDACC__
// word i of (a >> bits) / (a << bits) for a W-word little-endian bitset (std::bitset semantics)
__device__ __forceinline__ uint64_t shr_word(const uint64_t *a, int W, int i, int bits) {
  int k = i + (bits >> 6), bs = bits & 63;
  uint64_t lo = k < W ? a[k] : 0ull, hi = k + 1 < W ? a[k + 1] : 0ull;
  return bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
}
__device__ __forceinline__ uint64_t shl_word(const uint64_t *a, int W, int i, int bits) {
  int k = i - (bits >> 6), bs = bits & 63;
  uint64_t hi = (k >= 0 && k < W) ? a[k] : 0ull, lo = (k - 1 >= 0 && k - 1 < W) ? a[k - 1] : 0ull;
  return bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
}
// bits [pos, pos+nbits) of a, nbits <= 64
__device__ __forceinline__ uint64_t extract_bits(const uint64_t *a, int W, int pos, int nbits) {
  uint64_t v = shr_word(a, W, 0, pos);
  return nbits < 64 ? v & ((1ull << nbits) - 1ull) : v;
}
// mask of word i covering bits [lo, hi)
__device__ __forceinline__ uint64_t range_mask(int i, int lo, int hi) {
  int b0 = i << 6;
  int l = lo - b0, h = hi - b0;
  if (h <= 0 || l >= 64 || hi <= lo) return 0ull;
  uint64_t m = ~0ull;
  if (l > 0) m &= ~0ull << l;
  if (h < 64) m &= ~0ull >> (64 - h);
  return m;
}
// spread the 32 bits of x to the even bit positions of a 64-bit word
__device__ __forceinline__ uint64_t spread_bits(uint32_t x) {
  uint64_t v = x;
  v = (v | (v << 16)) & 0x0000FFFF0000FFFFull;
  v = (v | (v << 8)) & 0x00FF00FF00FF00FFull;
  v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0Full;
  v = (v | (v << 2)) & 0x3333333333333333ull;
  v = (v | (v << 1)) & 0x5555555555555555ull;
  return v;
}
__device__ __forceinline__ int base_code(const uint64_t *w, int j) { return (int)((w[j >> 5] >> (2 * (j & 31))) & 3ull); }

__device__ __forceinline__ bool filter_test(const uint32_t *filter, uint32_t nwords, uint64_t hk) {
  const uint32_t b = filter_bits(hk);
  return (__ldg(filter + filter_word(hk, nwords)) & b) == b;
}
// L2 eviction policies: the key filter should stay in L2, the slot table streams through it
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ bool filter_test_hint(const uint32_t *filter, uint32_t nwords, uint64_t hk, uint64_t pol) {
  const uint32_t b = filter_bits(hk);
  uint32_t w;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(w) : "l"(filter + filter_word(hk, nwords)), "l"(pol));
  return (w & b) == b;
}
__device__ __forceinline__ uint32_t filter_load_hint(const uint32_t *filter, uint32_t nwords, uint64_t hk, uint64_t pol) {
  uint32_t w;
  asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(w) : "l"(filter + filter_word(hk, nwords)), "l"(pol));
  return w;
}
// L2 (.cg) load: `live` is updated by other SMs between rounds, L1 must not serve it
__device__ __forceinline__ DictSlot load_slot(const DictSlot *p) {
  const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
  const uint4 w = __ldcg(reinterpret_cast<const uint4 *>(p) + 1);
  DictSlot s;
  s.key = (uint64_t)v.x | ((uint64_t)v.y << 32);
  s.start1 = v.z;
  s.live = v.w;
  s.count = w.x; s.rid[0] = w.y; s.rid[1] = w.z; s.rid[2] = w.w;
  return s;
}
__device__ __forceinline__ DictSlot load_slot_hint(const DictSlot *p, uint64_t pol) {
  uint4 v, w;
  asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
  asm volatile("ld.global.cg.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "l"(reinterpret_cast<const uint4 *>(p) + 1), "l"(pol));
  DictSlot s;
  s.key = (uint64_t)v.x | ((uint64_t)v.y << 32);
  s.start1 = v.z;
  s.live = v.w;
  s.count = w.x; s.rid[0] = w.y; s.rid[1] = w.z; s.rid[2] = w.w;
  return s;
}
__device__ __forceinline__ DictSlot load_slot_head(const DictSlot *p) {  // key / start1 / live only
  const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p));
  DictSlot s;
  s.key = (uint64_t)v.x | ((uint64_t)v.y << 32);
  s.start1 = v.z;
  s.live = v.w;
  s.count = 0; s.rid[0] = s.rid[1] = s.rid[2] = 0;
  return s;
}
// exact lookup in a dictionary nobody is updating: header index of the bin, or -1 (hk = mix64(key))
__device__ __forceinline__ long long dict_find(const DictView &d, uint64_t hk) {
  uint32_t h = slot_home(hk, d.slot_shift);
  for (;;) {
    DictSlot s = load_slot_head(d.slots + h);
    if (s.start1 == 0 || s.key > hk) return -1;
    if (s.key == hk) return (long long)s.start1 - 1;
    h++;
  }
}
#endif

// ---- grow-only named device / pinned buffers owned by a context --------------------------------
class BufferPool {
 public:
  ~BufferPool() { release(); }
  void *device(const char *name, size_t bytes) {
    Buf &b = dev_[name];
    if (bytes > b.cap) {
      if (b.p) SB_CUDA(cudaFree(b.p));
      b.p = nullptr; b.cap = 0;
      size_t want = bytes + bytes / 8 + 256;
      SB_CUDA(cudaMalloc(&b.p, want));
      b.cap = want;
    }
    return b.p;
  }
  void *pinned(const char *name, size_t bytes) {
    Buf &b = pin_[name];
    if (bytes > b.cap) {
      if (b.p) SB_CUDA(cudaFreeHost(b.p));
      b.p = nullptr; b.cap = 0;
      size_t want = bytes + bytes / 8 + 256;
      SB_CUDA(cudaMallocHost(&b.p, want));
      b.cap = want;
    }
    return b.p;
  }
  template <typename T> T *dev(const char *name, size_t n) { return static_cast<T *>(device(name, n * sizeof(T))); }
  template <typename T> T *pin(const char *name, size_t n) { return static_cast<T *>(pinned(name, n * sizeof(T))); }
  void release() {
    for (auto &kv : dev_) if (kv.second.p) cudaFree(kv.second.p);
    for (auto &kv : pin_) if (kv.second.p) cudaFreeHost(kv.second.p);
    dev_.clear(); pin_.clear();
  }
 private:
  struct Buf { void *p = nullptr; size_t cap = 0; };
  std::map<std::string, Buf> dev_, pin_;
};

}  // namespace sb
