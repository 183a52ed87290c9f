// chain_common.cuh -- definitions of the chain kernel (reorder.cu) that its helpers share.
#pragma once
#include "kernels.cuh"

namespace sb {
namespace chain {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr const char *kChainCfgDefault = "8x4";  // 32 chains per SM (64 registers, a few spills) beat 24 spill-free ones
// probes per lane in the first batch of a free-running search (fast tail): batches cover 8 k, 16 k, then all remaining shifts
constexpr int kBatch0 = 1;
enum { ST_SEARCH = 0, ST_NEWREAD = 1, ST_DONE = 2 };
enum { CTR_UNMATCHED = 0, CTR_ROUNDS, CTR_LOST, CTR_PROBES_ISSUED, CTR_PROBES_SEQ, CTR_COMPARES, CTR_ABORT,
       CTR_CYC_SEARCH, CTR_CYC_WAIT_A, CTR_CYC_COMMIT, CTR_CYC_WAIT_B, CTR_SLOT_PROBES, CTR_N };

struct ChainArgs {
  const uint64_t *reads; const uint16_t *lens; uint32_t N; int L, W, Lp, maxshift;
  DictView dict[2];
  uint32_t *claimed;   // bitmap, one bit per read
  uint32_t *winner;    // [N], kNoWinner until proposed
  uint4 *rec;          // [N] {pos lo, pos hi, index in the chain's log, chain id | meta << 24}: one store per claimed read
  uint32_t *chain_aligned; uint32_t *chain_single;
  uint32_t num_chains, per;
  unsigned long long *barrier; int *active; unsigned long long *ctr;
  unsigned long long max_rounds;
  uint32_t G;            // scan_bin: candidates verified per pass = 32 / W
  unsigned leader_mask;  // scan_bin: lanes g * W, g < G
  int generic_update;    // debugging aid: always use the per-column update_ref
  int prefetch_slots;    // pass 1 of a batch prefetches the slot of every filter positive into L2
  uint64_t pol_keep, pol_stream;  // L2 cache policies (evict_last for the filter words, evict_first for slot sectors)
  int fast_tail;         // free-running chains: batches of 8, 16, then all remaining shifts
  int batch0;            // ... times this many (tunable instantiation; kBatch0 in production)
  int filter_hint;       // filter words are loaded with the L2 evict_last policy (they should outlive the slot sectors)
  int steal_probes;      // free-running schedule: random slices an idle chain probes for an unclaimed read (0 = off)
  unsigned long long *chain_dbg;  // [2 * chains]: steps, globaltimer ns at finish (profiling aid)
};

__device__ __forceinline__ bool is_claimed(const uint32_t *claimed, uint32_t rid) {
  return (__ldcg(claimed + (rid >> 5)) >> (rid & 31)) & 1u;
}

// cnt[col] packs the four per-base counts of a column as u16 fields, rows A,C,T,G (reorder.h:120-123);
// 2-bit read codes are A0 G1 C2 T3 -> field shift 0, 48, 16, 32.  The reference counts in int
// (reorder.h:383-384); here a count SATURATES at 65535 instead of failing: exact up to 65535 reads of one
// base stacked on one column, a tie broken by the vote's fixed order beyond (the output stays decodable
// either way).
__device__ __forceinline__ uint64_t count_add(uint64_t v, int b) {
  const int sh = (int)((0x20103000u >> (8 * b)) & 0xFFu);
  if (((v >> sh) & 0xFFFFull) != 0xFFFFull) v += 1ull << sh;
  return v;
}

// word w of the reverse complement of the len-base sequence a[] (W words, zero beyond len): reverse the
// 2-bit groups of the whole array, complement, shift the padding out (reorder.h:215-217 by bit tricks)
__device__ __forceinline__ uint64_t revcomp_word(const uint64_t *a, int W, int len, int w) {
  const int pad = 64 * W - 2 * len, ws = pad >> 6, bs = pad & 63;
  uint64_t a0 = 0, a1 = 0;
  if (w + ws < W) {
    const uint64_t x = __brevll(a[W - 1 - (w + ws)]);
    a0 = ~(((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull));
  }
  if (w + ws + 1 < W) {
    const uint64_t x = __brevll(a[W - 2 - (w + ws)]);
    a1 = ~(((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull));
  }
  return bs ? (a0 >> bs) | (a1 << (64 - bs)) : a0;
}

// bits [pos, pos + nbits) of a bitset that is followed by one zero word (ref / revref in shared memory):
// no bounds checks, pos < 64 W
__device__ __forceinline__ uint64_t window_key(const uint64_t *a, int pos, int nbits) {
  const int k = pos >> 6, bs = pos & 63;
  const uint64_t v = (a[k] >> bs) | ((a[k + 1] << 1) << (63 - bs));
  return nbits < 64 ? v & ((1ull << nbits) - 1ull) : v;
}


}  // namespace chain
}  // namespace sb
