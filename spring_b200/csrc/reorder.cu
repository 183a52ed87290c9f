// reorder.cu -- the greedy overlap search (reference src/reorder.h:320-641) as one persistent sm_100a
// kernel.
//
// Mapping.  One warp = one "chain" = one reference OpenMP thread (reorder.h:351): it owns a
// consensus window (per-column base counts in shared memory, ref / revref bitsets), and repeats
// { search the dictionaries for an overlapping read, claim it, fold it into the consensus }.
//   * search: the reference tries shift 0,1,2,... one after the other, 4 dictionary probes per shift
//     (search_match, reorder.h:246-318).  Here lane l owns probe kind l & 3 (strand x dictionary) and the
//     shifts S + (l >> 2) + 8j of a batch: every probe tests one bit of an L2-resident key filter, the
//     positives load their 32-byte slot from HBM, and the warp walks the hits in the reference's priority
//     order (shift, forward before reverse, dict 0 before dict 1, highest read id first), so the read it
//     picks is the one the sequential search would pick.
//   * verification (scan_bin): the warp is cut into 32 / W groups of W lanes, one candidate per group, one
//     bitset word per lane: XOR with the shifted reference, mask the overlap, popcount, W-lane sum
//     (THRESH_REORDER = 4).
//   * updaterefcount (reorder.h:110-220): the four count-shift cases collapse to one remap
//     "new column i <- old column i + delta", 32 columns per step; the consensus is rebuilt by word
//     operations and re-voted only where read and old consensus differ (update_ref_fast).
//
// Scheduling.  The reference is racy (try-locks) and non-reproducible for more than one thread.
//   * free-running (default): chains claim reads with an atomic test-and-set like the reference's threads;
//     a chain that has drained its slice of the read ids seeds further contigs from randomly probed other
//     slices.  Output depends on timing for more than one chain, as the reference's does for -t > 1.
//   * deterministic (cooperative launch): chains run in lock step: phase A every chain searches against
//     the claim bitmap as of the round start and proposes one read (atomicMin of its chain id into
//     winner[rid]); grid barrier; phase B the winner claims and updates, losers retry; grid barrier.  The
//     result is a pure function of (input, num_chains) and is restated exactly by oracle/spring_oracle.c,
//     which for one chain is the reference's single-thread execution.
//
// Two instantiations of the free-running kernel: the production one has the lookup / compare counters and the tuning
// knobs compiled out (no spills at 64 registers); the counting one (spring_b200_set_chain_stats, and always the
// deterministic schedule) counts what the oracle counts.  What the kernel is sensitive to, in this order (measured,
// profiles/r02_chains2_experiment.txt): registers and code size, serialised dependent loads, instruction count --
// not DRAM traffic.  Hence: once-per-contig state in shared memory, the probe loop not unrolled, filter words of
// multi-probe batches requested four at a time, the count pass as straight-line code.
//
// Memory traffic per claimed read (L = 150): ~39 filter words (L2 while the filters fit it), ~3 slot sectors, one or
// two 40-byte candidate rows + their claim words, one 16-byte record; all other state stays in shared memory /
// registers for the life of the kernel (DESIGN.md sections 5, 6).
#include <algorithm>
#include <cub/cub.cuh>
#include "chain_common.cuh"

#ifndef SB_PROBE_GROUP
#define SB_PROBE_GROUP 4   // batches with more than one probe per lane: filter words requested in groups of this many before any is examined
#endif
#ifndef SB_PROBE_UNROLL
#define SB_PROBE_UNROLL 1  // the probe loop is not unrolled: most batches have one probe per lane, and the kernel's code size costs instruction-cache
#endif                     // misses (visit 24, k_chains ms on configs 2 / 3 / 5: unroll 4 15.6 / 191.0 / 227.5, unroll 1 15.5 / 190.0 / 212.6)
#ifndef SB_OPAQUE_TID
#define SB_OPAQUE_TID 1    // %tid is read once through volatile asm: lane / warp index stay in registers instead of being re-derived
#endif
#ifndef SB_COUNT_UNROLL
#define SB_COUNT_UNROLL 1  // count pass of update_ref_fast: straight-line code for the WT chunks (1) or the rolled loop (0)
#endif
#ifndef SB_NOINLINE_FIND
#define SB_NOINLINE_FIND 0 // find_unclaimed as a real function call (cold path; smaller kernel)
#endif
#define SB_STR2(x) #x
#define SB_STR(x) SB_STR2(x)
#define SB_UNROLL(n) _Pragma(SB_STR(unroll n))

namespace sb {

using namespace chain;

namespace {

// work/wait: SM cycles thread 0 of the block spent between barriers / spinning in this one
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long &target, long long &t_last,
                                             unsigned long long &work, unsigned long long &wait) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t_arr = clock64();
    work += t_arr - t_last;
    target += gridDim.x;
    unsigned long long v;
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(ctr) : "memory");
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory");
    } while (v < target);
    t_last = clock64();
    wait += t_last - t_arr;
  }
  __syncthreads();
}

// Fold the read staged in curw (cur_len bases; rev: use its reverse complement) into the window.
// new column i takes old column i+delta when that lies in [0, old_len), else starts from zero;
// the read covers new columns [cs, cs+cur_len).  Then majority -> ref, RC(ref) -> revref.
//
// fold > 0 reproduces a quirk of the reference that the output depends on: in the reverse case
// where the new read covers the whole window (reorder.h:159-165) the counts are moved UP by
// fold = cur_len - shift - old_len columns with an ascending in-place loop, so a source column
// that was already rewritten is read again: column i = q*fold + r ends up as
// old[r] + sum_{t=1..q} e(read base at t*fold + r) rather than old[i - fold] + e(read base at i).
__device__ __noinline__ void update_ref(uint64_t *ref, uint64_t *revref, const uint64_t *curw, uint64_t *cnt, int W, int lane,
                           int old_len, int delta, int cs, int cur_len, bool rev, int new_len, int fold) {
  const int nchunks = (new_len + 31) >> 5;
  for (int cc = 0; cc < nchunks; cc++) {
    const int ck = delta >= 0 ? cc : nchunks - 1 - cc;  // move direction decides the safe order
    const int i = (ck << 5) + lane;
    const bool in = i < new_len;
    const int src = i + delta;
    uint64_t v = 0;
    if (in && src >= 0 && src < old_len) {
      if (fold > 0) {
        const int r = i % fold, q = i / fold;
        v = cnt[r];
        for (int t = 1; t < q; t++) {  // the t == q term is the read's own base at column i, added below
          v = count_add(v, 3 - base_code(curw, cur_len - 1 - (t * fold + r)));
        }
      } else {
        v = cnt[src];
      }
    }
    __syncwarp();
    uint32_t code = 0;
    if (in) {
      const int ci = i - cs;
      if (ci >= 0 && ci < cur_len) {
        v = count_add(v, rev ? 3 - base_code(curw, cur_len - 1 - ci) : base_code(curw, ci));
      }
      cnt[i] = v;
      const uint32_t f0 = (uint32_t)v & 0xFFFFu, f1 = (uint32_t)(v >> 16) & 0xFFFFu;
      const uint32_t f2 = (uint32_t)(v >> 32) & 0xFFFFu, f3 = (uint32_t)(v >> 48);
      uint32_t mx = f0;  // first strict maximum over rows A,C,T,G (reorder.h:204-212) -> codes 0,2,3,1
      if (f1 > mx) { mx = f1; code = 2; }
      if (f2 > mx) { mx = f2; code = 3; }
      if (f3 > mx) { mx = f3; code = 1; }
    }
    // 32 two-bit codes -> one 64-bit word: lanes 0-15 fill the low half, 16-31 the high half
    const uint32_t part = code << (2 * (lane & 15));
    const uint32_t lo = __reduce_or_sync(FULL, lane < 16 ? part : 0u);
    const uint32_t hi = __reduce_or_sync(FULL, lane < 16 ? 0u : part);
    if (lane == 0) ref[ck] = (uint64_t)lo | ((uint64_t)hi << 32);
    __syncwarp();
  }
  if (lane >= nchunks && lane < W) ref[lane] = 0ull;
  __syncwarp();
  // revref = reverse complement of ref (reorder.h:215-217) by bit tricks: reverse the 2-bit groups of
  // the whole W-word array, complement, shift the padding out
  if (lane < W) {
    const int pad = 64 * W - 2 * new_len, ws = pad >> 6, bs = pad & 63;
    uint64_t a0 = 0, a1 = 0;
    if (lane + ws < W) {
      uint64_t x = __brevll(ref[W - 1 - (lane + ws)]);
      a0 = ~(((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull));
    }
    if (lane + ws + 1 < W) {
      uint64_t x = __brevll(ref[W - 1 - (lane + ws + 1)]);
      a1 = ~(((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull));
    }
    revref[lane] = bs ? (a0 >> bs) | (a1 << (64 - bs)) : a0;
  }
  __syncwarp();
}

// update_ref for every case but the fold quirk (delta >= 0): same result as the generic version above,
// with the consensus rebuilt by word operations instead of a majority vote per column.
//   * counts: one pass over the columns, 32 per step: cnt[i] = cnt[i + delta] (+ the read's base);
//   * consensus: ref[] always equals the column-wise majority of cnt[] (true after a reset, kept by
//     every update).  A column the read does not cover keeps its counts, hence its base; a column where
//     the read agrees with the old consensus keeps it too (the winner's count grows); a column with no
//     source is new and takes the read's base.  So new ref = (old ref >> delta columns) merged with the
//     read's words, and only the columns where read and old consensus DIFFER -- at most THRESH_REORDER
//     of them, the match passed the Hamming test on exactly these bits -- need the vote.
// curw is overwritten with the read as oriented in the contig.
// WT > 0: words per read known at compile time (the chunk loop unrolls into straight-line, branch-free code: the count
// pass was the largest single block of the kernel, 47 instructions per 32 columns with its bounds branches; now ~20).
// saturating: some column of this contig may have reached 65535 (more than 65534 reads folded in since the counts were
// reset) -- only then is the per-field saturation check of count_add needed.
template <int WT>
__device__ __forceinline__ void update_ref_fast(uint64_t *ref, uint64_t *revref, uint64_t *curw, uint64_t *cnt, int Wrt, int lane,
                                                int old_len, int delta, int cs, int cur_len, bool rev, int new_len, bool saturating) {
  const int W = WT ? WT : Wrt;
  if (rev) {
    uint64_t o = 0;
    if (lane < W) o = revcomp_word(curw, W, cur_len, lane);
    __syncwarp();
    if (lane < W) curw[lane] = o;
    __syncwarp();
  }
  // the read in window coordinates (shifted up by cs columns, nothing outside its columns), and what the consensus merge
  // below needs from the old consensus -- both before curw / ref are overwritten
  uint64_t A = 0, B = 0, MA = 0, MB = 0;
  if (lane < W) {
    MA = range_mask(lane, 0, 2 * (old_len - delta));        // columns that have a source column
    MB = range_mask(lane, 2 * cs, 2 * (cs + cur_len));      // columns the read covers
    A = shr_word(ref, W, lane, 2 * delta) & MA;
    B = shl_word(curw, W, lane, 2 * cs) & MB;
  }
  __syncwarp();
  if (lane < W) curw[lane] = B;
  __syncwarp();
  // counts: cnt[i] = cnt[i + delta] (+ the read's base), 32 columns per chunk, column i = 32 cc + lane.  All sources are
  // loaded before any column is stored (a chunk's sources overlap the columns the same chunk writes when delta < 32).
  // cnt[col] packs the four per-base counts as u16 fields, rows A,C,T,G (reorder.h:120-123); 2-bit read codes are
  // A0 G1 C2 T3 -> field 0, 3, 1, 2 = (0x9C >> 2 code) & 3.
  const int sh2 = 2 * (lane & 15);
  const bool hi_half = (lane & 16) != 0;
  if (WT && SB_COUNT_UNROLL) {
    uint64_t v[WT ? WT : 1];
#pragma unroll
    for (int cc = 0; cc < WT; cc++) {
      const int src = (cc << 5) + lane + delta;
      v[cc] = src < old_len ? cnt[src] : 0ull;
    }
    __syncwarp();
#pragma unroll
    for (int cc = 0; cc < WT; cc++) {
      const int i = (cc << 5) + lane;
      const uint2 bw = reinterpret_cast<const uint2 *>(curw)[cc];                 // same word for the whole warp: a broadcast
      const uint32_t code = ((hi_half ? bw.y : bw.x) >> sh2) & 3u;
      const bool covered = (unsigned)(i - cs) < (unsigned)cur_len;
      const int fsh = (int)((0x9Cu >> (2 * code)) & 3u) << 4;
      if (!saturating) v[cc] += (uint64_t)(covered ? 1u : 0u) << fsh;
      else if (covered && ((v[cc] >> fsh) & 0xFFFFull) != 0xFFFFull) v[cc] += 1ull << fsh;  // counts saturate at 65535 (count_add)
      if (i < new_len) cnt[i] = v[cc];
    }
  } else {
    const int nchunks = (new_len + 31) >> 5;
    for (int cc = 0; cc < nchunks; cc++) {  // ascending is safe: delta >= 0, sources lie at or above the column
      const int i = (cc << 5) + lane;
      uint64_t v = i + delta < old_len ? cnt[i + delta] : 0ull;
      __syncwarp();
      const uint2 bw = reinterpret_cast<const uint2 *>(curw)[cc];
      const uint32_t code = ((hi_half ? bw.y : bw.x) >> sh2) & 3u;
      const bool covered = (unsigned)(i - cs) < (unsigned)cur_len;
      const int fsh = (int)((0x9Cu >> (2 * code)) & 3u) << 4;
      if (covered && ((v >> fsh) & 0xFFFFull) != 0xFFFFull) v += 1ull << fsh;
      if (i < new_len) cnt[i] = v;
    }
  }
  uint64_t nw = 0, mm = 0;
  if (lane < W) {
    nw = A | (B & ~MA);
    const uint64_t X = (A ^ B) & MA & MB;
    mm = (X | (X >> 1)) & 0x5555555555555555ull;  // bit 2t: column 32*lane + t needs the vote
  }
  __syncwarp();  // counts written, old ref read
  while (mm) {
    const int bp = __ffsll((long long)mm) - 1;
    mm &= mm - 1;
    const uint64_t v = cnt[(lane << 5) + (bp >> 1)];
    const uint32_t f0 = (uint32_t)v & 0xFFFFu, f1 = (uint32_t)(v >> 16) & 0xFFFFu;
    const uint32_t f2 = (uint32_t)(v >> 32) & 0xFFFFu, f3 = (uint32_t)(v >> 48);
    uint32_t mx = f0, code = 0;  // first strict maximum over rows A,C,T,G (reorder.h:204-212) -> codes 0,2,3,1
    if (f1 > mx) { mx = f1; code = 2; }
    if (f2 > mx) { mx = f2; code = 3; }
    if (f3 > mx) { mx = f3; code = 1; }
    nw = (nw & ~(3ull << bp)) | ((uint64_t)code << bp);
  }
  if (lane < W) ref[lane] = nw;
  __syncwarp();
  if (lane < W) revref[lane] = revcomp_word(ref, W, new_len, lane);
  __syncwarp();
}

// Verify the live reads of one bin, highest id first, at most MAX_SEARCH of them (reorder.h:287-311).
// r0..r2: the bin's first three entries (from the slot); bins[] is only read for bins of > 3 reads.
//
// Lane layout: the warp is cut into G = 32 / W groups of W lanes; group g verifies candidate off + g and
// lane wig of the group handles word wig of the bitsets: one coalesced 8W-byte row load per candidate, one
// XOR / mask / popcount per lane, a W-lane segmented sum.  (A lane per candidate would spend W times the
// instructions on the one or two candidates a typical bin holds.)  The shifted reference word is the
// same for every candidate of the scan and is computed once.
template <int WT, bool STATS>
__device__ __forceinline__ bool scan_bin(const ChainArgs &a, const DictView &d, uint32_t bs, uint32_t bc, uint32_t r0, uint32_t r1, uint32_t r2,
                         const uint64_t *refsm, bool rev, int s, int ref_len, int lane, int grp, int wig, uint32_t &rid_out,
                         uint32_t &compares) {
  const int W = WT ? WT : a.W;  // WT > 0: words per read known at compile time (index arithmetic folds)
  const uint32_t G = WT ? 32u / (uint32_t)(WT ? WT : 1) : a.G;
  const bool act = (uint32_t)grp < G;
  unsigned leaders = a.leader_mask;                        // lanes with wig == 0 of the G groups
  if (WT) { leaders = 0; for (uint32_t q = 0; q < G; q++) leaders |= 1u << (q * (uint32_t)W); }  // a constant when W is
  const unsigned below = leaders & ((1u << (lane - wig)) - 1u);  // leaders of the groups before mine
  const uint64_t rw = rev ? shl_word(refsm, W, wig, 2 * s) : shr_word(refsm, W, wig, 2 * s);
  int live_before = 0;
  // big bins (repeats): skip the prefix of entries already known to be claimed, and extend that
  // hint when this scan meets more of them -- claims only grow, so the hint never hides a live read
  uint32_t t0 = 0, dead_to = 0;
  bool prefix_dead = bc > 3;
  if (bc > 3) { t0 = __ldcg(d.skip + (bs - 1)); dead_to = t0; }
  for (uint32_t off = t0; off < bc; off += G) {
    const uint32_t t = off + (uint32_t)grp;
    uint32_t rid = 0;
    bool live = false;
    uint64_t cw = 0;
    int len = 0;
    if (act && t < bc) {
      rid = bc <= 3 ? (t == 0 ? r0 : t == 1 ? r1 : r2) : __ldg(d.bins + bs + t);
      // row word and length are fetched before the claim bit is known: one memory round trip per
      // pass instead of two (a claimed candidate costs a wasted sector, not a serialised latency)
      cw = __ldg(a.reads + (size_t)rid * W + wig);
      len = __ldg(a.lens + rid);
      live = !is_claimed(a.claimed, rid);
    }
    const unsigned lm = __ballot_sync(FULL, live) & leaders;  // one bit per live candidate, in scan order
    if (prefix_dead) {
      if (lm == 0) dead_to = min(bc, off + G);
      else {
        dead_to = off + (uint32_t)__popc(leaders & ((1u << (__ffs(lm) - 1)) - 1u));
        prefix_dead = false;
        if (lane == 0 && dead_to > t0) atomicMax(d.skip + (bs - 1), dead_to);
      }
    }
    const int rank = live_before + __popc(lm & below);
    const bool ev = live && rank < kMaxSearch;
    int h = 0;
    if (ev) {
      int lo, hi;
      if (!rev) { lo = 0; hi = 2 * min(ref_len - s, len); }
      else { lo = 2 * s; hi = 2 * min(ref_len + s, len); }
      h = __popcll((rw ^ cw) & range_mask(wig, lo, hi));
    }
    for (int o = 1; o < W; o <<= 1) {  // sum over the group's W lanes, into its leader
      const int t2 = __shfl_down_sync(FULL, h, o);
      if (wig + o < W) h += t2;
    }
    const unsigned em = __ballot_sync(FULL, ev) & leaders;
    const unsigned pm = __ballot_sync(FULL, ev && h <= kThreshReorder) & leaders;
    if (pm) {
      const int wl = __ffs(pm) - 1;
      rid_out = __shfl_sync(FULL, rid, wl);
      if (STATS) compares += __popc(em & ((2u << wl) - 1u));
      return true;
    }
    if (STATS) compares += __popc(em);
    live_before += __popc(lm);
    if (live_before >= kMaxSearch) break;
  }
  if (prefix_dead && lane == 0 && dead_to > t0) atomicMax(d.skip + (bs - 1), dead_to);
  return false;
}

// Phase A for a searching chain: the first read, in the reference's order, that matches.
//
// Lane l of the warp owns probe kind (l & 3) = 2*strand + dict and shifts S + (l >> 2) + 8*j:
// a batch covers 8*n consecutive shifts with n probes per lane, n = 1, 2, 4, 8, 16 (most matches sit
// within the first 8 shifts, so most rounds cost 32 probes; a dead end walks all L/2 shifts in
// log-many batches).  Every probe first tests one bit of the dictionary's key filter (32 MB per
// dictionary at 10 M keys: L2-resident, no DRAM access); only filter positives (true keys + ~5 %
// false positives) go on to the slot table in HBM.  The warp then walks its hits in the reference's
// order (shift, forward before reverse, dict 0 before dict 1; bins from the highest id down).
// One call examines ONE batch: shifts [S, S + 8n), n = 1, 2, 4, 8, 16, 16, ... probes per lane for
// batch b = 0, 1, ...; a chain that finds nothing continues with the next batch in the next round
// (bounded work per round keeps the lock-step chains balanced; claims only grow, so earlier batches
// cannot turn productive later -- same result as a full search, see oracle/spring_oracle.c).
template <int WT, bool FAST_TAIL, bool STATS>
__device__ __forceinline__ bool chain_search(const ChainArgs &a, const uint64_t *ref, const uint64_t *revref, int ref_len, int lane, int grp,
                             int wig, int b, int S, uint32_t &prop_rid, int &prop_shift, int &prop_rev, uint32_t &probes_issued,
                             uint32_t &probes_seq, uint32_t &compares, uint32_t &slot_probes) {
  const int W = WT ? WT : a.W;
  const int kind = lane & 3, rev = kind >> 1, sub = lane >> 2;
  const DictView &d = a.dict[kind & 1];
  const uint64_t *src = rev ? revref : ref;
  const uint64_t pol_keep = a.pol_keep, pol_stream = a.pol_stream;  // createpolicy results, made once on the host side's behalf
  // shifts this lane's probe kind may use: forward d.end + s < ref_len (reorder.h:264-265), reverse
  // d.end < ref_len + s and s < d.start (:266-267), all below maxshift
  const int s_lo = rev ? d.end - ref_len + 1 : 0;
  const int s_hi = min(a.maxshift, rev ? d.start : ref_len - d.end);
  // bit position of the window key in src at shift s: kbase + ksign * 2s
  const int kbase = 2 * d.start, kstep = rev ? -2 : 2;
  {
    // probes per lane: the deterministic schedule keeps the oracle's rounds of 8, 16, 32, 64, then 128 shifts; a free-running
    // chain goes 8, 16, then everything that is left (up to 128 shifts): 99 % of the matches sit below shift 24, so the
    // third batch is almost always the last one of a dead end -- three round trips per dead end instead of four or five
    // (SPRING_B200_FAST_TAIL=0 restores the oracle's batches: -6.5 % kernel time on config 5's contig-start-heavy input, neutral
    // on configs 2 and 3; profiles/r02_chains2_experiment.txt)
    const int n = (FAST_TAIL && (STATS ? a.fast_tail : 1)) ? (b < 2 ? (STATS ? a.batch0 : kBatch0) << b : 16) : (b < 4 ? 1 << b : 16);
    // ---- pass 1: bounds + filter bit for this lane's n probes ---------------------------------------------------------
    // The production instantiation (STATS = false) has the tuning knobs compiled in: L2 evict_last on the filter words,
    // slot prefetch for filter positives, fast tail.
    unsigned okm = 0, cand = 0;
    const bool hint = STATS ? a.filter_hint != 0 : true, pf = STATS ? a.prefetch_slots != 0 : true;
    if (SB_PROBE_GROUP <= 1 || n == 1) {
      SB_UNROLL(SB_PROBE_UNROLL)
      for (int j = 0; j < n; j++) {
        const int s = S + sub + 8 * j;
        if (s >= s_lo && s < s_hi) {
          okm |= 1u << j;
          const uint64_t hk = mix64(window_key(src, kbase + kstep * s, d.key_bits));
          if (hint ? filter_test_hint(d.filter, d.filter_words, hk, pol_keep) : filter_test(d.filter, d.filter_words, hk)) {
            cand |= 1u << j;
            if (pf) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.slots + slot_home(hk, d.slot_shift)));
          }
        }
      }
    } else {
      // later batches (2 .. 16 probes per lane; a dead end walks them all): the filter words of SB_PROBE_GROUP probes are
      // requested back to back and examined afterwards -- one memory round trip per group instead of one per probe
      for (int j0 = 0; j0 < n; j0 += SB_PROBE_GROUP) {
        uint32_t fw[SB_PROBE_GROUP], fb[SB_PROBE_GROUP], hm[SB_PROBE_GROUP];
        unsigned okg = 0;
#pragma unroll
        for (int u = 0; u < SB_PROBE_GROUP; u++) {
          const int s = S + sub + 8 * (j0 + u);
          fw[u] = 0; fb[u] = 1; hm[u] = 0;
          if (j0 + u < n && s >= s_lo && s < s_hi) {
            okg |= 1u << u;
            const uint64_t hk = mix64(window_key(src, kbase + kstep * s, d.key_bits));
            fb[u] = filter_bits(hk);
            hm[u] = slot_home(hk, d.slot_shift);
            fw[u] = hint ? filter_load_hint(d.filter, d.filter_words, hk, pol_keep) : __ldg(d.filter + filter_word(hk, d.filter_words));
          }
        }
#pragma unroll
        for (int u = 0; u < SB_PROBE_GROUP; u++) {
          if ((fw[u] & fb[u]) == fb[u]) {  // never true for a probe that was not issued (fw = 0, fb = 1)
            cand |= 1u << (j0 + u);
            if (pf) asm volatile("prefetch.global.L2 [%0];" ::"l"(d.slots + hm[u]));
          }
        }
        okm |= okg << j0;
      }
    }
    if (STATS) probes_issued += (unsigned)__popc(okm);  // per-lane partial sum, reduced when the chain ends
    // ---- pass 2: resolve hits in priority order ----------------------------------------------------
    int cur_j = -1, found_p = -1;
    uint32_t cur_start1 = 0, cur_count = 0, cur_r0 = 0, cur_r1 = 0, cur_r2 = 0;
    for (;;) {
      while (cur_j < 0 && cand) {  // this lane's next filter positive -> slot table (HBM)
        const int j = __ffs(cand) - 1;
        cand &= cand - 1;
        const int s = S + sub + 8 * j;
        const uint64_t hk = mix64(window_key(src, kbase + kstep * s, d.key_bits));
        uint32_t h = slot_home(hk, d.slot_shift);
        if (STATS) slot_probes++;
        for (;;) {  // ordered probing: every key between the home slot and hk's own slot is smaller than hk
          const DictSlot sl = load_slot_hint(d.slots + h, pol_stream);
          if (sl.start1 == 0 || sl.key > hk) break;
          if (sl.key == hk) {
            // a bin with no live read is an "empty_bin" (reorder.h:277-281): skipped without a visit
            if (sl.live) { cur_j = j; cur_start1 = sl.start1; cur_count = sl.count; cur_r0 = sl.rid[0]; cur_r1 = sl.rid[1]; cur_r2 = sl.rid[2]; }
            break;
          }
          h++;
        }
      }
      const int myp = cur_j >= 0 ? (((S + sub + 8 * cur_j) << 2) | kind) : 0x7FFFFFFF;
      const int p = __reduce_min_sync(FULL, myp);
      if (p == 0x7FFFFFFF) break;
      const int ps = p >> 2, pk = p & 3;
      const int owner = (((ps - S) & 7) << 2) | pk;
      const uint32_t mb = __shfl_sync(FULL, cur_start1, owner);  // entries follow the header at bins[mb - 1]
      const uint32_t mc = __shfl_sync(FULL, cur_count, owner);
      const uint32_t r0 = __shfl_sync(FULL, cur_r0, owner), r1 = __shfl_sync(FULL, cur_r1, owner), r2 = __shfl_sync(FULL, cur_r2, owner);
      const DictView &pd = a.dict[pk & 1];
      uint32_t rid;
      if (scan_bin<WT, STATS>(a, pd, mb, mc, r0, r1, r2, (pk >> 1) ? revref : ref, pk >> 1, ps, ref_len, lane, grp, wig, rid, compares)) {
        prop_rid = rid; prop_shift = ps; prop_rev = pk >> 1; found_p = p;
        break;
      }
      if (lane == owner) cur_j = -1;
    }
    // lookups a sequential search would have issued: all of this batch, or those up to the hit
    unsigned seqmask = okm;
    if (STATS && found_p >= 0) {
      const int fs = found_p >> 2, fk = found_p & 3;
      const int rel = fs - S - sub;  // this lane's shifts <= fs are j <= rel / 8
      if (rel < 0) seqmask = 0;
      else {
        int jm = rel >> 3;
        if ((rel & 7) == 0 && kind > fk) jm--;  // same shift, later kind: not reached
        seqmask = jm < 0 ? 0u : (jm >= 31 ? seqmask : seqmask & ((2u << jm) - 1u));
      }
    }
    if (STATS) probes_seq += (unsigned)__popc(seqmask);
    if (found_p >= 0) return true;
  }
  return false;
}

// One 16-byte record per read, written by the chain that claims it (one store): position in the contig, index in the
// chain's aligned / singleton log, chain id + meta (bit 0 reverse, bit 1 flag '1', bit 2 singleton) in the top byte.
__device__ __forceinline__ uint4 make_rec(long long pos, uint32_t k, uint32_t chain, uint32_t meta) {
  return make_uint4((uint32_t)(unsigned long long)pos, (uint32_t)((unsigned long long)pos >> 32), k, chain | (meta << 24));
}

// Highest unclaimed read in [lo, cursor] (reorder.h:576-592), 32 bitmap words per step.
#if SB_NOINLINE_FIND
__device__ __noinline__ bool find_unclaimed(const uint32_t *claimed, long long lo, long long cursor, int lane, uint32_t &rid) {
#else
__device__ bool find_unclaimed(const uint32_t *claimed, long long lo, long long cursor, int lane, uint32_t &rid) {
#endif
  if (cursor < lo) return false;
  const long long whi = cursor >> 5, wlo = lo >> 5;
  for (long long wbase = whi; wbase >= wlo; wbase -= 32) {
    const long long wi = wbase - lane;
    uint32_t fb = 0;
    if (wi >= wlo) {
      fb = ~__ldcg(claimed + wi);
      if (wi == whi) { const int top = (int)(cursor & 31); if (top < 31) fb &= (1u << (top + 1)) - 1u; }
      if (wi == wlo) fb &= ~0u << (int)(lo & 31);
    }
    const unsigned m = __ballot_sync(FULL, fb != 0);
    if (m) {
      const int wl = __ffs(m) - 1;
      const uint32_t f = __shfl_sync(FULL, fb, wl);
      rid = (uint32_t)((wbase - wl) * 32 + (31 - __clz(f)));
      return true;
    }
  }
  return false;
}

// shared memory of one chain, in uint64 words: ref and revref (W words + one zero word each, so that
// window_key may read one word past the bitset), the staged read, Lp packed count columns
// + the chain's cold state (ColdState: touched once per contig, kept out of the register file)
__host__ __device__ inline size_t chain_smem_words(int W, int Lp) { return 3 * (size_t)W + 2 + (size_t)Lp + 4; }  // sizeof(ColdState) = 32
// Warp-uniform chain state that is read or written once per contig, not once per step: kept in shared memory (every lane
// reads the same word -- a broadcast -- and writes the same value), so that it does not occupy six registers per lane for the
// life of the kernel; the hot loops were rematerialising lane ids and shared-memory bases for want of them.
struct ColdState {
  int cursor, slice_lo;                                     // this chain's slice of the read ids still to seed from
  uint32_t first_rid, prev, num_unmatched_1m, n_single;      // contig's first read; last read without a record; reorder.h:433-439; singletons logged
  int first_len, pad;                                        // length of the contig's first read (the left search starts from it again)
};

// WPB warps (= chains) per block, at least MINB blocks per SM: the register budget is the knob that
// decides how many chains co-reside (run_reorder picks the configuration)
// WT > 0: specialised for reads of WT words (W = 5: up to 160 bases, W = 8: up to 256) -- every row address, lane / W
// split, per-chain shared-memory offset and loop bound becomes a constant; WT = 0: any length, from the arguments
// STATS: count the dictionary lookups / Hamming evaluations the oracle counts (parity tests, the roofline's
// algorithmic bytes); the production instantiation leaves the counting out of the hot loops
template <bool LOCKSTEP, int WPB, int MINB, int WT, bool STATS = true>
__global__ void __launch_bounds__(WPB * 32, MINB) k_chains(ChainArgs a) {
  constexpr int kWarpsPerBlock = WPB;
  extern __shared__ __align__(16) uint64_t smem[];
  // read once through volatile asm: the compiler otherwise re-reads %tid (S2R) and rebuilds lane / warp / shared-memory
  // bases dozens of times in the hot loops rather than keep them in registers
#if SB_OPAQUE_TID
  uint32_t tid_once;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_once));
  const int lane = (int)(tid_once & 31u), wib = (int)(tid_once >> 5);
#else
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
#endif
  const uint32_t cid = blockIdx.x * kWarpsPerBlock + wib;
  const int W = WT ? WT : a.W, Lp = WT ? 32 * WT : a.Lp;  // Lp = 32 W for every L (words_for)
  const int grp = lane / W, wig = lane - grp * W;  // scan_bin's lane layout
  const size_t per_chain = chain_smem_words(W, Lp);
  uint64_t *ref = smem + wib * per_chain, *revref = ref + W + 1, *curw = revref + W + 1;
  uint64_t *cnt = curw + W;  // one word per column: four u16 counts {A,C,T,G}
  if (lane == 0) { ref[W] = 0ull; revref[W] = 0ull; }
  __syncwarp();

  int state = cid < a.num_chains ? ST_SEARCH : ST_DONE;
  int ref_len = 0, prev_unmatched = 0, left_search = 0, iter_started = 0, stop_searching = 0, batch = 0, batch_S = 0;
  long long ref_pos = 0, cur_read_pos = 0;
  ColdState &cold = *reinterpret_cast<ColdState *>(cnt + Lp);
  int &cursor = cold.cursor, &slice_lo = cold.slice_lo;  // read ids fit 31 bits (check_input)
  uint32_t &first_rid = cold.first_rid, &prev = cold.prev, &num_unmatched_1m = cold.num_unmatched_1m, &n_single = cold.n_single;
  // One writer: lane 0 stores, __syncwarp() on both sides orders the store against the other lanes' reads (every lane
  // storing the same value is what compute-sanitizer's racecheck rightly calls a hazard).
  auto cold_set = [&](auto &field, auto v) { __syncwarp(); if (lane == 0) field = v; __syncwarp(); };
  if (lane == 0) { cursor = -1; slice_lo = 0; first_rid = 0; prev = 0; num_unmatched_1m = 0; n_single = 0; cold.first_len = 0; }
  __syncwarp();
  uint32_t num_reads_thr = 0, n_aligned = 0, window_left = 0;
  // statistics: c_issued / c_seq / c_slot are per-lane partial sums, c_cmp / c_unmatched / c_lost are
  // warp-uniform; 32-bit in registers, flushed to the 64-bit totals before they can wrap
  uint32_t c_unmatched = 0, c_lost = 0, c_issued = 0, c_seq = 0, c_cmp = 0, c_slot = 0;
  unsigned long long target = 0, round = 0;
  auto flush_counters = [&](bool force) {
    if (!STATS && !force) return;  // only c_lost is counted: one increment per lost claim, flushed at the end
    if (!force && !__any_sync(FULL, (c_issued | c_seq | c_slot | c_cmp | c_lost) >> 30)) return;
    auto wsum = [&](uint32_t v) {  // exact 64-bit warp sum of 32-bit lane values
      return (unsigned long long)__reduce_add_sync(FULL, v & 0xFFFFu) + ((unsigned long long)__reduce_add_sync(FULL, v >> 16) << 16);
    };
    const unsigned long long s_issued = wsum(c_issued), s_seq = wsum(c_seq), s_slot = wsum(c_slot);
    if (lane == 0) {
      atomicAdd(a.ctr + CTR_PROBES_ISSUED, s_issued);
      atomicAdd(a.ctr + CTR_PROBES_SEQ, s_seq);
      atomicAdd(a.ctr + CTR_SLOT_PROBES, s_slot);
      atomicAdd(a.ctr + CTR_COMPARES, (unsigned long long)c_cmp);
      atomicAdd(a.ctr + CTR_LOST, (unsigned long long)c_lost);
    }
    c_issued = c_seq = c_slot = c_cmp = c_lost = 0;
  };

  // claim bit + "remove from both dictionaries" (reorder.h:458-472): one decrement per bin
  auto claim_pre = [&](uint32_t rid, uint32_t sidx) {  // slot index fetched in phase A
    if (lane == 0) atomicOr(a.claimed + (rid >> 5), 1u << (rid & 31));
    if (lane < kNumDict && sidx != 0xFFFFFFFFu) atomicSub(&a.dict[lane].slots[sidx].live, 1u);
  };
  auto stage_read = [&](uint32_t rid) {
    if (lane < W) curw[lane] = __ldg(a.reads + (size_t)rid * W + lane);
    __syncwarp();
  };
  // fold the read staged in curw into the window: word-parallel fast path, or the per-column generic
  // version for the reference's in-place "fold" quirk (and on request, as a cross-check)
  uint32_t depth = 0;  // reads folded into the counts since they were last reset (update_ref_fast: saturation only beyond 65534)
  auto upd = [&](int old_len, int delta, int cs, int cur_len, bool rev, int new_len, int fold) {
    if (old_len == 0) depth = 0;  // counts start from zero: a new contig, or its left search
    if (depth < 65535u) depth++;
    if (fold > 0 || (STATS && a.generic_update)) {
      update_ref(ref, revref, curw, cnt, W, lane, old_len, delta, cs, cur_len, rev, new_len, fold);
      if (fold > 0) depth = 65535u;  // the fold adds several of the read's bases to one column: no bound on the counts after it
    } else {
      update_ref_fast<WT>(ref, revref, curw, cnt, W, lane, old_len, delta, cs, cur_len, rev, new_len, depth > 65534u);
    }
  };
  // the read must already be staged in curw
  auto new_contig = [&](uint32_t rid, int len) {  // updaterefcount(..., resetcount = true, rev = false) + reorder.h:426-430,:601-612
    __syncwarp();
    if (lane == 0) { cold.first_len = len; first_rid = rid; prev = rid; }
    __syncwarp();
    upd(0, 0, 0, len, false, len, 0);
    ref_len = len; ref_pos = 0; cur_read_pos = 0;
    prev_unmatched = 1; left_search = 0;
    state = ST_SEARCH; iter_started = 0; batch = 0; batch_S = 0;
    flush_counters(false);
  };

  if (state == ST_SEARCH) {  // reorder.h:405-431
    const uint32_t first = cid * a.per;
    if (lane == 0) {
      slice_lo = (int)first;
      cursor = cid == a.num_chains - 1 ? (int)a.N - 1 : (int)((cid + 1) * a.per) - 1;
    }
    __syncwarp();
    // reorder.h:411-419: a thread gives up its start read if somebody already took it.  Chains of the
    // free-running schedule start whenever their block gets an SM, so an earlier chain may have claimed
    // `first` through a dictionary match: test-and-set, and on failure pick from the own slice instead.
    unsigned old = 0;
    if (lane == 0) old = atomicOr(a.claimed + (first >> 5), 1u << (first & 31));
    old = __shfl_sync(FULL, old, 0);
    if ((old >> (first & 31)) & 1u) {
      state = ST_NEWREAD;
    } else {
      if (lane < kNumDict) {
        const uint32_t sidx = __ldg(a.dict[lane].slot_of_read + first);
        if (sidx != 0xFFFFFFFFu) atomicSub(&a.dict[lane].slots[sidx].live, 1u);
      }
      c_unmatched++;
      const int first_len = __ldg(a.lens + first);
      stage_read(first);
      new_contig(first, first_len);
    }
  }
  unsigned long long cy_search = 0, cy_wait_a = 0, cy_commit = 0, cy_wait_b = 0;
  long long t_last = clock64();
  if (!LOCKSTEP) {
    // ---------------- free-running schedule ----------------------------------------------------------
    // Every chain runs at its own pace, exactly like a reference thread: a candidate is claimed on
    // the spot with an atomic test-and-set of its claim bit (the reference's remainingreads[] under
    // read_lock, reorder.h:303-309); losing that race just continues the search.  No grid barrier,
    // no proposals; the result depends on timing for more than one chain -- as the reference's does.
    while (state != ST_DONE) {
      if (state == ST_SEARCH) {
        if (!iter_started) {  // loop top, reorder.h:433-439 (the window's end is counted down: no modulo per step)
          if (window_left == 0) {
            if (num_unmatched_1m > kStopUnmatched) stop_searching = 1;
            cold_set(num_unmatched_1m, 0u);
            window_left = kStopWindow;
          }
          window_left--;
          num_reads_thr++;
          iter_started = 1;
        }
        bool found = false;
        uint32_t k = 0, pre_sidx = 0xFFFFFFFFu;
        uint64_t pre_word = 0;
        int shift = 0, prev_rev = 0, pre_len = 0;
        if (!stop_searching) {
          int b = 0, S = 0;
          while (S < a.maxshift) {
            if (chain_search<WT, true, STATS>(a, ref, revref, ref_len, lane, grp, wig, b, S, k, shift, prev_rev, c_issued, c_seq, c_cmp, c_slot)) {
              // the claim's round trip overlaps the loads the update will need (row, length, slot indices)
              unsigned old = 0;
              if (lane == 0) old = atomicOr(a.claimed + (k >> 5), 1u << (k & 31));
              if (lane < W) pre_word = __ldg(a.reads + (size_t)k * W + lane);
              pre_len = __ldg(a.lens + k);
              if (lane < kNumDict) pre_sidx = __ldg(a.dict[lane].slot_of_read + k);
              old = __shfl_sync(FULL, old, 0);
              if (!((old >> (k & 31)) & 1u)) { found = true; break; }
              c_lost++;  // another chain took it between the check and the claim: search this batch again
              continue;
            }
            S += 8 * ((STATS ? a.fast_tail : 1) ? (b < 2 ? (STATS ? a.batch0 : kBatch0) << b : 16) : (b < 4 ? 1 << b : 16));
            b++;
          }
        }
        if (found) {
          if (lane < kNumDict && pre_sidx != 0xFFFFFFFFu) atomicSub(&a.dict[lane].slots[pre_sidx].live, 1u);
          if (lane < W) curw[lane] = pre_word;
          __syncwarp();
          const int len = pre_len, old = ref_len;
          int delta, cs, nl, fold = 0;
          if (!prev_rev) { delta = shift; cs = 0; nl = max(old - shift, len); }
          else if (len - shift >= old) { fold = len - shift - old; delta = -fold; cs = 0; nl = len; }
          else if (old + shift <= a.L) { delta = 0; cs = old - len + shift; nl = old + shift; }
          else { delta = old + shift - a.L; cs = a.L - len; nl = a.L; }
          upd(old, delta, cs, len, prev_rev != 0, nl, fold);
          ref_len = nl;
          if (!prev_rev) {
            if (!left_search) { cur_read_pos = ref_pos + shift; ref_pos = cur_read_pos; }
            else { cur_read_pos = ref_pos + old - shift - len; ref_pos = ref_pos + old - shift - nl; }
          } else {
            if (!left_search) { cur_read_pos = ref_pos + old + shift - len; ref_pos = ref_pos + old + shift - nl; }
            else { cur_read_pos = ref_pos - shift; ref_pos = cur_read_pos; }
          }
          if (lane == 0) {
            if (prev_unmatched) a.rec[prev] = make_rec(0, n_aligned, cid, 0);
            const uint32_t kk = n_aligned + (prev_unmatched ? 1u : 0u);
            const int is_r = prev_rev ? !left_search : left_search;
            a.rec[k] = make_rec(cur_read_pos, kk, cid, 2u | (is_r ? 1u : 0u));
          }
          n_aligned += prev_unmatched ? 2u : 1u;
          prev_unmatched = 0;
          iter_started = 0;
        } else {
          cold_set(num_unmatched_1m, num_unmatched_1m + 1u);
          if (!left_search) {
            left_search = 1;
            stage_read(first_rid);
            const int len = cold.first_len;
            upd(0, 0, 0, len, true, len, 0);
            ref_len = len; ref_pos = 0; cur_read_pos = 0;
            iter_started = 0;
          } else {
            left_search = 0;
            state = ST_NEWREAD;
          }
        }
      } else {  // ST_NEWREAD, reorder.h:576-612
        uint32_t j = 0;
        bool got = false;
        // as for a matched read, the loads the new contig needs (row, length, slot indices) travel under the claim's round trip
        uint32_t seed_sidx = 0xFFFFFFFFu;
        uint64_t seed_word = 0;
        int seed_len = 0;
        auto claim_seed = [&](uint32_t r) {
          unsigned old = 0;
          if (lane == 0) old = atomicOr(a.claimed + (r >> 5), 1u << (r & 31));
          if (lane < W) seed_word = __ldg(a.reads + (size_t)r * W + lane);
          seed_len = __ldg(a.lens + r);
          if (lane < kNumDict) seed_sidx = __ldg(a.dict[lane].slot_of_read + r);
          old = __shfl_sync(FULL, old, 0);
          return !((old >> (r & 31)) & 1u);
        };
        while (find_unclaimed(a.claimed, slice_lo, cursor, lane, j)) {
          const bool mine = claim_seed(j);
          cold_set(cursor, (int)j - 1);
          if (mine) { got = true; break; }
        }
        // Own slice exhausted: instead of idling until the slowest chain is done, seed the next contig from
        // the slice of a randomly chosen other chain (the reference's threads all pick from ONE pool,
        // reorder.h:576-592).  Random victims keep thieves apart; a bounded number of probes bounds the
        // work once nothing is left.  Every slice is still drained by its owner, so coverage is unaffected.
        if (!got && a.steal_probes > 0) {
          uint32_t rnd = cid * 2654435761u + num_reads_thr;
          for (int t = 0; t < a.steal_probes && !got; t++) {
            rnd = rnd * 1664525u + 1013904223u;
            const uint32_t v = (rnd >> 8) % a.num_chains;
            const long long vlo = (long long)v * a.per;
            const long long vhi = v == a.num_chains - 1 ? (long long)a.N - 1 : vlo + a.per - 1;
            if (!find_unclaimed(a.claimed, vlo, vhi, lane, j)) continue;
            if (claim_seed(j)) got = true;
          }
        }
        if (prev_unmatched) {
          if (lane == 0) a.rec[prev] = make_rec(0, n_single, cid, 4);
          cold_set(n_single, n_single + 1u);
        }
        if (got) {
          if (lane < kNumDict && seed_sidx != 0xFFFFFFFFu) atomicSub(&a.dict[lane].slots[seed_sidx].live, 1u);
          c_unmatched++;
          if (lane < W) curw[lane] = seed_word;
          __syncwarp();
          new_contig(j, seed_len);
        } else {
          prev_unmatched = 0;
          state = ST_DONE;
        }
      }
      round++;
    }
    if (lane == 0 && a.chain_dbg) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      a.chain_dbg[2 * cid] = round; a.chain_dbg[2 * cid + 1] = ns;
    }
  }
  if (LOCKSTEP) grid_barrier(a.barrier, target, t_last, cy_commit, cy_wait_b);
  cy_commit = 0; cy_wait_b = 0;
  while (LOCKSTEP) {
    // ---------------- phase A: search / pick against the round-start claim state -------------
    bool has_prop = false, search_more = false;
    uint32_t prop_rid = 0;
    int prop_shift = 0, prop_rev = 0;
    if (state == ST_SEARCH) {
      if (!iter_started) {  // loop top, reorder.h:433-439
        if (num_reads_thr % kStopWindow == 0) {
          if (num_unmatched_1m > kStopUnmatched) stop_searching = 1;
          cold_set(num_unmatched_1m, 0u);
        }
        num_reads_thr++;
        iter_started = 1;
      }
      if (!stop_searching) {
        has_prop = chain_search<WT, false, true>(a, ref, revref, ref_len, lane, grp, wig, batch, batch_S, prop_rid, prop_shift, prop_rev, c_issued,
                                c_seq, c_cmp, c_slot);
        if (!has_prop) {
          const int nshift = 8 * (batch < 4 ? 1 << batch : 16);
          if (batch_S + nshift < a.maxshift) { batch++; batch_S += nshift; search_more = true; }
        }
      }
    } else if (state == ST_NEWREAD) {
      has_prop = find_unclaimed(a.claimed, slice_lo, cursor, lane, prop_rid);
    }
    if (has_prop && lane == 0) atomicMin(a.winner + prop_rid, cid);
    // everything phase B will need is fetched now, so its latency hides behind the barrier:
    // the read to fold in (or the contig's first read when a left search is about to start) and the
    // dictionary slots to decrement
    uint32_t pre_sidx = 0xFFFFFFFFu;
    if (has_prop) {
      if (lane < kNumDict) pre_sidx = __ldg(a.dict[lane].slot_of_read + prop_rid);
      stage_read(prop_rid);
    } else if (state == ST_SEARCH && !search_more && !left_search) {
      stage_read(first_rid);
    }
    grid_barrier(a.barrier, target, t_last, cy_search, cy_wait_a);

    // ---------------- phase B: winners claim and update ----------------------------------------
    if (state == ST_SEARCH) {
      if (has_prop) {
        if (__ldcg(a.winner + prop_rid) == cid) {
          const uint32_t k = prop_rid;
          claim_pre(k, pre_sidx);
          const int len = __ldg(a.lens + k), shift = prop_shift, old = ref_len;
          int delta, cs, nl, fold = 0;
          if (!prop_rev) { delta = shift; cs = 0; nl = max(old - shift, len); }                 // reorder.h:144-156
          else if (len - shift >= old) { fold = len - shift - old; delta = -fold; cs = 0; nl = len; }  // :159-174
          else if (old + shift <= a.L) { delta = 0; cs = old - len + shift; nl = old + shift; } // :175-184
          else { delta = old + shift - a.L; cs = a.L - len; nl = a.L; }                         // :185-199
          upd(old, delta, cs, len, prop_rev != 0, nl, fold);
          ref_len = nl;
          if (!prop_rev) {  // reorder.h:490-497
            if (!left_search) { cur_read_pos = ref_pos + shift; ref_pos = cur_read_pos; }
            else { cur_read_pos = ref_pos + old - shift - len; ref_pos = ref_pos + old - shift - nl; }
          } else {          // reorder.h:528-535
            if (!left_search) { cur_read_pos = ref_pos + old + shift - len; ref_pos = ref_pos + old + shift - nl; }
            else { cur_read_pos = ref_pos - shift; ref_pos = cur_read_pos; }
          }
          if (lane == 0) {
            if (prev_unmatched) {  // the contig's first read is written lazily, reorder.h:498-507
              a.rec[prev] = make_rec(0, n_aligned, cid, 0);
            }
            const uint32_t kk = n_aligned + (prev_unmatched ? 1u : 0u);
            const int is_r = prop_rev ? !left_search : left_search;  // reorder.h:508, :546
            a.rec[k] = make_rec(cur_read_pos, kk, cid, 2u | (is_r ? 1u : 0u));
          }
          n_aligned += prev_unmatched ? 2u : 1u;
          prev_unmatched = 0;
          iter_started = 0; batch = 0; batch_S = 0;
        } else {
          c_lost++;
        }
      } else if (!search_more) {  // no match, reorder.h:559-615
        cold_set(num_unmatched_1m, num_unmatched_1m + 1u);
        if (!left_search) {
          left_search = 1;
          const int len = cold.first_len;
          upd(0, 0, 0, len, true, len, 0);
          ref_len = len; ref_pos = 0; cur_read_pos = 0;
          iter_started = 0; batch = 0; batch_S = 0;
        } else {
          left_search = 0;
          state = ST_NEWREAD;
        }
      }
    } else if (state == ST_NEWREAD) {
      if (has_prop) {
        if (__ldcg(a.winner + prop_rid) == cid) {
          const uint32_t j = prop_rid;
          claim_pre(j, pre_sidx);
          if (lane == 0 && prev_unmatched) a.rec[prev] = make_rec(0, n_single, cid, 4);
          if (prev_unmatched) cold_set(n_single, n_single + 1u);
          cold_set(cursor, (int)j - 1);
          c_unmatched++;
          new_contig(j, __ldg(a.lens + j));
        } else {
          c_lost++;
        }
      } else {
        if (prev_unmatched) {
          if (lane == 0) a.rec[prev] = make_rec(0, n_single, cid, 4);
          cold_set(n_single, n_single + 1u);
        }
        state = ST_DONE;
        if (lane == 0) atomicSub(a.active, 1);
      }
    }
    round++;
    grid_barrier(a.barrier, target, t_last, cy_commit, cy_wait_b);
    if (__ldcg(a.active) <= 0) break;
    if (round >= a.max_rounds) {  // watchdog: uniform across the grid
      if (cid == 0 && lane == 0) a.ctr[CTR_ABORT] = 1ull;
      break;
    }
  }
  flush_counters(true);
  if (lane == 0) {
    a.chain_aligned[cid] = n_aligned;
    a.chain_single[cid] = n_single;
    atomicAdd(a.ctr + CTR_UNMATCHED, (unsigned long long)c_unmatched);
    if (threadIdx.x == 0) {  // per-block numbers (thread 0 runs the barrier)
      atomicAdd(a.ctr + CTR_CYC_SEARCH, cy_search);
      atomicAdd(a.ctr + CTR_CYC_WAIT_A, cy_wait_a);
      atomicAdd(a.ctr + CTR_CYC_COMMIT, cy_commit);
      atomicAdd(a.ctr + CTR_CYC_WAIT_B, cy_wait_b);
    }
    if (cid == 0) a.ctr[CTR_ROUNDS] = round;
  }
}

__global__ void k_make_policies(unsigned long long *out) {
  out[0] = l2_policy_evict_last();
  out[1] = l2_policy_evict_first();
}

// Chain logs -> one stream, chain after chain (what the merge of per-thread files gives,
// encoder.h:386-423): record k of chain c lands at offset[c] + k.
// Two passes: the scatter writes ONE 16-byte record per read at its (random) stream position -- one sector instead of four
// (order, flag, pos, rev live in four arrays) -- and a coalesced pass splits the records into the arrays the encoder reads.
__global__ void k_scatter_records(const uint4 *__restrict__ rec, uint32_t n, const uint32_t *__restrict__ off_aligned,
                                  const uint32_t *__restrict__ off_single, uint4 *__restrict__ stream, uint32_t *s_order) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 r = rec[i];
  const uint32_t m = r.w >> 24, c = r.w & 0xFFFFFFu, k = r.z;
  if (m & 4) s_order[off_single[c] + k] = i;
  else stream[off_aligned[c] + k] = make_uint4(r.x, r.y, i, m);
}
__global__ void k_unpack_records(const uint4 *__restrict__ stream, uint32_t n, const uint32_t *__restrict__ num_aligned, uint32_t *order,
                                 uint8_t *flag, int64_t *pos, uint8_t *rev) {
  uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n || o >= *num_aligned) return;  // the number of aligned records is still on the device (last entry of the scan)
  const uint4 r = stream[o];
  order[o] = r.z;
  flag[o] = (r.w >> 1) & 1;
  pos[o] = (int64_t)((unsigned long long)r.x | ((unsigned long long)r.y << 32));
  rev[o] = (r.w & 1) ? 'r' : 'd';
}

}  // namespace

void run_reorder(Ctx &c, const uint64_t *reads, const uint16_t *lens, uint32_t n, int L, uint32_t num_chains,
                 const DictBuild dict[2], ReorderDev &out) {
  cudaStream_t st = c.stream;
  out = ReorderDev{};
  const int W = words_for(L), Lp = (L + 31) & ~31;
  const uint32_t nn = n ? n : 1;
  out.order = c.pool.dev<uint32_t>("ro.order", nn);
  out.flag = c.pool.dev<uint8_t>("ro.flag", nn);
  out.pos = c.pool.dev<int64_t>("ro.pos", nn);
  out.rev = c.pool.dev<uint8_t>("ro.rev", nn);
  out.s_order = c.pool.dev<uint32_t>("ro.s_order", nn);
  if (n == 0) return;

  const bool lockstep = c.lockstep;
  // launch configuration: warps per block x minimum blocks per SM.  The deterministic schedule is a cooperative launch
  // and keeps 8 x 3; the free-running one defaults to kChainCfgDefault (SPRING_B200_KCFG=8x3|8x4|8x5|8x6|4x7|4x9
  // overrides it: occupancy experiments).
  void (*kern)(ChainArgs) = k_chains<true, 8, 3, 0>;
  int kWarpsPerBlock = 8;
  if (!lockstep) {
    const char *cfg = getenv("SPRING_B200_KCFG");
    const std::string want = cfg ? cfg : kChainCfgDefault;
    const bool generic_w = getenv("SPRING_B200_GENERIC_W") != nullptr;  // A/B: the any-length instantiation
    if (want == "8x4") {
      kWarpsPerBlock = 8;
      // the tuning knobs exist in the counting instantiation only; the production one has their defaults compiled in
      const bool stats = c.chain_stats || getenv("SPRING_B200_FAST_TAIL") || getenv("SPRING_B200_FILTER_HINT") ||
                         getenv("SPRING_B200_PREFETCH") || getenv("SPRING_B200_GENERIC_UPDATE") || getenv("SPRING_B200_BATCH0");
      if (W == 5 && !generic_w) kern = stats ? k_chains<false, 8, 4, 5> : k_chains<false, 8, 4, 5, false>;       // 129..160 bases (150 bp reads)
      else if (W == 8 && !generic_w) kern = stats ? k_chains<false, 8, 4, 8> : k_chains<false, 8, 4, 8, false>;  // 225..256 bases (250 bp reads)
      else if (W == 4 && !generic_w) kern = stats ? k_chains<false, 8, 4, 4> : k_chains<false, 8, 4, 4, false>;  // 97..128 bases (100 / 125 bp reads)
      else kern = stats ? k_chains<false, 8, 4, 0> : k_chains<false, 8, 4, 0, false>;
    }
    else if (want == "8x5") { kern = k_chains<false, 8, 5, 0>; kWarpsPerBlock = 8; }
    else if (want == "8x6") { kern = k_chains<false, 8, 6, 0>; kWarpsPerBlock = 8; }
    else if (want == "4x9") { kern = k_chains<false, 4, 9, 0>; kWarpsPerBlock = 4; }
    else if (want == "4x7") { kern = k_chains<false, 4, 7, 0>; kWarpsPerBlock = 4; }
    else { kern = k_chains<false, 8, 3, 0>; kWarpsPerBlock = 8; }
  }
  const size_t smem = kWarpsPerBlock * chain_smem_words(W, Lp) * sizeof(uint64_t);
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpsPerBlock * 32, smem));
  if (per_sm < 1) throw CudaError("k_chains does not fit on an SM");
  const uint32_t chains_per_block = (uint32_t)kWarpsPerBlock;
  const uint32_t max_chains = (uint32_t)per_sm * c.num_sms * chains_per_block;
  uint32_t C = num_chains;
  if (C == 0) {
    // auto: every co-resident chain, as long as a chain's slice keeps >= 2048 reads.  Every chain costs contig starts:
    // measured (tests/tools/ratio_check.py) the `Reads:` stream grows by about 64 * chains / reads against the reference
    // (+7.6 % at 4 M reads with 4736 chains), so 2048 reads per chain bound the cost at ~3 % for small inputs, where the
    // kernel's speed does not matter, and change nothing from 10 M reads up (4736 chains: 3 % at 10 M, 0.3 % at 100 M).
    // SPRING_B200_READS_PER_CHAIN=6400 keeps it under 1 % at any size; 256 is the round-1 policy.
    static const uint32_t kReadsPerChain = [] {
      const char *e = getenv("SPRING_B200_READS_PER_CHAIN");
      const long v = e ? atol(e) : 2048;
      return (uint32_t)(v < 1 ? 1 : v);
    }();
    C = n / kReadsPerChain; if (C < 1) C = 1; if (C > max_chains) C = max_chains;
    if (const char *e = getenv("SPRING_B200_MAX_CHAINS")) { const uint32_t m = (uint32_t)strtoul(e, nullptr, 10); if (m && C > m) C = m; }
  }
  if (C > max_chains) C = max_chains;
  if (C > n) C = n;
  const uint32_t grid = (C + chains_per_block - 1) / chains_per_block;
  if (C >= (1u << 24)) throw LimitError("reorder: more than 2^24 chains");  // a record keeps the chain id in 24 bits
  out.num_chains = C;

  ChainArgs a{};
  a.reads = reads; a.lens = lens; a.N = n; a.L = L; a.W = W; a.Lp = Lp; a.maxshift = L / 2;
  a.dict[0] = dict[0].view; a.dict[1] = dict[1].view;
  const size_t bm_words = ((size_t)n + 31) / 32;
  a.claimed = c.pool.dev<uint32_t>("ro.claimed", bm_words);
  a.winner = c.pool.dev<uint32_t>("ro.winner", nn);
  a.rec = c.pool.dev<uint4>("ro.rec", nn);
  const uint32_t nslots = grid * chains_per_block;
  a.chain_aligned = c.pool.dev<uint32_t>("ro.chain_aligned", nslots + 1);
  a.chain_single = c.pool.dev<uint32_t>("ro.chain_single", nslots + 1);
  uint32_t *off_aligned = c.pool.dev<uint32_t>("ro.off_aligned", nslots + 1);
  uint32_t *off_single = c.pool.dev<uint32_t>("ro.off_single", nslots + 1);
  unsigned long long *sync = c.pool.dev<unsigned long long>("ro.sync", 2 + CTR_N);
  a.barrier = sync; a.active = reinterpret_cast<int *>(sync + 1); a.ctr = sync + 2;
  a.chain_dbg = getenv("SPRING_B200_CHAIN_DBG") ? c.pool.dev<unsigned long long>("ro.chain_dbg", 2 * (size_t)nslots + 2) : nullptr;
  if (a.chain_dbg) SB_CUDA(cudaMemsetAsync(a.chain_dbg, 0, (2 * (size_t)nslots + 2) * sizeof(unsigned long long), st));
  a.num_chains = C; a.per = n / C;
  a.G = 32u / (uint32_t)W;
  a.leader_mask = 0;
  for (uint32_t g = 0; g < a.G; g++) a.leader_mask |= 1u << (g * W);
  a.generic_update = getenv("SPRING_B200_GENERIC_UPDATE") ? 1 : 0;
  a.prefetch_slots = getenv("SPRING_B200_PREFETCH") ? atoi(getenv("SPRING_B200_PREFETCH")) : 1;
  if (!c.l2_policies[0]) {  // createpolicy.fractional.L2::evict_last / evict_first, evaluated once per context
    unsigned long long *d_pol = c.pool.dev<unsigned long long>("ro.l2pol", 2);
    k_make_policies<<<1, 1, 0, st>>>(d_pol);
    SB_CUDA(cudaMemcpyAsync(c.l2_policies, d_pol, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
  }
  a.pol_keep = c.l2_policies[0]; a.pol_stream = c.l2_policies[1];
  a.fast_tail = getenv("SPRING_B200_FAST_TAIL") ? atoi(getenv("SPRING_B200_FAST_TAIL")) : 1;
  a.batch0 = getenv("SPRING_B200_BATCH0") ? std::max(1, std::min(4, atoi(getenv("SPRING_B200_BATCH0")))) : kBatch0;
  a.filter_hint = getenv("SPRING_B200_FILTER_HINT") ? atoi(getenv("SPRING_B200_FILTER_HINT")) : 1;  // L2 evict_last on the filter words  // -1 to -2 % on configs 2 and 3
  a.steal_probes = lockstep ? 0 : (getenv("SPRING_B200_STEAL") ? atoi(getenv("SPRING_B200_STEAL")) : 64);
  a.max_rounds = 8ull * n + 4096ull;
  SB_CUDA(cudaMemsetAsync(a.claimed, 0, bm_words * sizeof(uint32_t), st));
  if (lockstep) SB_CUDA(cudaMemsetAsync(a.winner, 0xFF, (size_t)n * sizeof(uint32_t), st));  // proposals: deterministic schedule only
  SB_CUDA(cudaMemsetAsync(sync, 0, (2 + CTR_N) * sizeof(unsigned long long), st));
  SB_CUDA(cudaMemsetAsync(a.chain_aligned, 0, (nslots + 1) * sizeof(uint32_t), st));
  SB_CUDA(cudaMemsetAsync(a.chain_single, 0, (nslots + 1) * sizeof(uint32_t), st));
  int active = (int)C;
  SB_CUDA(cudaMemcpyAsync(a.active, &active, sizeof(int), cudaMemcpyHostToDevice, st));
  void *args[] = {&a};
  if (!c.ev_k0) { SB_CUDA(cudaEventCreate(&c.ev_k0)); SB_CUDA(cudaEventCreate(&c.ev_k1)); }
  SB_CUDA(cudaEventRecord(c.ev_k0, st));
  if (lockstep) SB_CUDA(cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(kWarpsPerBlock * 32), args, smem, st));
  else kern<<<grid, kWarpsPerBlock * 32, smem, st>>>(a);
  SB_CUDA(cudaEventRecord(c.ev_k1, st));
  c.launches++;

  size_t need = 0, tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, need, a.chain_aligned, off_aligned, (int)nslots + 1, st); tmp_bytes = need;
  void *tmp = c.pool.device("ro.cubtmp", tmp_bytes);
  need = tmp_bytes; cub::DeviceScan::ExclusiveSum(tmp, need, a.chain_aligned, off_aligned, (int)nslots + 1, st);
  need = tmp_bytes; cub::DeviceScan::ExclusiveSum(tmp, need, a.chain_single, off_single, (int)nslots + 1, st);
  c.launches += 2;
  uint4 *stream = c.pool.dev<uint4>("ro.stream", nn);
  k_scatter_records<<<(n + 255) / 256, 256, 0, st>>>(a.rec, n, off_aligned, off_single, stream, out.s_order);
  k_unpack_records<<<(n + 255) / 256, 256, 0, st>>>(stream, n, off_aligned + nslots, out.order, out.flag, out.pos, out.rev);
  c.launches += 2;
  unsigned long long *h = c.pool.pin<unsigned long long>("ro.hsync", CTR_N + 4);
  SB_CUDA(cudaMemcpyAsync(h, a.ctr, CTR_N * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  uint32_t *htot = reinterpret_cast<uint32_t *>(h + CTR_N);
  SB_CUDA(cudaMemcpyAsync(htot, off_aligned + nslots, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaMemcpyAsync(htot + 1, off_single + nslots, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaStreamSynchronize(st));
  SB_CUDA(cudaGetLastError());
  if (a.chain_dbg) {
    std::vector<unsigned long long> dbg(2 * (size_t)nslots);
    SB_CUDA(cudaMemcpy(dbg.data(), a.chain_dbg, dbg.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> steps, ends;
    unsigned long long t0 = ~0ull;
    for (uint32_t i = 0; i < C; i++) { steps.push_back(dbg[2 * i]); ends.push_back(dbg[2 * i + 1]); if (dbg[2 * i + 1] && dbg[2 * i + 1] < t0) t0 = dbg[2 * i + 1]; }
    std::sort(steps.begin(), steps.end()); std::sort(ends.begin(), ends.end());
    auto pct = [&](std::vector<unsigned long long> &v, double q) { return v[(size_t)(q * (v.size() - 1))]; };
    fprintf(stderr, "[chain_dbg] steps min %llu p50 %llu p90 %llu p99 %llu max %llu | finish (ms after first finisher) p10 %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f\n",
            steps.front(), pct(steps, .5), pct(steps, .9), pct(steps, .99), steps.back(), (pct(ends, .1) - t0) / 1e6,
            (pct(ends, .5) - t0) / 1e6, (pct(ends, .9) - t0) / 1e6, (pct(ends, .99) - t0) / 1e6, (ends.back() - t0) / 1e6);
  }
  if (h[CTR_ABORT]) throw LimitError("reorder: watchdog hit (round limit) -- chain kernel did not converge");
  out.num = htot[0];
  out.num_singletons = htot[1];
  if (out.num + out.num_singletons != n) throw LimitError("reorder: records do not cover all reads");
  out.unmatched = (uint32_t)h[CTR_UNMATCHED];
  out.rounds = h[CTR_ROUNDS]; out.lost = h[CTR_LOST];
  SB_CUDA(cudaEventElapsedTime(&out.ms_kernel, c.ev_k0, c.ev_k1));
  out.cyc[0] = h[CTR_CYC_SEARCH]; out.cyc[1] = h[CTR_CYC_WAIT_A]; out.cyc[2] = h[CTR_CYC_COMMIT]; out.cyc[3] = h[CTR_CYC_WAIT_B];
  out.slot_probes = h[CTR_SLOT_PROBES];
  out.probes_issued = h[CTR_PROBES_ISSUED]; out.probes_seq = h[CTR_PROBES_SEQ]; out.compares = h[CTR_COMPARES];
}

}  // namespace sb
