// chains2.cu -- the free-running chain kernel (default schedule of reorder<>(), reference src/reorder.h:320-641).
//
// What changed against the warp-per-chain kernel in reorder.cu (which stays for the deterministic schedule):
//
//   * A chain owns GL = 16 lanes, two chains share a warp.  Nothing in a chain step needs 32 lanes any more (see
//     the update below), and the kernel is bound by dependent memory round trips, not by bandwidth: two chains
//     per warp put twice the loads in flight per register file and -- whenever the two are on the same path,
//     which is the common "hit in the first batch, verify, claim, update" step -- share every instruction issue.
//     All warp primitives run on the chain's own lane mask; the two halves of a warp are independent threads of
//     control that reconverge at the top of the (flattened) step loop.
//   * One loop iteration = ONE batch of shifts (like the round of the deterministic schedule): a dead-end search
//     is a run of cheap probe-only iterations, so a chain in a dead end holds up its warp partner for one batch,
//     not for the whole search.
//   * updaterefcount (reorder.h:110-220) without a pass over the columns.  The per-column base counts live in
//     shared memory BIT-SLICED: plane k of base b is a bitset over the columns holding bit k of that count, laid
//     out exactly like a read (bits 2c, 2c+1 for column c; bases A / G share one word as its even / odd bits,
//     C / T the other), so that
//        - sliding the window by delta columns is the same funnel shift as for a read, applied to the
//          planes in use (about log2(coverage) of them),
//        - adding a read is a ripple-carry add of its per-base column masks into the planes (about two
//          planes deep on average),
//        - the 4-way vote is only needed in the columns where the read disagrees with the old consensus
//          (at most THRESH_REORDER of them): one ballot gathers the bits of all four counts of a column.
//     About 150 warp instructions per step instead of about 450 for the packed-u16 pass of reorder.cu.
//   * filter-positive probes prefetch their slot (prefetch.global.L2) while the batch's other filter words are
//     still in flight, so the walk over a lane's positives finds its slots in L2.
//
// The search order, the claim protocol, the records and the counters are those of reorder.cu; with one chain the
// output is the reference's single-thread result bit for bit (tests/test_gpu_parity.py::test_free_running_schedule).
#include <cub/cub.cuh>
#include "chain_common.cuh"

namespace sb {
namespace chain {
namespace {

constexpr int kPlanes = 16;                              // planes per base: 16-bit counts
constexpr uint64_t kEven = 0x5555555555555555ull;

// shared memory of one chain, in uint64 words: ref and revref (W words + one zero word each, window_key reads
// one word past the bitset), the staged read, 2 x kPlanes x W count planes
__host__ __device__ inline size_t chain2_smem_words(int W) { return 3 * (size_t)W + 2 + 2 * (size_t)kPlanes * W; }

template <int GL>
struct Lanes {
  int lane, gl, gbase;
  unsigned gmask;
  __device__ __forceinline__ Lanes() {
    lane = threadIdx.x & 31;
    gl = lane & (GL - 1);
    gbase = lane - gl;
    gmask = GL == 32 ? FULL : ((GL == 32 ? 0u : ((1u << (GL & 31)) - 1u)) << gbase);
  }
  __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(gmask, p) >> gbase; }  // bit i: lane gl == i
  __device__ __forceinline__ bool any(bool p) const { return __any_sync(gmask, p) != 0; }
  __device__ __forceinline__ void sync() const { __syncwarp(gmask); }
  __device__ __forceinline__ int reduce_min(int v) const { return __reduce_min_sync(gmask, v); }
  __device__ __forceinline__ unsigned reduce_add(unsigned v) const { return __reduce_add_sync(gmask, v); }
  __device__ __forceinline__ unsigned reduce_max(unsigned v) const { return __reduce_max_sync(gmask, v); }
  template <typename T> __device__ __forceinline__ T shfl(T v, int src_gl) const { return __shfl_sync(gmask, v, gbase + src_gl); }
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// plane k of base pair dp (0: A even / G odd bits, 1: C even / T odd bits), word w
__device__ __forceinline__ uint64_t *plane(uint64_t *pl, int W, int dp, int k) { return pl + ((size_t)dp * kPlanes + k) * W; }

// updaterefcount (reorder.h:110-220) for every case but the in-place "fold" quirk (delta >= 0), on bit-sliced counts.
// The read is staged in curw (cur_len bases; rev: its reverse complement is what enters the window).  New column i
// takes old column i + delta when that lies in [0, old_len), the read covers new columns [cs, cs + cur_len).
// ref[] always equals the column-wise majority of the counts (true after a reset, kept by every update): a column
// the read does not cover keeps its counts, one where the read agrees with the old consensus keeps its base, a new
// column takes the read's base; only columns where read and old consensus differ are re-voted.
// curw is overwritten with the read as oriented in the contig.  kmax: planes that may be non-zero.
template <int GL>
__device__ void planes_update(const Lanes<GL> &g, uint64_t *ref, uint64_t *revref, uint64_t *curw, uint64_t *pl, int W, int &kmax,
                              int old_len, int delta, int cs, int cur_len, bool rev, int new_len) {
  const int gl = g.gl;
  if (rev) {
    uint64_t o = 0;
    if (gl < W) o = revcomp_word(curw, W, cur_len, gl);
    g.sync();
    if (gl < W) curw[gl] = o;
    g.sync();
  }
  // ---- counts: slide the window ------------------------------------------------------------------------
  if (old_len == 0) {  // reset (reorder.h:134-143)
    for (int t = gl; t < kmax * W; t += GL) { pl[t] = 0ull; pl[(size_t)kPlanes * W + t] = 0ull; }
    kmax = 0;
  } else if (delta > 0) {
    for (int q = gl; g.any(q < 2 * W); q += GL) {  // lane = (base pair, word); uniform trip count
      const bool mine = q < 2 * W;
      const int dp = mine && q >= W, w = mine ? q - dp * W : 0;
      for (int k = 0; k < kmax; k++) {
        uint64_t *p = plane(pl, W, dp, k);
        const uint64_t v = mine ? shr_word(p, W, w, 2 * delta) : 0ull;
        g.sync();  // every lane has read plane k before anybody overwrites it
        if (mine) p[w] = v;
      }
    }
  }
  g.sync();
  // ---- counts: add the read (ripple carry through the planes) ---------------------------------------
  for (int q = gl; g.any(q < 2 * W); q += GL) {
    const bool mine = q < 2 * W;
    const int dp = mine && q >= W, w = mine ? q - dp * W : 0;
    uint64_t carry = 0;
    if (mine) {
      const uint64_t cover = range_mask(w, 2 * cs, 2 * (cs + cur_len));
      const uint64_t B = shl_word(curw, W, w, 2 * cs) & cover;
      const uint64_t lo = B & kEven, hi = (B >> 1) & kEven;
      const uint64_t sel = (dp ? hi : ~hi) & cover & kEven;   // columns whose base belongs to this pair
      carry = (sel & ~lo) | ((sel & lo) << 1);                // A / C counted in the even bit, G / T in the odd one
    }
    uint64_t *p0 = plane(pl, W, dp, 0) + w;
    int k = 0;
    while (g.any(carry != 0ull)) {
      if (k == kPlanes) {  // a count would pass 65535: it stays there (count_add saturates the same way)
        if (carry) for (int kk = 0; kk < kPlanes; kk++) p0[(size_t)kk * W] |= carry;
        break;
      }
      if (carry) {
        const uint64_t t = p0[(size_t)k * W];
        p0[(size_t)k * W] = t ^ carry;
        carry &= t;
      }
      k++;
    }
    kmax = max(kmax, min(k, kPlanes));
  }
  // ---- consensus -----------------------------------------------------------------------------------------
  uint64_t nw = 0, mm = 0;
  if (gl < W) {
    const uint64_t MA = range_mask(gl, 0, 2 * (old_len - delta));        // columns that have a source column
    const uint64_t MB = range_mask(gl, 2 * cs, 2 * (cs + cur_len));      // columns the read covers
    const uint64_t A = shr_word(ref, W, gl, 2 * delta) & MA;
    const uint64_t B = shl_word(curw, W, gl, 2 * cs) & MB;
    nw = A | (B & ~MA);
    const uint64_t X = (A ^ B) & MA & MB;
    mm = (X | (X >> 1)) & kEven;  // bit 2t: column 32 * gl + t needs the vote
  }
  g.sync();  // counts written, old ref read
  constexpr int KL = GL / 4;      // planes gathered per ballot: lane = (base, plane within the pass)
  for (;;) {
    const unsigned need = g.ballot(mm != 0ull);
    if (!need) break;
    const int src = __ffs(need) - 1;
    const int bp = g.shfl(mm ? __ffsll((long long)mm) - 1 : 0, src);
    const int b = gl / KL, kk = gl - b * KL;  // b = 2-bit code of the base: A0 G1 C2 T3
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    for (int k0 = 0; k0 < kmax; k0 += KL) {
      const int k = k0 + kk;
      const bool bit = k < kmax && ((plane(pl, W, b >> 1, k)[src] >> (bp + (b & 1))) & 1ull);
      const unsigned bal = g.ballot(bit);
      constexpr unsigned M = (1u << KL) - 1u;
      c0 |= (bal & M) << k0; c1 |= ((bal >> KL) & M) << k0; c2 |= ((bal >> (2 * KL)) & M) << k0; c3 |= ((bal >> (3 * KL)) & M) << k0;
    }
    uint32_t mx = c0, code = 0;  // first strict maximum over rows A,C,T,G (reorder.h:204-212) -> codes 0,2,3,1
    if (c2 > mx) { mx = c2; code = 2; }
    if (c3 > mx) { mx = c3; code = 3; }
    if (c1 > mx) { mx = c1; code = 1; }
    if (gl == src) { nw = (nw & ~(3ull << bp)) | ((uint64_t)code << bp); mm &= mm - 1ull; }
  }
  if (gl < W) ref[gl] = nw;
  g.sync();
  if (gl < W) revref[gl] = revcomp_word(ref, W, new_len, gl);
  g.sync();
}

// The reference's in-place "fold" (reorder.h:159-165: counts moved UP with an ascending loop, so already rewritten
// columns are read again; see update_ref in reorder.cu) and the debugging cross-check: per-column update on packed
// u16 counts in a global scratch row, converted from and back to the planes.  Variable-length input only; rare.
template <int GL>
__device__ __noinline__ void planes_update_generic(const Lanes<GL> &g, uint64_t *ref, uint64_t *revref, const uint64_t *curw,
                                                   uint64_t *pl, int W, int &kmax, uint64_t *cnt, int old_len, int delta, int cs,
                                                   int cur_len, bool rev, int new_len, int fold) {
  const int gl = g.gl, Lp = 32 * W;
  // planes -> packed counts (field shifts of count_add: codes A0 G1 C2 T3 -> 0, 48, 16, 32)
  for (int i = gl; i < Lp; i += GL) {
    uint64_t v = 0;
    if (i < old_len)
      for (int b = 0; b < 4; b++) {
        uint64_t f = 0;
        for (int k = 0; k < kmax; k++) f |= ((plane(pl, W, b >> 1, k)[i >> 5] >> (2 * (i & 31) + (b & 1))) & 1ull) << k;
        v |= f << ((0x20103000u >> (8 * b)) & 0xFFu);
      }
    cnt[i] = v;
  }
  g.sync();
  for (int t = gl; t < kPlanes * W; t += GL) { pl[t] = 0ull; pl[(size_t)kPlanes * W + t] = 0ull; }
  if (gl < W) ref[gl] = 0ull;
  g.sync();
  const int nchunks = (new_len + GL - 1) / GL;
  unsigned fmax = 0;
  for (int cc = 0; cc < nchunks; cc++) {
    const int ck = delta >= 0 ? cc : nchunks - 1 - cc;  // move direction decides the safe order
    const int i = ck * GL + gl;
    const bool in = i < new_len;
    const int src = i + delta;
    uint64_t v = 0;
    if (in && src >= 0 && src < old_len) {
      if (fold > 0) {
        const int r = i % fold, q = i / fold;
        v = cnt[r];
        for (int t = 1; t < q; t++) v = count_add(v, 3 - base_code(curw, cur_len - 1 - (t * fold + r)));
      } else {
        v = cnt[src];
      }
    }
    g.sync();
    if (in) {
      const int ci = i - cs;
      if (ci >= 0 && ci < cur_len) v = count_add(v, rev ? 3 - base_code(curw, cur_len - 1 - ci) : base_code(curw, ci));
      cnt[i] = v;
      const uint32_t f0 = (uint32_t)v & 0xFFFFu, f1 = (uint32_t)(v >> 16) & 0xFFFFu;
      const uint32_t f2 = (uint32_t)(v >> 32) & 0xFFFFu, f3 = (uint32_t)(v >> 48);
      uint32_t mx = f0, code = 0;  // first strict maximum over rows A,C,T,G (reorder.h:204-212) -> codes 0,2,3,1
      if (f1 > mx) { mx = f1; code = 2; }
      if (f2 > mx) { mx = f2; code = 3; }
      if (f3 > mx) { mx = f3; code = 1; }
      fmax = max(fmax, mx);
      if (code) atomicOr(reinterpret_cast<unsigned long long *>(ref + (i >> 5)), (unsigned long long)code << (2 * (i & 31)));
      // packed counts -> planes
      for (int b = 0; b < 4; b++) {
        uint32_t f = (uint32_t)(v >> ((0x20103000u >> (8 * b)) & 0xFFu)) & 0xFFFFu;
        while (f) {
          const int k = __ffs(f) - 1;
          f &= f - 1;
          atomicOr(reinterpret_cast<unsigned long long *>(plane(pl, W, b >> 1, k) + (i >> 5)), 1ull << (2 * (i & 31) + (b & 1)));
        }
      }
    }
    g.sync();
  }
  fmax = g.reduce_max(fmax);
  kmax = fmax ? 32 - __clz(fmax) : 0;
  if (gl < W) revref[gl] = revcomp_word(ref, W, new_len, gl);
  g.sync();
}

// Verify the live reads of one bin, highest id first, at most MAX_SEARCH of them (reorder.h:287-311); see scan_bin
// in reorder.cu.  The chain's GL lanes are cut into GL / W groups of W lanes, one candidate per group.
template <int GL>
__device__ bool scan_bin2(const Lanes<GL> &g, const ChainArgs &a, const DictView &d, uint32_t bs, uint32_t bc, uint32_t r0, uint32_t r1,
                          uint32_t r2, const uint64_t *refsm, bool rev, int s, int ref_len, uint32_t &rid_out, uint32_t &compares) {
  const int W = a.W, gl = g.gl;
  const int grp = gl / W, wig = gl - grp * W;
  const uint32_t G = (uint32_t)(GL / W);
  const bool act = (uint32_t)grp < G;
  unsigned leaders = 0;
  for (uint32_t q = 0; q < G; q++) leaders |= 1u << (q * W);
  const unsigned below = leaders & ((1u << (gl - wig)) - 1u);  // leaders of the groups before mine
  const uint64_t rw = rev ? shl_word(refsm, W, wig, 2 * s) : shr_word(refsm, W, wig, 2 * s);
  int live_before = 0;
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
      cw = __ldg(a.reads + (size_t)rid * W + wig);
      len = __ldg(a.lens + rid);
      live = !is_claimed(a.claimed, rid);
    }
    const unsigned lm = g.ballot(live) & leaders;  // one bit per live candidate, in scan order
    if (prefix_dead) {
      if (lm == 0) dead_to = min(bc, off + G);
      else {
        dead_to = off + (uint32_t)__popc(leaders & ((1u << (__ffs(lm) - 1)) - 1u));
        prefix_dead = false;
        if (gl == 0 && dead_to > t0) atomicMax(d.skip + (bs - 1), dead_to);
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
      const int t2 = __shfl_down_sync(g.gmask, h, o, GL);
      if (wig + o < W) h += t2;
    }
    const unsigned em = g.ballot(ev) & leaders;
    const unsigned pm = g.ballot(ev && h <= kThreshReorder) & leaders;
    if (pm) {
      const int wl = __ffs(pm) - 1;
      rid_out = g.shfl(rid, wl);
      compares += __popc(em & ((2u << wl) - 1u));
      return true;
    }
    compares += __popc(em);
    live_before += __popc(lm);
    if (live_before >= kMaxSearch) break;
  }
  if (prefix_dead && gl == 0 && dead_to > t0) atomicMax(d.skip + (bs - 1), dead_to);
  return false;
}

// One batch of the search (chain_search in reorder.cu): lane gl owns probe kind (gl & 3) = 2 * strand + dict and
// the shifts S + (gl >> 2) + KL * j, j < n, of the batch [S, S + KL * n).
template <int GL>
__device__ bool search_batch(const Lanes<GL> &g, const ChainArgs &a, const uint64_t *ref, const uint64_t *revref, int ref_len, int n, int S,
                             uint32_t &prop_rid, int &prop_shift, int &prop_rev, uint32_t &probes_issued, uint32_t &probes_seq,
                             uint32_t &compares, uint32_t &slot_probes) {
  constexpr int KL = GL / 4;
  const int gl = g.gl;
  const int kind = gl & 3, rev = kind >> 1, sub = gl >> 2;
  const DictView &d = a.dict[kind & 1];
  const uint64_t *src = rev ? revref : ref;
  const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
  // shifts this lane's probe kind may use: forward d.end + s < ref_len (reorder.h:264-265), reverse
  // d.end < ref_len + s and s < d.start (:266-267), all below maxshift
  const int s_lo = rev ? d.end - ref_len + 1 : 0;
  const int s_hi = min(a.maxshift, rev ? d.start : ref_len - d.end);
  const int kbase = 2 * d.start, kstep = rev ? -2 : 2;  // bit position of the window key in src at shift s
  // ---- pass 1: bounds + filter bit for this lane's n probes; positives prefetch their slot ----------------
  unsigned okm = 0, cand = 0;
#pragma unroll 2
  for (int j = 0; j < n; j++) {
    const int s = S + sub + KL * j;
    if (s >= s_lo && s < s_hi) {
      okm |= 1u << j;
      const uint64_t hk = mix64(window_key(src, kbase + kstep * s, d.key_bits));
      if (filter_test_hint(d.filter, d.filter_mask, hk, pol_keep)) {
        cand |= 1u << j;
        prefetch_l2(d.slots + ((uint32_t)hk & d.slot_mask));
      }
    }
  }
  probes_issued += (unsigned)__popc(okm);  // per-lane partial sum, reduced when the chain ends
  // ---- pass 2: resolve hits in priority order ----------------------------------------------------
  int cur_j = -1, found_p = -1;
  uint32_t cur_start1 = 0, cur_count = 0, cur_r0 = 0, cur_r1 = 0, cur_r2 = 0;
  for (;;) {
    while (cur_j < 0 && cand) {  // this lane's next filter positive -> slot table
      const int j = __ffs(cand) - 1;
      cand &= cand - 1;
      const int s = S + sub + KL * j;
      const uint64_t key = window_key(src, kbase + kstep * s, d.key_bits);
      uint32_t h = (uint32_t)mix64(key) & d.slot_mask;
      slot_probes++;
      for (;;) {
        const DictSlot sl = load_slot_hint(d.slots + h, pol_stream);
        if (sl.start1 == 0) break;
        if (sl.key == key) {
          // a bin with no live read is an "empty_bin" (reorder.h:277-281): skipped without a visit
          if (sl.live) { cur_j = j; cur_start1 = sl.start1; cur_count = sl.count; cur_r0 = sl.rid[0]; cur_r1 = sl.rid[1]; cur_r2 = sl.rid[2]; }
          break;
        }
        h = (h + 1) & d.slot_mask;
      }
    }
    const int myp = cur_j >= 0 ? (((S + sub + KL * cur_j) << 2) | kind) : 0x7FFFFFFF;
    const int p = g.reduce_min(myp);
    if (p == 0x7FFFFFFF) break;
    const int ps = p >> 2, pk = p & 3;
    const int owner = (((ps - S) & (KL - 1)) << 2) | pk;
    const uint32_t mb = g.shfl(cur_start1, owner);  // entries follow the header at bins[mb - 1]
    const uint32_t mc = g.shfl(cur_count, owner);
    const uint32_t r0 = g.shfl(cur_r0, owner), r1 = g.shfl(cur_r1, owner), r2 = g.shfl(cur_r2, owner);
    uint32_t rid;
    if (scan_bin2<GL>(g, a, a.dict[pk & 1], mb, mc, r0, r1, r2, (pk >> 1) ? revref : ref, pk >> 1, ps, ref_len, rid, compares)) {
      prop_rid = rid; prop_shift = ps; prop_rev = pk >> 1; found_p = p;
      break;
    }
    if (gl == owner) cur_j = -1;
  }
  // lookups a sequential search would have issued: all of this batch, or those up to the hit
  unsigned seqmask = okm;
  if (found_p >= 0) {
    const int fs = found_p >> 2, fk = found_p & 3;
    const int rel = fs - S - sub;  // this lane's shifts <= fs are j <= rel / KL
    if (rel < 0) seqmask = 0;
    else {
      int jm = rel / KL;
      if ((rel & (KL - 1)) == 0 && kind > fk) jm--;  // same shift, later kind: not reached
      seqmask = jm < 0 ? 0u : (jm >= 31 ? seqmask : seqmask & ((2u << jm) - 1u));
    }
  }
  probes_seq += (unsigned)__popc(seqmask);
  return found_p >= 0;
}

// Highest unclaimed read in [lo, cursor] (reorder.h:576-592), GL bitmap words per step.
template <int GL>
__device__ bool find_unclaimed2(const Lanes<GL> &g, const uint32_t *claimed, long long lo, long long cursor, uint32_t &rid) {
  if (cursor < lo) return false;
  const long long whi = cursor >> 5, wlo = lo >> 5;
  for (long long wbase = whi; wbase >= wlo; wbase -= GL) {
    const long long wi = wbase - g.gl;
    uint32_t fb = 0;
    if (wi >= wlo) {
      fb = ~__ldcg(claimed + wi);
      if (wi == whi) { const int top = (int)(cursor & 31); if (top < 31) fb &= (1u << (top + 1)) - 1u; }
      if (wi == wlo) fb &= ~0u << (int)(lo & 31);
    }
    const unsigned m = g.ballot(fb != 0);
    if (m) {
      const int wl = __ffs(m) - 1;
      const uint32_t f = g.shfl(fb, wl);
      rid = (uint32_t)((wbase - wl) * 32 + (31 - __clz(f)));
      return true;
    }
  }
  return false;
}

// TPB threads per block = TPB / GL chains, at least MINB blocks per SM
template <int GL, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) k_chains2(const __grid_constant__ ChainArgs a) {
  constexpr int kChainsPerBlock = TPB / GL;
  constexpr int KL = GL / 4;
  extern __shared__ __align__(16) uint64_t smem[];
  const Lanes<GL> g;
  const int gl = g.gl;
  const int cib = threadIdx.x / GL;  // chain in block
  const uint32_t cid = blockIdx.x * kChainsPerBlock + cib;
  const int W = a.W;
  const size_t per_chain = chain2_smem_words(W);
  uint64_t *ref = smem + cib * per_chain, *revref = ref + W + 1, *curw = revref + W + 1;
  uint64_t *pl = curw + W;
  for (size_t t = gl; t < per_chain; t += GL) ref[t] = 0ull;  // zero pad words, planes all clear
  g.sync();

  int state = cid < a.num_chains ? ST_SEARCH : ST_DONE;
  int ref_len = 0, prev_unmatched = 0, left_search = 0, iter_started = 0, stop_searching = 0, batch = 0, batch_S = 0, kmax = 0;
  long long ref_pos = 0, cur_read_pos = 0;
  int cursor = -1, slice_lo = 0;  // read ids fit 31 bits (check_input)
  uint32_t first_rid = 0, prev = 0, num_reads_thr = 0, num_unmatched_1m = 0, n_aligned = 0, n_single = 0;
  // statistics: c_issued / c_seq / c_slot are per-lane partial sums, the others chain-uniform
  uint32_t c_unmatched = 0, c_lost = 0, c_issued = 0, c_seq = 0, c_cmp = 0, c_slot = 0, steps = 0;
  auto flush_counters = [&](bool force) {
    if (!force && !g.any(((c_issued | c_seq | c_slot | c_cmp | c_lost) >> 30) != 0)) return;
    auto wsum = [&](uint32_t v) {  // exact 64-bit sum of 32-bit lane values
      return (unsigned long long)g.reduce_add(v & 0xFFFFu) + ((unsigned long long)g.reduce_add(v >> 16) << 16);
    };
    const unsigned long long s_issued = wsum(c_issued), s_seq = wsum(c_seq), s_slot = wsum(c_slot);
    if (gl == 0) {
      atomicAdd(a.ctr + CTR_PROBES_ISSUED, s_issued);
      atomicAdd(a.ctr + CTR_PROBES_SEQ, s_seq);
      atomicAdd(a.ctr + CTR_SLOT_PROBES, s_slot);
      atomicAdd(a.ctr + CTR_COMPARES, (unsigned long long)c_cmp);
      atomicAdd(a.ctr + CTR_LOST, (unsigned long long)c_lost);
    }
    c_issued = c_seq = c_slot = c_cmp = c_lost = 0;
  };
  auto stage_read = [&](uint32_t rid) {
    if (gl < W) curw[gl] = __ldg(a.reads + (size_t)rid * W + gl);
    g.sync();
  };
  // fold the read staged in curw into the window
  auto upd = [&](int old_len, int delta, int cs, int cur_len, bool rev, int new_len, int fold) {
    if (fold > 0 || a.generic_update)
      planes_update_generic<GL>(g, ref, revref, curw, pl, W, kmax, a.cnt_scratch + (size_t)cid * 32 * W, old_len, delta, cs, cur_len, rev,
                                new_len, fold);
    else planes_update<GL>(g, ref, revref, curw, pl, W, kmax, old_len, delta, cs, cur_len, rev, new_len);
  };
  // the read must already be staged in curw
  auto new_contig = [&](uint32_t rid) {  // updaterefcount(..., resetcount = true, rev = false) + reorder.h:426-430,:601-612
    const int len = __ldg(a.lens + rid);
    upd(0, 0, 0, len, false, len, 0);
    ref_len = len; ref_pos = 0; cur_read_pos = 0;
    prev_unmatched = 1; first_rid = rid; prev = rid; left_search = 0;
    state = ST_SEARCH; iter_started = 0; batch = 0; batch_S = 0;
    flush_counters(false);
  };
  auto test_and_set = [&](uint32_t rid) -> bool {  // true: this chain now owns the read
    unsigned old = 0;
    if (gl == 0) old = atomicOr(a.claimed + (rid >> 5), 1u << (rid & 31));
    old = g.shfl(old, 0);
    return !((old >> (rid & 31)) & 1u);
  };
  auto leave_bins = [&](uint32_t rid) {  // "remove from both dictionaries" (reorder.h:458-472): one decrement per bin
    if (gl < kNumDict) {
      const uint32_t sidx = __ldg(a.dict[gl].slot_of_read + rid);
      if (sidx != 0xFFFFFFFFu) atomicSub(&a.dict[gl].slots[sidx].live, 1u);
    }
  };

  if (state == ST_SEARCH) {  // reorder.h:405-431
    const uint32_t first = cid * a.per;
    slice_lo = (int)first;
    cursor = cid == a.num_chains - 1 ? (int)a.N - 1 : (int)((cid + 1) * a.per) - 1;
    // reorder.h:411-419: a thread gives up its start read if somebody already took it (chains start whenever their
    // block gets an SM, so an earlier chain may have claimed `first` through a dictionary match)
    if (!test_and_set(first)) {
      state = ST_NEWREAD;
    } else {
      leave_bins(first);
      c_unmatched++;
      stage_read(first);
      new_contig(first);
    }
  }
  while (state != ST_DONE) {
    if (state == ST_SEARCH) {
      if (!iter_started) {  // loop top, reorder.h:433-439
        if (num_reads_thr % kStopWindow == 0) {
          if (num_unmatched_1m > kStopUnmatched) stop_searching = 1;
          num_unmatched_1m = 0;
        }
        num_reads_thr++;
        iter_started = 1;
        batch = 0; batch_S = 0;
      }
      bool found = false, exhausted = stop_searching != 0;
      uint32_t k = 0, pre_sidx = 0xFFFFFFFFu;
      uint64_t pre_word = 0;
      int shift = 0, prev_rev = 0, pre_len = 0;
      if (!stop_searching) {
        // batch b covers 8, 16, 32, 64, then 128 shifts (at most 32 probes per lane)
        const int nshift = min(8 << min(batch, 4), 32 * KL);
        if (search_batch<GL>(g, a, ref, revref, ref_len, nshift / KL, batch_S, k, shift, prev_rev, c_issued, c_seq, c_cmp, c_slot)) {
          // the claim's round trip overlaps the loads the update will need (row, length, slot indices)
          unsigned old = 0;
          if (gl == 0) old = atomicOr(a.claimed + (k >> 5), 1u << (k & 31));
          if (gl < W) pre_word = __ldg(a.reads + (size_t)k * W + gl);
          pre_len = __ldg(a.lens + k);
          if (gl < kNumDict) pre_sidx = __ldg(a.dict[gl].slot_of_read + k);
          old = g.shfl(old, 0);
          if (!((old >> (k & 31)) & 1u)) found = true;
          else c_lost++;  // another chain took it between the check and the claim: this batch is searched again
        } else {
          batch_S += nshift;
          batch++;
          exhausted = batch_S >= a.maxshift;
        }
      }
      if (found) {
        if (gl < kNumDict && pre_sidx != 0xFFFFFFFFu) atomicSub(&a.dict[gl].slots[pre_sidx].live, 1u);
        if (gl < W) curw[gl] = pre_word;
        g.sync();
        const int len = pre_len, old = ref_len;
        int delta, cs, nl, fold = 0;
        if (!prev_rev) { delta = shift; cs = 0; nl = max(old - shift, len); }                         // reorder.h:144-156
        else if (len - shift >= old) { fold = len - shift - old; delta = -fold; cs = 0; nl = len; }   // :159-174
        else if (old + shift <= a.L) { delta = 0; cs = old - len + shift; nl = old + shift; }         // :175-184
        else { delta = old + shift - a.L; cs = a.L - len; nl = a.L; }                                 // :185-199
        upd(old, delta, cs, len, prev_rev != 0, nl, fold);
        ref_len = nl;
        if (!prev_rev) {  // reorder.h:490-497
          if (!left_search) { cur_read_pos = ref_pos + shift; ref_pos = cur_read_pos; }
          else { cur_read_pos = ref_pos + old - shift - len; ref_pos = ref_pos + old - shift - nl; }
        } else {          // reorder.h:528-535
          if (!left_search) { cur_read_pos = ref_pos + old + shift - len; ref_pos = ref_pos + old + shift - nl; }
          else { cur_read_pos = ref_pos - shift; ref_pos = cur_read_pos; }
        }
        if (gl == 0) {
          if (prev_unmatched) {  // the contig's first read is written lazily, reorder.h:498-507
            a.rec_chain[prev] = cid; a.rec_k[prev] = n_aligned; a.rec_pos[prev] = 0; a.rec_meta[prev] = 0;
          }
          const uint32_t kk = n_aligned + (prev_unmatched ? 1u : 0u);
          const int is_r = prev_rev ? !left_search : left_search;  // reorder.h:508, :546
          a.rec_chain[k] = cid; a.rec_k[k] = kk; a.rec_pos[k] = cur_read_pos; a.rec_meta[k] = (uint8_t)(2 | (is_r ? 1 : 0));
        }
        n_aligned += prev_unmatched ? 2u : 1u;
        prev_unmatched = 0;
        iter_started = 0;
        steps++;
      } else if (exhausted) {  // no match, reorder.h:559-615
        num_unmatched_1m++;
        if (!left_search) {
          left_search = 1;
          stage_read(first_rid);
          const int len = __ldg(a.lens + first_rid);
          upd(0, 0, 0, len, true, len, 0);
          ref_len = len; ref_pos = 0; cur_read_pos = 0;
          iter_started = 0;
        } else {
          left_search = 0;
          state = ST_NEWREAD;
        }
        steps++;
      }
    } else {  // ST_NEWREAD, reorder.h:576-612
      uint32_t j = 0;
      bool got = false;
      while (find_unclaimed2<GL>(g, a.claimed, slice_lo, cursor, j)) {
        cursor = (int)j - 1;
        if (test_and_set(j)) { got = true; break; }
      }
      // Own slice exhausted: seed the next contig from the slice of a randomly chosen other chain (the reference's
      // threads all pick from ONE pool, reorder.h:576-592); see reorder.cu.
      if (!got && a.steal_probes > 0) {
        uint32_t rnd = cid * 2654435761u + num_reads_thr;
        for (int t = 0; t < a.steal_probes && !got; t++) {
          rnd = rnd * 1664525u + 1013904223u;
          const uint32_t v = (rnd >> 8) % a.num_chains;
          const long long vlo = (long long)v * a.per;
          const long long vhi = v == a.num_chains - 1 ? (long long)a.N - 1 : vlo + a.per - 1;
          if (!find_unclaimed2<GL>(g, a.claimed, vlo, vhi, j)) continue;
          if (test_and_set(j)) got = true;
        }
      }
      if (prev_unmatched) {
        if (gl == 0) { a.rec_chain[prev] = cid; a.rec_k[prev] = n_single; a.rec_meta[prev] = 4; }
        n_single++;
      }
      if (got) {
        leave_bins(j);
        c_unmatched++;
        stage_read(j);
        new_contig(j);
      } else {
        prev_unmatched = 0;
        state = ST_DONE;
      }
      steps++;
    }
  }
  if (cid < a.num_chains) {
    if (gl == 0 && a.chain_dbg) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      a.chain_dbg[2 * cid] = steps; a.chain_dbg[2 * cid + 1] = ns;
    }
    flush_counters(true);
    if (gl == 0) {
      a.chain_aligned[cid] = n_aligned;
      a.chain_single[cid] = n_single;
      atomicAdd(a.ctr + CTR_UNMATCHED, (unsigned long long)c_unmatched);
      if (cid == 0) a.ctr[CTR_ROUNDS] = steps;
    }
  }
}

template <int GL, int TPB, int MINB>
Chains2Config config_of(int W) {
  Chains2Config c{};
  c.block_threads = TPB;
  c.chains_per_block = TPB / GL;
  c.smem_bytes = (size_t)c.chains_per_block * chain2_smem_words(W) * sizeof(uint64_t);
  auto kern = k_chains2<GL, TPB, MINB>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem_bytes) != cudaSuccess) {
    cudaGetLastError();
    c.max_blocks_per_sm = 0;
    return c;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, c.smem_bytes) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
  c.max_blocks_per_sm = per_sm;
  return c;
}

}  // namespace

// lanes_per_chain: 16 (default; any read length) or 32 (one chain per warp); blocks of 256 threads (128 for long
// reads, whose planes would not leave room for a second block of 256)
Chains2Config chains2_config(int W, int lanes_per_chain) {
  if (lanes_per_chain == 32) return W <= 8 ? config_of<32, 256, 4>(W) : config_of<32, 128, 4>(W);
  return W <= 8 ? config_of<16, 256, 4>(W) : config_of<16, 128, 4>(W);
}

void chains2_launch(const ChainArgs &a, int lanes_per_chain, uint32_t grid, cudaStream_t st) {
  const int W = a.W;
  const Chains2Config c = chains2_config(W, lanes_per_chain);
  if (lanes_per_chain == 32) {
    if (W <= 8) k_chains2<32, 256, 4><<<grid, c.block_threads, c.smem_bytes, st>>>(a);
    else k_chains2<32, 128, 4><<<grid, c.block_threads, c.smem_bytes, st>>>(a);
  } else {
    if (W <= 8) k_chains2<16, 256, 4><<<grid, c.block_threads, c.smem_bytes, st>>>(a);
    else k_chains2<16, 128, 4><<<grid, c.block_threads, c.smem_bytes, st>>>(a);
  }
}

}  // namespace chain
}  // namespace sb
