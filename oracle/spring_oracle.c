/*
 * spring_oracle.c -- CPU restatement of SPRING's short-read reorder + encode hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's library.
 *
 * What it restates (all citations relative to /root/reference/src):
 *   - 2-bit read layout and dictionary windows        reorder.h:95-108, :750-759
 *   - constructdictionary (keys -> bins, ascending ids) bitset_util.h:74-221
 *   - findpos/remove semantics (live range of a bin)  bitset_util.cpp:20-63
 *   - search_match                                     reorder.h:246-318
 *   - updaterefcount                                   reorder.h:110-220
 *   - the greedy chain driver                          reorder.h:351-627
 *   - correct_order                                    encoder.cpp:177-222
 *   - encoder dictionaries + contig loop               encoder.h:206-369, :609-624
 *   - buildcontig / writecontig / enc_noise            encoder.cpp:32-109, encoder.h:518-537
 *   - unaligned tail + 4-bit records                   encoder.h:426-453, util.cpp:322-348
 *   - read reconstruction (inverse)                    decompress.cpp:263-283, :664-685
 *
 * Scheduling.  The reference runs T OpenMP threads racing through striped locks and is not
 * reproducible for T > 1.  This restatement runs C "chains" (one per reference thread) under
 * a deterministic round-synchronous schedule: in every round each chain searches against the
 * claim state of the round start and *proposes* one read; the lowest chain id wins a
 * contested read, losers retry next round.  For C == 1 this is exactly the reference's
 * single-thread execution (no lock is ever contended), which is how the oracle is pinned:
 * tests/test_oracle.py compares every output stream byte for byte with
 * oracle/_ref/spring_ref --hotpath -t 1.  For C > 1 it is one legal interleaving of the
 * reference's threads, except that chain c seeds new contigs only from its own slice
 * [c*(N/C), (c+1)*(N/C)) instead of from the global top (keeps chains independent).
 *
 * Dictionary representation.  The reference maps keys to bins with BooPHF (BooPHF.h:851-883),
 * which returns arbitrary bins for non-keys and therefore verifies the key against the first
 * read of the bin (reorder.h:282-285).  Any exact key->bin map gives identical results; the
 * oracle uses sorted unique keys + binary search.  Deletion is lazy: a read is "removed"
 * exactly when it has been claimed (true for the single-thread reference: every claimed read
 * is removed from both dictionaries before the next search, reorder.h:458-472).
 *
 * PARITY STATUS: pinned against the reference build (oracle/_ref) at T = C = 1; the reference
 * holds no golden vectors for this path.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_READ_LEN 511          /* params.h:22 */
#define MAXW 16                   /* ceil(2*511/64) */
#define NUM_DICT 2                /* params.h:26,33 */
#define MAX_SEARCH 1000           /* params.h:27,34 */
#define THRESH_REORDER 4          /* params.h:28 */
#define THRESH_ENCODER 24         /* params.h:35 */
#define STOP_WINDOW 1000000u      /* reorder.h:433 */
#define STOP_FRACTION 0.5         /* params.h:31 */
#define MAX_LIST_SIZE_DEFAULT 10000000u   /* encoder.h:215; SPRING_ORACLE_MAX_LIST lowers it so that tests reach it */

/* ------------------------------------------------------------------------------------ */
/* small helpers                                                                          */
/* ------------------------------------------------------------------------------------ */
static void *xmalloc(size_t n) {
  void *p = malloc(n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); }
  return p;
}
static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s ? s : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}
static void *xrealloc(void *q, size_t n) {
  void *p = realloc(q, n ? n : 1);
  if (!p) { fprintf(stderr, "oracle: out of memory (%zu)\n", n); abort(); }
  return p;
}

typedef struct { uint8_t *p; size_t n, cap; } bytevec;
static void bv_reserve(bytevec *v, size_t extra) {
  if (v->n + extra > v->cap) {
    size_t c = v->cap ? v->cap * 2 : 4096;
    while (c < v->n + extra) c *= 2;
    v->p = (uint8_t *)xrealloc(v->p, c);
    v->cap = c;
  }
}
static void bv_push(bytevec *v, const void *src, size_t n) {
  bv_reserve(v, n);
  memcpy(v->p + v->n, src, n);
  v->n += n;
}
static void bv_push1(bytevec *v, uint8_t b) { bv_push(v, &b, 1); }

/* 2-bit code used by reorder (reorder.h:97-106, util.cpp:271-274): A0 G1 C2 T3 */
static const char CODE2CHAR[4] = {'A', 'G', 'C', 'T'};
static inline int char2code(char c) {
  switch (c) { case 'A': return 0; case 'G': return 1; case 'C': return 2; default: return 3; }
}
static inline char revchar(char c) { /* util.h:23-29 */
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C';
               case 'T': return 'A'; default: return 'N'; }
}
static void reverse_complement(const char *s, char *out, int len) { /* util.cpp:376-381 */
  for (int j = 0; j < len; j++) out[j] = revchar(s[len - j - 1]);
}

static int words_for(int max_readlen) { return (2 * max_readlen - 1) / 64 + 1; } /* call_template_functions.cpp:10 */

/* bitsettostring, reorder.h:76-92 */
static void packed_to_string(const uint64_t *w, int len, char *s) {
  for (int j = 0; j < len; j++) s[j] = CODE2CHAR[(w[j >> 5] >> (2 * (j & 31))) & 3];
}
/* chartobitset, bitset_util.h:238-244 */
static void string_to_packed(const char *s, int len, uint64_t *w, int W) {
  memset(w, 0, sizeof(uint64_t) * W);
  for (int j = 0; j < len; j++) w[j >> 5] |= (uint64_t)char2code(s[j]) << (2 * (j & 31));
}
/* std::bitset >>= / <<= on W words */
static void mw_shr(uint64_t *w, int W, int bits) {
  int ws = bits >> 6, bs = bits & 63;
  for (int i = 0; i < W; i++) {
    uint64_t lo = (i + ws < W) ? w[i + ws] : 0, hi = (i + ws + 1 < W) ? w[i + ws + 1] : 0;
    w[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
  }
}
static void mw_shl(uint64_t *w, int W, int bits) {
  int ws = bits >> 6, bs = bits & 63;
  for (int i = W - 1; i >= 0; i--) {
    uint64_t hi = (i - ws >= 0) ? w[i - ws] : 0, lo = (i - ws - 1 >= 0) ? w[i - ws - 1] : 0;
    w[i] = bs ? (hi << bs) | (lo >> (64 - bs)) : hi;
  }
}
/* ((b & mask1) >> 2*start).to_ullong() : nbits <= 64 starting at bit position pos */
static uint64_t mw_extract(const uint64_t *w, int W, int pos, int nbits) {
  int wi = pos >> 6, bs = pos & 63;
  uint64_t lo = (wi < W) ? w[wi] : 0, hi = (wi + 1 < W) ? w[wi + 1] : 0;
  uint64_t v = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
  if (nbits < 64) v &= (((uint64_t)1 << nbits) - 1);
  return v;
}
/* popcount((a ^ b) & mask) with mask = bits [lo, hi)  -- generatemasks, bitset_util.h:223-236 */
static int mw_hamming(const uint64_t *a, const uint64_t *b, int W, int lo, int hi) {
  int h = 0;
  if (hi <= lo) return 0;
  for (int i = 0; i < W; i++) {
    int b0 = i * 64, b1 = b0 + 64;
    if (b1 <= lo || b0 >= hi) continue;
    uint64_t m = ~(uint64_t)0;
    if (lo > b0) m &= ~(uint64_t)0 << (lo - b0);
    if (hi < b1) m &= ~(uint64_t)0 >> (b1 - hi);
    h += __builtin_popcountll((a[i] ^ b[i]) & m);
  }
  return h;
}

/* ------------------------------------------------------------------------------------ */
/* dictionary (bitset_util.h:74-221 + bitset_util.cpp)                                  */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  int start, end;          /* base window [start, end] inclusive */
  uint32_t numkeys, dict_numreads;
  uint64_t *keys;          /* sorted unique */
  uint32_t *bin_start;     /* numkeys + 1 */
  uint32_t *read_id;       /* ascending inside a bin (bitset_util.h:188-206) */
} orc_dict;

typedef struct { uint64_t key; uint32_t rid; } keyrid;
static int cmp_keyrid(const void *a, const void *b) {
  const keyrid *x = (const keyrid *)a, *y = (const keyrid *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->rid < y->rid ? -1 : (x->rid > y->rid);
}
static void dict_from_pairs(orc_dict *d, keyrid *kr, uint32_t n) {
  qsort(kr, n, sizeof(keyrid), cmp_keyrid);
  d->dict_numreads = n;
  d->keys = (uint64_t *)xmalloc(sizeof(uint64_t) * n);
  d->bin_start = (uint32_t *)xmalloc(sizeof(uint32_t) * ((size_t)n + 1));
  d->read_id = (uint32_t *)xmalloc(sizeof(uint32_t) * n);
  uint32_t k = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (i == 0 || kr[i].key != kr[i - 1].key) { d->keys[k] = kr[i].key; d->bin_start[k] = i; k++; }
    d->read_id[i] = kr[i].rid;
  }
  d->bin_start[k] = n;
  d->numkeys = k;
}
static void dict_free(orc_dict *d) { free(d->keys); free(d->bin_start); free(d->read_id); memset(d, 0, sizeof(*d)); }
static int64_t dict_lookup(const orc_dict *d, uint64_t key) {
  int64_t lo = 0, hi = (int64_t)d->numkeys - 1;
  while (lo <= hi) {
    int64_t m = (lo + hi) >> 1;
    if (d->keys[m] == key) return m;
    if (d->keys[m] < key) lo = m + 1; else hi = m - 1;
  }
  return -1;
}

/* reorder dictionaries over 2-bit reads: windows per reorder.h:752-759 */
static void reorder_windows(int L, int start[2], int end[2]) {
  start[0] = L > 100 ? L / 2 - 32 : L / 2 - L * 32 / 100;
  end[0] = L / 2 - 1;
  start[1] = L / 2;
  end[1] = L > 100 ? L / 2 - 1 + 32 : L / 2 - 1 + L * 32 / 100;
}
static void build_reorder_dict(orc_dict *d, const uint64_t *reads, const uint16_t *lens, uint32_t N, int W) {
  keyrid *kr = (keyrid *)xmalloc(sizeof(keyrid) * (size_t)N);
  uint32_t n = 0;
  int nb = 2 * (d->end - d->start + 1);
  for (uint32_t i = 0; i < N; i++) {
    if (lens[i] <= d->end) continue; /* bitset_util.h:99-105 */
    kr[n].key = mw_extract(reads + (size_t)i * W, W, 2 * d->start, nb);
    kr[n].rid = i;
    n++;
  }
  dict_from_pairs(d, kr, n);
  free(kr);
}

/* exported: keys + CSR of one reorder dictionary, for parity tests of the GPU dictionary build.
 * keys_out/bin_start_out/read_id_out must hold N, N+1, N entries.  returns numkeys. */
uint32_t orc_reorder_dict(const uint64_t *reads, const uint16_t *lens, uint32_t N, int max_readlen, int which,
                          uint64_t *keys_out, uint32_t *bin_start_out, uint32_t *read_id_out,
                          uint32_t *dict_numreads_out) {
  int W = words_for(max_readlen), s[2], e[2];
  reorder_windows(max_readlen, s, e);
  orc_dict d; memset(&d, 0, sizeof(d));
  d.start = s[which]; d.end = e[which];
  build_reorder_dict(&d, reads, lens, N, W);
  memcpy(keys_out, d.keys, sizeof(uint64_t) * d.numkeys);
  memcpy(bin_start_out, d.bin_start, sizeof(uint32_t) * ((size_t)d.numkeys + 1));
  memcpy(read_id_out, d.read_id, sizeof(uint32_t) * d.dict_numreads);
  *dict_numreads_out = d.dict_numreads;
  uint32_t nk = d.numkeys;
  dict_free(&d);
  return nk;
}

/* ------------------------------------------------------------------------------------ */
/* reorder                                                                                */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  uint64_t probes;        /* dictionary lookups issued (search_match loop bodies reaching lookup) */
  uint64_t probe_hits;    /* lookups that found a non-empty bin with the right key */
  uint64_t compares;      /* Hamming evaluations */
  uint64_t passes;        /* Hamming <= thresh */
  uint64_t rounds;        /* scheduler rounds */
  uint64_t lost_proposals;
  uint32_t unmatched;     /* "Reordering done, X were unmatched" (reorder.h:633-635) */
} orc_counters;

typedef struct {
  /* aligned stream, chain after chain (what reorder threads write, reorder.h:498-512) */
  uint32_t *order; uint8_t *flag; int64_t *pos; uint8_t *rc; uint64_t n;
  /* singletons, chain after chain (reorder.h:594-609) */
  uint32_t *s_order; uint64_t n_s;
  orc_counters ctr;
} orc_reorder_out;

enum { ST_SEARCH = 0, ST_NEWREAD = 1, ST_DONE = 2 };

typedef struct {
  uint64_t ref[MAXW], revref[MAXW];
  int *count[4];
  int ref_len;
  int64_t ref_pos, cur_read_pos;
  int64_t first_rid, current, prev;
  int prev_unmatched, left_search;
  int state, iter_started, stop_searching;
  int batch, batch_S, search_more; /* shift batching of the schedule, see chain_propose */
  uint32_t num_reads_thr, num_unmatched_past_1M;
  int64_t cursor, slice_lo;
  int has_prop, prop_shift, prop_rev;
  uint32_t prop_rid;
  /* log */
  uint32_t *order; uint8_t *flag; int64_t *pos; uint8_t *rc; uint64_t n, cap;
  uint32_t *s_order; uint64_t n_s, cap_s;
} chain_t;

typedef struct {
  const uint64_t *reads; const uint16_t *lens; uint32_t N; int L, W, maxshift;
  orc_dict dict[NUM_DICT];
  uint8_t *claimed;
  orc_counters ctr;
} reorder_ctx;

static void chain_emit(chain_t *c, uint32_t order, uint8_t flag, int64_t pos, uint8_t rc) {
  if (c->n == c->cap) {
    c->cap = c->cap ? c->cap * 2 : 256;
    c->order = (uint32_t *)xrealloc(c->order, c->cap * sizeof(uint32_t));
    c->flag = (uint8_t *)xrealloc(c->flag, c->cap);
    c->pos = (int64_t *)xrealloc(c->pos, c->cap * sizeof(int64_t));
    c->rc = (uint8_t *)xrealloc(c->rc, c->cap);
  }
  c->order[c->n] = order; c->flag[c->n] = flag; c->pos[c->n] = pos; c->rc[c->n] = rc; c->n++;
}
static void chain_emit_singleton(chain_t *c, uint32_t order) {
  if (c->n_s == c->cap_s) {
    c->cap_s = c->cap_s ? c->cap_s * 2 : 256;
    c->s_order = (uint32_t *)xrealloc(c->s_order, c->cap_s * sizeof(uint32_t));
  }
  c->s_order[c->n_s++] = order;
}

/* updaterefcount, reorder.h:110-220.  count index order A,C,T,G (reorder.h:120-123). */
static void updaterefcount(const reorder_ctx *x, chain_t *c, const uint64_t *cur, int resetcount, int rev,
                           int shift, int cur_readlen) {
  static const char inttochar[4] = {'A', 'C', 'T', 'G'};
  char s[MAX_READ_LEN + 1], s1[MAX_READ_LEN + 1], revcur[MAX_READ_LEN + 1], *current;
  int L = x->L, ref_len = c->ref_len;
  int **count = c->count;
#define CH2I(a) ((((uint8_t)(a)) & 0x06) >> 1)
  packed_to_string(cur, cur_readlen, s);
  if (!rev) current = s;
  else { reverse_complement(s, s1, cur_readlen); current = s1; }
  if (resetcount) { /* :133-142 */
    for (int j = 0; j < 4; j++) memset(count[j], 0, sizeof(int) * L);
    for (int i = 0; i < cur_readlen; i++) count[CH2I(current[i])][i] = 1;
    ref_len = cur_readlen;
  } else {
    if (!rev) { /* :144-156 */
      for (int i = 0; i < ref_len - shift; i++) {
        for (int j = 0; j < 4; j++) count[j][i] = count[j][i + shift];
        if (i < cur_readlen) count[CH2I(current[i])][i] += 1;
      }
      for (int i = ref_len - shift; i < cur_readlen; i++) {
        for (int j = 0; j < 4; j++) count[j][i] = 0;
        count[CH2I(current[i])][i] = 1;
      }
      ref_len = (ref_len - shift > cur_readlen) ? ref_len - shift : cur_readlen;
    } else { /* :157-201 */
      if (cur_readlen - shift >= ref_len) {
        for (int i = cur_readlen - shift - ref_len; i < cur_readlen - shift; i++) {
          for (int j = 0; j < 4; j++) count[j][i] = count[j][i - (cur_readlen - shift - ref_len)];
          count[CH2I(current[i])][i] += 1;
        }
        for (int i = 0; i < cur_readlen - shift - ref_len; i++) {
          for (int j = 0; j < 4; j++) count[j][i] = 0;
          count[CH2I(current[i])][i] = 1;
        }
        for (int i = cur_readlen - shift; i < cur_readlen; i++) {
          for (int j = 0; j < 4; j++) count[j][i] = 0;
          count[CH2I(current[i])][i] = 1;
        }
        ref_len = cur_readlen;
      } else if (ref_len + shift <= L) {
        for (int i = ref_len - cur_readlen + shift; i < ref_len; i++)
          count[CH2I(current[i - (ref_len - cur_readlen + shift)])][i] += 1;
        for (int i = ref_len; i < ref_len + shift; i++) {
          for (int j = 0; j < 4; j++) count[j][i] = 0;
          count[CH2I(current[i - (ref_len - cur_readlen + shift)])][i] = 1;
        }
        ref_len = ref_len + shift;
      } else {
        for (int i = 0; i < L - shift; i++)
          for (int j = 0; j < 4; j++) count[j][i] = count[j][i + (ref_len + shift - L)];
        for (int i = L - cur_readlen; i < L - shift; i++) count[CH2I(current[i - (L - cur_readlen)])][i] += 1;
        for (int i = L - shift; i < L; i++) {
          for (int j = 0; j < 4; j++) count[j][i] = 0;
          count[CH2I(current[i - (L - cur_readlen)])][i] = 1;
        }
        ref_len = L;
      }
    }
    for (int i = 0; i < ref_len; i++) { /* :204-212 */
      int max = 0, indmax = 0;
      for (int j = 0; j < 4; j++)
        if (count[j][i] > max) { max = count[j][i]; indmax = j; }
      current[i] = inttochar[indmax];
    }
  }
#undef CH2I
  string_to_packed(current, ref_len, c->ref, x->W);
  reverse_complement(current, revcur, ref_len);
  string_to_packed(revcur, ref_len, c->revref, x->W);
  c->ref_len = ref_len;
}

/* search_match, reorder.h:246-318, against the claim state frozen at round start.
 * b = ref (or revref) already shifted by `shift` bases. returns 1 and *k on success. */
static int search_match(reorder_ctx *x, const uint64_t *b, int rev, int shift, int ref_len, uint32_t *k) {
  for (int l = 0; l < NUM_DICT; l++) {
    const orc_dict *d = &x->dict[l];
    if (!rev) { if (d->end + shift >= ref_len) continue; }
    else { if (d->end >= ref_len + shift || d->start <= shift) continue; }
    uint64_t key = mw_extract(b, x->W, 2 * d->start, 2 * (d->end - d->start + 1));
    x->ctr.probes++;
    int64_t bin = dict_lookup(d, key);
    if (bin < 0) continue;
    int64_t lo = d->bin_start[bin], hi = d->bin_start[bin + 1];
    int live = 0, counted_hit = 0;
    for (int64_t i = hi - 1; i >= lo; i--) {
      uint32_t rid = d->read_id[i];
      if (x->claimed[rid]) continue;          /* removed from the bin (bitset_util.cpp:37-63) */
      if (!counted_hit) { x->ctr.probe_hits++; counted_hit = 1; }
      if (++live > MAX_SEARCH) break;          /* :287-288 */
      int len = x->lens[rid], h;
      if (!rev) { int m = ref_len - shift < len ? ref_len - shift : len; h = mw_hamming(b, x->reads + (size_t)rid * x->W, x->W, 0, 2 * m); }
      else { int m = ref_len + shift < len ? ref_len + shift : len; h = mw_hamming(b, x->reads + (size_t)rid * x->W, x->W, 2 * shift, 2 * m); }
      x->ctr.compares++;
      if (h <= THRESH_REORDER) { x->ctr.passes++; *k = rid; return 1; }
    }
  }
  return 0;
}

static void chain_new_contig(reorder_ctx *x, chain_t *c, int64_t rid) {
  c->current = rid;
  updaterefcount(x, c, x->reads + (size_t)rid * x->W, 1, 0, 0, x->lens[rid]);
  c->ref_pos = 0; c->cur_read_pos = 0;
  c->prev_unmatched = 1; c->first_rid = rid; c->prev = rid;
  c->left_search = 0;
  c->state = ST_SEARCH; c->iter_started = 0; c->batch = 0; c->batch_S = 0;
}

/* phase A of a round for one chain */
static void chain_propose(reorder_ctx *x, chain_t *c) {
  c->has_prop = 0;
  if (c->state == ST_SEARCH) {
    if (!c->iter_started) { /* loop top, reorder.h:433-439 */
      if (c->num_reads_thr % STOP_WINDOW == 0) {
        if (c->num_unmatched_past_1M > STOP_FRACTION * STOP_WINDOW) c->stop_searching = 1;
        c->num_unmatched_past_1M = 0;
      }
      c->num_reads_thr++;
      c->iter_started = 1;
    }
    c->search_more = 0;
    if (c->stop_searching) return;
    /* The shift loop of reorder.h:479-558 is cut into batches of 8, 16, 32, 64, 128, 128, ... shifts
     * and a chain examines ONE batch per round (bounded work per round keeps the lock-step chains
     * balanced).  Claims only ever grow, so a batch that found nothing in an earlier round would
     * find nothing now either: the read proposed is exactly the one a full search against the
     * current round's claim state returns. */
    int nshift = 8 * (c->batch < 4 ? 1 << c->batch : 16);
    int s_end = c->batch_S + nshift < x->maxshift ? c->batch_S + nshift : x->maxshift;
    uint64_t ref[MAXW], revref[MAXW];
    memcpy(ref, c->ref, sizeof(ref)); memcpy(revref, c->revref, sizeof(revref));
    mw_shr(ref, x->W, 2 * c->batch_S); mw_shl(revref, x->W, 2 * c->batch_S);
    for (int shift = c->batch_S; shift < s_end; shift++) { /* :479-558 */
      uint32_t k;
      if (search_match(x, ref, 0, shift, c->ref_len, &k)) { c->has_prop = 1; c->prop_rid = k; c->prop_shift = shift; c->prop_rev = 0; return; }
      if (search_match(x, revref, 1, shift, c->ref_len, &k)) { c->has_prop = 1; c->prop_rid = k; c->prop_shift = shift; c->prop_rev = 1; return; }
      mw_shl(revref, x->W, 2); mw_shr(ref, x->W, 2);
    }
    if (s_end < x->maxshift) { c->batch++; c->batch_S = s_end; c->search_more = 1; } /* next batch next round */
  } else if (c->state == ST_NEWREAD) { /* :576-592 */
    for (int64_t j = c->cursor; j >= c->slice_lo; j--)
      if (!x->claimed[j]) { c->has_prop = 1; c->prop_rid = (uint32_t)j; return; }
  }
}

/* phase B of a round for one chain */
static void chain_commit(reorder_ctx *x, chain_t *c, const uint32_t *winner, int cid) {
  if (c->state == ST_SEARCH) {
    if (c->has_prop) {
      if (winner[c->prop_rid] != (uint32_t)cid) { x->ctr.lost_proposals++; return; } /* lost: retry the same iteration */
      uint32_t k = c->prop_rid; int shift = c->prop_shift, rev = c->prop_rev;
      x->claimed[k] = 1;
      c->current = k;
      int ref_len_old = c->ref_len, len = x->lens[k];
      updaterefcount(x, c, x->reads + (size_t)k * x->W, 0, rev, shift, len);
      if (!rev) { /* :490-497 */
        if (!c->left_search) { c->cur_read_pos = c->ref_pos + shift; c->ref_pos = c->cur_read_pos; }
        else { c->cur_read_pos = c->ref_pos + ref_len_old - shift - len; c->ref_pos = c->ref_pos + ref_len_old - shift - c->ref_len; }
      } else { /* :528-535 */
        if (!c->left_search) { c->cur_read_pos = c->ref_pos + ref_len_old + shift - len; c->ref_pos = c->ref_pos + ref_len_old + shift - c->ref_len; }
        else { c->cur_read_pos = c->ref_pos - shift; c->ref_pos = c->cur_read_pos; }
      }
      if (c->prev_unmatched) chain_emit(c, (uint32_t)c->prev, 0, 0, 'd'); /* :498-507 */
      uint8_t rc = rev ? (c->left_search ? 'd' : 'r') : (c->left_search ? 'r' : 'd'); /* :508,:546 */
      chain_emit(c, k, 1, c->cur_read_pos, rc);
      c->prev_unmatched = 0;
      c->iter_started = 0; c->batch = 0; c->batch_S = 0;
      return;
    }
    if (c->search_more) return; /* more shifts to examine next round */
    /* no match, :559-615 */
    c->num_unmatched_past_1M++;
    if (!c->left_search) {
      c->left_search = 1;
      updaterefcount(x, c, x->reads + (size_t)c->first_rid * x->W, 1, 1, 0, x->lens[c->first_rid]);
      c->ref_pos = 0; c->cur_read_pos = 0;
      c->iter_started = 0; c->batch = 0; c->batch_S = 0;
    } else {
      c->left_search = 0;
      c->state = ST_NEWREAD;
    }
    return;
  }
  if (c->state == ST_NEWREAD) {
    if (c->has_prop) {
      if (winner[c->prop_rid] != (uint32_t)cid) { x->ctr.lost_proposals++; return; }
      uint32_t j = c->prop_rid;
      x->claimed[j] = 1;
      c->cursor = (int64_t)j - 1;
      x->ctr.unmatched++;
      if (c->prev_unmatched) chain_emit_singleton(c, (uint32_t)c->prev); /* :606-609 */
      chain_new_contig(x, c, j);
    } else {
      if (c->prev_unmatched) chain_emit_singleton(c, (uint32_t)c->prev); /* :594-598 */
      c->state = ST_DONE;
    }
  }
}

void orc_reorder_free(orc_reorder_out *o) {
  free(o->order); free(o->flag); free(o->pos); free(o->rc); free(o->s_order);
  memset(o, 0, sizeof(*o));
}

/* reorder<>(), reorder.h:320-641, C chains under the round-synchronous schedule. */
int orc_reorder(const uint64_t *reads, const uint16_t *lens, uint32_t N, int max_readlen, int num_chains,
                orc_reorder_out *out) {
  memset(out, 0, sizeof(*out));
  if (max_readlen < 1 || max_readlen > MAX_READ_LEN) return -1;
  reorder_ctx x; memset(&x, 0, sizeof(x));
  x.reads = reads; x.lens = lens; x.N = N; x.L = max_readlen; x.W = words_for(max_readlen);
  x.maxshift = max_readlen / 2; /* reorder.h:750 */
  int C = num_chains < 1 ? 1 : num_chains;
  if (N == 0) return 0;
  if ((uint32_t)C > N) C = (int)N; /* every chain needs a distinct seed read */
  int s[2], e[2];
  reorder_windows(max_readlen, s, e);
  for (int l = 0; l < NUM_DICT; l++) { x.dict[l].start = s[l]; x.dict[l].end = e[l]; build_reorder_dict(&x.dict[l], reads, lens, N, x.W); }
  x.claimed = (uint8_t *)xcalloc(N, 1);
  uint32_t *winner = (uint32_t *)xmalloc(sizeof(uint32_t) * (size_t)N);
  memset(winner, 0xff, sizeof(uint32_t) * (size_t)N);
  chain_t *ch = (chain_t *)xcalloc(C, sizeof(chain_t));
  uint32_t per = N / (uint32_t)C; /* reorder.h:419-420 */
  for (int c = 0; c < C; c++) {
    chain_t *k = &ch[c];
    for (int j = 0; j < 4; j++) k->count[j] = (int *)xcalloc(max_readlen, sizeof(int));
    int64_t first = (int64_t)c * per;
    k->slice_lo = first;
    k->cursor = (c == C - 1) ? (int64_t)N - 1 : (int64_t)(c + 1) * per - 1;
    if (first >= (int64_t)N || x.claimed[first]) { k->state = ST_DONE; continue; } /* :411-418 */
    x.claimed[first] = 1;
    x.ctr.unmatched++;
    chain_new_contig(&x, k, first);
  }
  for (;;) {
    int active = 0;
    for (int c = 0; c < C; c++) if (ch[c].state != ST_DONE) { active = 1; chain_propose(&x, &ch[c]); }
    if (!active) break;
    x.ctr.rounds++;
    for (int c = 0; c < C; c++)
      if (ch[c].state != ST_DONE && ch[c].has_prop && winner[ch[c].prop_rid] > (uint32_t)c) winner[ch[c].prop_rid] = (uint32_t)c;
    for (int c = 0; c < C; c++) if (ch[c].state != ST_DONE) chain_commit(&x, &ch[c], winner, c);
  }
  uint64_t n = 0, ns = 0;
  for (int c = 0; c < C; c++) { n += ch[c].n; ns += ch[c].n_s; }
  out->order = (uint32_t *)xmalloc(n * sizeof(uint32_t)); out->flag = (uint8_t *)xmalloc(n);
  out->pos = (int64_t *)xmalloc(n * sizeof(int64_t)); out->rc = (uint8_t *)xmalloc(n);
  out->s_order = (uint32_t *)xmalloc(ns * sizeof(uint32_t));
  for (int c = 0; c < C; c++) {
    chain_t *k = &ch[c];
    memcpy(out->order + out->n, k->order, k->n * sizeof(uint32_t));
    memcpy(out->flag + out->n, k->flag, k->n);
    memcpy(out->pos + out->n, k->pos, k->n * sizeof(int64_t));
    memcpy(out->rc + out->n, k->rc, k->n);
    out->n += k->n;
    memcpy(out->s_order + out->n_s, k->s_order, k->n_s * sizeof(uint32_t));
    out->n_s += k->n_s;
    free(k->order); free(k->flag); free(k->pos); free(k->rc); free(k->s_order);
    for (int j = 0; j < 4; j++) free(k->count[j]);
  }
  out->ctr = x.ctr;
  free(ch); free(winner); free(x.claimed);
  for (int l = 0; l < NUM_DICT; l++) dict_free(&x.dict[l]);
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* encoder                                                                                */
/* ------------------------------------------------------------------------------------ */
/* 3-bit code of encoder bitsets (encoder.h:499-514): A0 N1 G2 C4 T6 */
static inline int code3(char c) {
  switch (c) { case 'A': return 0; case 'N': return 1; case 'G': return 2; case 'C': return 4; default: return 6; }
}
static char ENC_NOISE[128][128]; /* encoder.h:518-537 */
static char DEC_NOISE[128][128]; /* decompress.cpp:664-685 */
static int tables_ready = 0;
static void init_tables(void) {
  if (tables_ready) return;
  memset(ENC_NOISE, 0, sizeof(ENC_NOISE)); memset(DEC_NOISE, 0, sizeof(DEC_NOISE));
  const char *rows[5] = {"ACGTN", "CAGTN", "GTACN", "TGCAN", "NAGCT"};
  for (int r = 0; r < 5; r++)
    for (int k = 0; k < 4; k++) {
      ENC_NOISE[(uint8_t)rows[r][0]][(uint8_t)rows[r][k + 1]] = (char)('0' + k);
      DEC_NOISE[(uint8_t)rows[r][0]][(uint8_t)('0' + k)] = rows[r][k + 1];
    }
  tables_ready = 1;
}

typedef struct {
  bytevec seq;        /* consensus, ASCII (read_seq.bin.<t> before 2-bit packing, encoder.cpp:87) */
  bytevec pos;        /* u64 per aligned read, absolute (read_pos.bin after encoder.h:473-487) */
  bytevec noise;      /* read_noise.txt */
  bytevec noisepos;   /* read_noisepos.bin (u16) */
  bytevec rc;         /* read_rev.txt */
  bytevec order;      /* read_order.bin (u32): aligned then unaligned */
  bytevec lengths;    /* read_lengths.bin (u16): aligned then unaligned */
  bytevec unaligned;  /* read_unaligned.txt (4-bit records) */
  uint64_t unaligned_len;  /* read_unaligned.txt.count */
  uint64_t num_aligned;
  uint32_t matched_s, matched_N;  /* encoder.h:490-492 */
  uint64_t enc_probes, enc_compares;
} orc_encode_out;

void orc_encode_free(orc_encode_out *o) {
  free(o->seq.p); free(o->pos.p); free(o->noise.p); free(o->noisepos.p); free(o->rc.p);
  free(o->order.p); free(o->lengths.p); free(o->unaligned.p);
  memset(o, 0, sizeof(*o));
}

typedef struct { size_t off; int64_t pos; uint8_t rc; uint32_t order; uint16_t len; uint64_t seqno; } contig_read; /* off: string offset in the contig arena */
static int cmp_contig_read(const void *a, const void *b) { /* stable (list::sort) via seqno */
  const contig_read *x = (const contig_read *)a, *y = (const contig_read *)b;
  if (x->pos != y->pos) return x->pos < y->pos ? -1 : 1;
  return x->seqno < y->seqno ? -1 : (x->seqno > y->seqno);
}

/* write_dnaN_in_bits, util.cpp:322-348 */
static void write_dnaN_record(bytevec *v, const char *s, uint16_t len) {
  bv_push(v, &len, 2);
  for (int i = 0; i < (len + 1) / 2; i++) {
    uint8_t b = 0;
    for (int j = 0; j < 2 && 2 * i + j < len; j++) {
      char c = s[2 * i + j];
      int code = c == 'N' ? 4 : char2code(c);
      b |= (uint8_t)(code << (4 * j));
    }
    bv_push1(v, b);
  }
}

/* correct_order, encoder.cpp:177-222: clean index -> index in the original FASTQ.
 * order_N ascending original indices of N reads. in-place on `order` (n entries). */
void orc_correct_order(uint32_t *order, uint64_t n, const uint32_t *order_N, uint32_t num_N, uint32_t num_total) {
  uint32_t num_clean = num_total - num_N;
  uint8_t *flagN = (uint8_t *)xcalloc(num_total, 1);
  for (uint32_t i = 0; i < num_N; i++) flagN[order_N[i]] = 1;
  uint32_t *cum = (uint32_t *)xmalloc(sizeof(uint32_t) * (size_t)num_clean);
  uint32_t p = 0, nn = 0;
  for (uint32_t i = 0; i < num_total; i++) { if (flagN[i]) nn++; else cum[p++] = nn; }
  for (uint64_t i = 0; i < n; i++) order[i] += cum[order[i]];
  free(flagN); free(cum);
}

/* encoder_main + encode, encoder.h:124-494, 572-633, single stream (T == 1 semantics).
 *   reads/lens         : the N clean reads, 2-bit packed, W words each
 *   st_*               : aligned stream from reorder (clean indices)
 *   s_order            : singleton clean indices, n_s
 *   nrec, nrec_bytes   : input_N.dna contents (4-bit records), num_N of them
 *   order_N            : read_order_N.bin
 *   num_total          : cp.num_reads                                                     */
int orc_encode(const uint64_t *reads, const uint16_t *lens, uint32_t N, int max_readlen,
               const uint32_t *st_order, const uint8_t *st_flag, const int64_t *st_pos, const uint8_t *st_rc, uint64_t n_st,
               const uint32_t *s_order, uint64_t n_s,
               const uint8_t *nrec, uint64_t nrec_bytes, const uint32_t *order_N, uint32_t num_N,
               uint32_t num_total, orc_encode_out *out) {
  init_tables();
  memset(out, 0, sizeof(*out));
  int L = max_readlen, W = words_for(L);
  (void)N;
  /* readsingletons, encoder.h:541-570: singleton strings then N strings */
  uint64_t n_sn = n_s + num_N;
  char **sread = (char **)xmalloc(sizeof(char *) * n_sn);
  uint16_t *slen = (uint16_t *)xmalloc(sizeof(uint16_t) * n_sn);
  uint32_t *order_s = (uint32_t *)xmalloc(sizeof(uint32_t) * n_sn);
  for (uint64_t i = 0; i < n_s; i++) {
    uint32_t rid = s_order[i];
    slen[i] = lens[rid];
    sread[i] = (char *)xmalloc((size_t)slen[i] + 1);
    packed_to_string(reads + (size_t)rid * W, slen[i], sread[i]);
    order_s[i] = rid;
  }
  {
    static const char int2dnaN[16] = {'A','G','C','T','N','A','A','A','A','A','A','A','A','A','A','A'}; /* util.cpp:353 */
    uint64_t off = 0;
    for (uint32_t i = 0; i < num_N; i++) {
      if (off + 2 > nrec_bytes) return -2;
      uint16_t len; memcpy(&len, nrec + off, 2); off += 2;
      uint64_t nb = ((uint64_t)len + 1) / 2;
      if (off + nb > nrec_bytes) return -2;
      char *s = (char *)xmalloc((size_t)len + 1);
      for (int j = 0; j < len; j++) s[j] = int2dnaN[(nrec[off + j / 2] >> (4 * (j & 1))) & 15];
      off += nb;
      sread[n_s + i] = s; slen[n_s + i] = len; order_s[n_s + i] = order_N[i];
    }
  }
  /* correct_order: singletons and the aligned stream */
  uint32_t *ord = (uint32_t *)xmalloc(sizeof(uint32_t) * (n_st ? n_st : 1));
  memcpy(ord, st_order, sizeof(uint32_t) * n_st);
  orc_correct_order(order_s, n_s, order_N, num_N, num_total);
  orc_correct_order(ord, n_st, order_N, num_N, num_total);

  /* encoder dictionaries, encoder.h:609-624 */
  orc_dict dict[NUM_DICT]; memset(dict, 0, sizeof(dict));
  if (L > 50) { dict[0].start = 0; dict[0].end = 20; dict[1].start = 21; dict[1].end = 41; }
  else { dict[0].start = 0; dict[0].end = 20 * L / 50; dict[1].start = 20 * L / 50 + 1; dict[1].end = 41 * L / 50; }
  if (n_sn > 0)
    for (int l = 0; l < NUM_DICT; l++) {
      keyrid *kr = (keyrid *)xmalloc(sizeof(keyrid) * n_sn);
      uint32_t n = 0;
      for (uint64_t i = 0; i < n_sn; i++) {
        if (slen[i] <= dict[l].end) continue;
        uint64_t key = 0;
        for (int k = dict[l].start; k <= dict[l].end; k++) key |= (uint64_t)code3(sread[i][k]) << (3 * (k - dict[l].start));
        kr[n].key = key; kr[n].rid = (uint32_t)i; n++;
      }
      dict_from_pairs(&dict[l], kr, n);
      free(kr);
    }
  uint8_t *remaining = (uint8_t *)xmalloc(n_sn ? n_sn : 1);
  memset(remaining, 1, n_sn);
  uint8_t *removed = (uint8_t *)xcalloc(n_sn ? n_sn : 1, 1); /* removed from dicts (lags `remaining` within a probe) */

  /* oriented strings of stream reads are produced lazily per contig */
  uint64_t max_list_size = MAX_LIST_SIZE_DEFAULT;
  { const char *e = getenv("SPRING_ORACLE_MAX_LIST"); if (e && atol(e) > 0) max_list_size = (uint64_t)atol(e); }
  contig_read *list = NULL; uint64_t list_n = 0, list_cap = 0;
  char *arena = NULL; size_t arena_n = 0, arena_cap = 0; /* strings owned by current contig */
  uint64_t seqno = 0, abs_pos = 0;
  char *ref = NULL; size_t ref_cap = 0;
  long (*cnt)[4] = NULL; size_t cnt_cap = 0;
  uint32_t *newly = NULL; size_t newly_cap = 0;
  char *win_rc = (char *)xmalloc((size_t)L + 1);

#define LIST_PUSH(STR, LEN, POS, RC, ORD) do { \
    if (list_n == list_cap) { list_cap = list_cap ? list_cap * 2 : 64; \
      list = (contig_read *)xrealloc(list, list_cap * sizeof(contig_read)); } \
    if (arena_n + (LEN) + 1 > arena_cap) { arena_cap = arena_cap ? arena_cap * 2 : 4096; \
      while (arena_cap < arena_n + (LEN) + 1) { arena_cap *= 2; } \
      arena = (char *)xrealloc(arena, arena_cap); } \
    memcpy(arena + arena_n, (STR), (LEN)); list[list_n].off = arena_n; arena_n += (LEN) + 1; \
    list[list_n].pos = (POS); list[list_n].rc = (RC); list[list_n].order = (ORD); list[list_n].len = (uint16_t)(LEN); \
    list[list_n].seqno = seqno++; list_n++; } while (0)

  char tmp[MAX_READ_LEN + 1], tmp2[MAX_READ_LEN + 1];
  for (uint64_t i = 0; i <= n_st; i++) {
    int done = (i == n_st);
    if (done || st_flag[i] == 0 || list_n > max_list_size) { /* encoder.h:215 */
      if (list_n != 0) {
        qsort(list, list_n, sizeof(contig_read), cmp_contig_read); /* :221 */
        int64_t first_pos = list[0].pos;
        for (uint64_t q = 0; q < list_n; q++) list[q].pos -= first_pos;
        /* buildcontig, encoder.cpp:32-74 (count index A,C,G,T) */
        size_t clen = 0;
        for (uint64_t q = 0; q < list_n; q++) {
          size_t endp = (size_t)list[q].pos + list[q].len;
          if (q == 0) clen = list[q].len; else if (endp > clen) clen = endp;
        }
        if (clen + 1 > ref_cap) { ref_cap = clen * 2 + 64; ref = (char *)xrealloc(ref, ref_cap); }
        if (clen > cnt_cap) { cnt_cap = clen * 2 + 64; cnt = (long(*)[4])xrealloc(cnt, cnt_cap * sizeof(*cnt)); }
        memset(cnt, 0, clen * sizeof(*cnt));
        for (uint64_t q = 0; q < list_n; q++)
          for (int t = 0; t < list[q].len; t++) {
            char ch = arena[list[q].off + t];
            int idx = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 4;
            if (idx < 4) cnt[list[q].pos + t][idx]++;
            /* chartolong['N'] = 4 writes past the 4-entry row in the reference; N never
               reaches buildcontig (only clean reads are in the list at this point) */
          }
        for (size_t t = 0; t < clen; t++) {
          long max = 0; int indmax = 0;
          for (int j = 0; j < 4; j++) if (cnt[t][j] > max) { max = cnt[t][j]; indmax = j; }
          ref[t] = "ACGT"[indmax];
        }
        /* singleton alignment, encoder.h:231-352 */
        if ((int64_t)clen >= L && n_sn > 0) {
          for (long j = 0; j < (long)clen - L + 1; j++) {
            const char *win = ref + j;
            reverse_complement(win, win_rc, L);
            for (int rev = 0; rev < 2; rev++) {
              const char *wb = rev ? win_rc : win;
              for (int l = 0; l < NUM_DICT; l++) {
                uint64_t key = 0;
                for (int k = dict[l].start; k <= dict[l].end; k++) key |= (uint64_t)code3(wb[k]) << (3 * (k - dict[l].start));
                out->enc_probes++;
                int64_t bin = dict[l].numkeys ? dict_lookup(&dict[l], key) : -1;
                if (bin < 0) continue;
                int64_t lo = dict[l].bin_start[bin], hi = dict[l].bin_start[bin + 1];
                int live = 0; size_t n_new = 0;
                for (int64_t q = hi - 1; q >= lo; q--) {
                  uint32_t rid = dict[l].read_id[q];
                  if (removed[rid]) continue;
                  if (++live > MAX_SEARCH) break; /* :270-271 */
                  int h = 0;
                  for (int t = 0; t < slen[rid]; t++) h += __builtin_popcount(code3(wb[t]) ^ code3(sread[rid][t])); /* :274-283 */
                  out->enc_compares++;
                  if (h <= THRESH_ENCODER && remaining[rid]) {
                    remaining[rid] = 0;
                    long pos = rev ? (j + L - slen[rid]) : j; /* :297-299 */
                    const char *rs = sread[rid];
                    if (rev) { reverse_complement(sread[rid], tmp, slen[rid]); rs = tmp; }
                    LIST_PUSH(rs, slen[rid], pos, rev ? 'r' : 'd', order_s[rid]);
                    if (n_new == newly_cap) { newly_cap = newly_cap ? newly_cap * 2 : 64; newly = (uint32_t *)xrealloc(newly, newly_cap * sizeof(uint32_t)); }
                    newly[n_new++] = rid;
                  }
                }
                for (size_t q = 0; q < n_new; q++) removed[newly[q]] = 1; /* :319-333 */
              }
            }
          }
        }
        qsort(list, list_n, sizeof(contig_read), cmp_contig_read); /* :354-356 */
        /* writecontig, encoder.cpp:76-109 */
        bv_push(&out->seq, ref, clen);
        for (uint64_t q = 0; q < list_n; q++) {
          long currentpos = (long)list[q].pos, prevj = 0;
          for (long t = 0; t < list[q].len; t++)
            if (arena[list[q].off + t] != ref[currentpos + t]) {
              bv_push1(&out->noise, (uint8_t)ENC_NOISE[(uint8_t)ref[currentpos + t]][(uint8_t)arena[list[q].off + t]]);
              uint16_t pv = (uint16_t)(t - prevj);
              bv_push(&out->noisepos, &pv, 2);
              prevj = t;
            }
          bv_push1(&out->noise, '\n');
          uint64_t ap = abs_pos + (uint64_t)currentpos;
          bv_push(&out->pos, &ap, 8);
          bv_push(&out->order, &list[q].order, 4);
          bv_push(&out->lengths, &list[q].len, 2);
          bv_push1(&out->rc, list[q].rc);
          out->num_aligned++;
        }
        abs_pos += clen;
      }
      list_n = 0; arena_n = 0;
    }
    if (!done) { /* :360-368: push current record (oriented as writetofile does, reorder.h:667-678) */
      uint32_t rid = st_order[i];
      int len = lens[rid];
      packed_to_string(reads + (size_t)rid * W, len, tmp);
      const char *rs = tmp;
      if (st_rc[i] == 'r') { reverse_complement(tmp, tmp2, len); rs = tmp2; }
      LIST_PUSH(rs, len, st_pos[i], st_rc[i], ord[i]);
    }
  }
  /* unaligned tail, encoder.h:426-453 */
  out->matched_s = (uint32_t)n_s; out->matched_N = num_N;
  for (uint64_t i = 0; i < n_sn; i++)
    if (remaining[i]) {
      if (i < n_s) out->matched_s--; else out->matched_N--;
      bv_push(&out->order, &order_s[i], 4);
      bv_push(&out->lengths, &slen[i], 2);
      write_dnaN_record(&out->unaligned, sread[i], slen[i]);
      out->unaligned_len += slen[i];
    }
  for (uint64_t i = 0; i < n_sn; i++) free(sread[i]);
  free(sread); free(slen); free(order_s); free(ord); free(remaining); free(removed);
  free(list); free(arena); free(ref); free(cnt); free(newly); free(win_rc);
  for (int l = 0; l < NUM_DICT; l++) dict_free(&dict[l]);
  return 0;
#undef LIST_PUSH
}

/* pack_compress_seq's packing, encoder.cpp:126-147: A0 C1 G2 T3, 4 bases/byte LSB first.
 * packed must hold len/4 bytes; the len%4 tail stays ASCII. returns len/4. */
uint64_t orc_pack_seq(const uint8_t *seq, uint64_t len, uint8_t *packed) {
  for (uint64_t i = 0; i < len / 4; i++) {
    uint8_t b = 0;
    for (int j = 0; j < 4; j++) {
      char c = (char)seq[4 * i + j];
      int v = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
      b |= (uint8_t)(v << (2 * j));
    }
    packed[i] = b;
  }
  return len / 4;
}

/* ------------------------------------------------------------------------------------ */
/* inverse: rebuild reads from the encoder-level streams (decompress.cpp:263-283, :310-312) */
/* reads_out: num_reads * stride bytes, read i of the stream written at row i (ASCII).   */
/* ------------------------------------------------------------------------------------ */
int orc_decode(const uint8_t *seq, uint64_t seq_len, const uint64_t *pos, const uint8_t *noise, uint64_t noise_len,
               const uint16_t *noisepos, uint64_t n_noisepos, const uint8_t *rc, uint64_t num_aligned,
               const uint16_t *lengths, uint64_t num_reads, const uint8_t *unaligned, uint64_t unaligned_bytes,
               uint8_t *reads_out, int stride) {
  init_tables();
  static const char int2dnaN[16] = {'A','G','C','T','N','A','A','A','A','A','A','A','A','A','A','A'};
  uint64_t np = 0, nq = 0;
  char tmp[MAX_READ_LEN + 1];
  for (uint64_t i = 0; i < num_aligned; i++) {
    int len = lengths[i];
    if (pos[i] + (uint64_t)len > seq_len || len > stride) return -1;
    char *r = (char *)reads_out + (size_t)i * stride;
    memcpy(r, seq + pos[i], len);
    uint32_t prev = 0;
    while (np < noise_len && noise[np] != '\n') {
      if (nq >= n_noisepos) return -2;
      uint32_t p = prev + noisepos[nq++];
      if ((int)p >= len) return -3;
      r[p] = DEC_NOISE[(uint8_t)r[p]][noise[np]];
      prev = p; np++;
    }
    if (np >= noise_len) return -4;
    np++;
    if (rc[i] == 'r') { memcpy(tmp, r, len); reverse_complement(tmp, r, len); }
  }
  uint64_t off = 0;
  for (uint64_t i = num_aligned; i < num_reads; i++) {
    if (off + 2 > unaligned_bytes) return -5;
    uint16_t len; memcpy(&len, unaligned + off, 2); off += 2;
    if (len != lengths[i] || len > stride) return -6;
    char *r = (char *)reads_out + (size_t)i * stride;
    for (int j = 0; j < len; j++) r[j] = int2dnaN[(unaligned[off + j / 2] >> (4 * (j & 1))) & 15];
    off += ((uint64_t)len + 1) / 2;
  }
  if (np != noise_len || nq != n_noisepos || off != unaligned_bytes) return -7;
  return 0;
}

/* struct sizes for the ctypes side */
int orc_sizeof_reorder_out(void) { return (int)sizeof(orc_reorder_out); }
int orc_sizeof_encode_out(void) { return (int)sizeof(orc_encode_out); }
