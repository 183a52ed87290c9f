/*
 * reblock_oracle.c -- CPU restatement of the two host stages that follow the hot path in
 * `spring -c`: pe_encode (reference src/pe_encode.cpp:24-84) and the re-blocking half of
 * reorder_compress_streams (src/reorder_compress_streams.cpp:83-361; the BSC calls at :363-428 are
 * not part of it).  SURVEY.md section 8(f) rows 1 and 3.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/pyoracle.py): the checker of spring_b200_reblock_streams.
 * Pinned byte for byte against the reference itself (oracle/_ref/spring_ref --reblock) by
 * tests/test_reblock_oracle.py.
 *
 * Everything is kept in memory: the per-block files read_flag.txt.<b>, read_pos.bin.<b>, ... are
 * returned as nine byte streams, blocks concatenated, with per-block byte offsets.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { RB_FLAG = 0, RB_POS, RB_NOISE, RB_NOISEPOS, RB_RC, RB_UNALIGNED, RB_LENGTHS, RB_POS_PAIR, RB_RC_PAIR, RB_NSTREAMS };

typedef struct {
  uint8_t *data[RB_NSTREAMS];
  uint64_t size[RB_NSTREAMS], cap[RB_NSTREAMS];
  uint64_t *off[RB_NSTREAMS]; /* [num_blocks + 1] byte offset of each block in data[] */
  uint32_t num_blocks;
} orc_blocks_out;

size_t orc_sizeof_blocks_out(void) { return sizeof(orc_blocks_out); }

static void put(orc_blocks_out *o, int s, const void *p, uint64_t n) {
  if (o->size[s] + n > o->cap[s]) {
    uint64_t c = o->cap[s] ? o->cap[s] * 2 : 4096;
    while (c < o->size[s] + n) c *= 2;
    o->data[s] = (uint8_t *)realloc(o->data[s], c);
    o->cap[s] = c;
  }
  memcpy(o->data[s] + o->size[s], p, n);
  o->size[s] += n;
}
static void put1(orc_blocks_out *o, int s, uint8_t b) { put(o, s, &b, 1); }

void orc_blocks_free(orc_blocks_out *o) {
  for (int s = 0; s < RB_NSTREAMS; s++) { free(o->data[s]); free(o->off[s]); }
  memset(o, 0, sizeof(*o));
}

/* pe_encode.cpp:24-84: order[i] = original index of stream read i (file-2 reads offset by numreads/2)
 * -> position of stream read i in the decompressed output: file-1 reads keep their stream order,
 * every file-2 read follows its mate. */
int orc_pe_encode(uint32_t *order, uint32_t numreads) {
  const uint32_t half = numreads / 2;
  uint32_t *inverse = (uint32_t *)malloc(sizeof(uint32_t) * (numreads ? numreads : 1));
  if (!inverse) return -1;
  for (uint32_t i = 0; i < numreads; i++) inverse[order[i]] = i; /* :43-48 */
  uint32_t pos_in_file_1 = 0;
  for (uint32_t i = 0; i < numreads; i++)                          /* :53-56 */
    if (order[i] < half) order[i] = pos_in_file_1++;
  /* :60-69.  The loop reads order[] entries it may already have rewritten; a rewritten file-2 entry is
   * >= half as well, and the mate looked up is always a file-1 read (rewritten in the first pass,
   * < half), so the in-place update is well defined.  BUT the test "order[i] >= half" is applied to the
   * value as it is when the loop reaches i -- still the original one for file-2 reads. */
  for (uint32_t i = 0; i < numreads; i++) {
    if (order[i] >= half) {
      const uint32_t mate_original = order[i] - half;
      const uint32_t mate_stream = inverse[mate_original];
      order[i] = order[mate_stream] + half;
    }
  }
  free(inverse);
  return 0;
}

static const char N4[5] = {'A', 'G', 'C', 'T', 'N'}; /* util.cpp:353 (read_dnaN_from_bits) */

/* reorder_compress_streams.cpp:83-361.
 * Streams as the encoder leaves them: pos / rc per aligned read, noise ('\n' terminated per aligned read),
 * noisepos (one u16 per noise symbol), order + lengths for all reads (aligned first), unaligned 4-bit records.
 * order is ignored unless paired_end || preserve_order (:114-115,:139: SE -r keeps the stream order). */
int orc_reblock(const uint64_t *pos, const uint8_t *noise, uint64_t noise_bytes, const uint16_t *noisepos, const uint8_t *rc,
                uint64_t num_aligned, const uint32_t *order, const uint16_t *lengths, uint64_t num_reads,
                const uint8_t *unaligned, uint64_t unaligned_bytes, int paired_end, int preserve_order,
                uint32_t num_reads_per_block, orc_blocks_out *o) {
  memset(o, 0, sizeof(*o));
  const uint64_t n = num_reads, nn = n ? n : 1;
  char *RC_arr = (char *)calloc(nn, 1);
  uint16_t *len_arr = (uint16_t *)calloc(nn, 2);
  uint8_t *flag_arr = (uint8_t *)calloc(nn, 1);
  uint64_t *pos_in_noise = (uint64_t *)calloc(nn, 8), *pos_arr = (uint64_t *)calloc(nn, 8);
  uint16_t *noise_len = (uint16_t *)calloc(nn, 2);
  const int use_order = paired_end || preserve_order;
  /* aligned reads, :112-137 (noise symbols are kept in place: noise_arr == noise without the newlines) */
  uint8_t *noise_arr = (uint8_t *)malloc(noise_bytes ? noise_bytes : 1);
  uint64_t np_noise = 0, rp = 0;
  for (uint64_t i = 0; i < num_aligned; i++) {
    const uint64_t ord = use_order ? order[i] : i;
    RC_arr[ord] = (char)rc[i];
    len_arr[ord] = lengths[i];
    flag_arr[ord] = 1;
    pos_arr[ord] = pos[i];
    pos_in_noise[ord] = np_noise;
    uint16_t k = 0;
    while (rp < noise_bytes && noise[rp] != '\n') { noise_arr[np_noise++] = noise[rp++]; k++; }
    rp++; /* the newline */
    noise_len[ord] = k;
  }
  /* unaligned reads, :141-172: records -> one char array; pos_arr = offset of the read in it */
  const uint64_t num_unaligned = n - num_aligned;
  uint64_t total_unal = 0;
  for (uint64_t i = 0; i < num_unaligned; i++) total_unal += lengths[num_aligned + i];
  char *unal = (char *)malloc(total_unal ? total_unal : 1);
  uint64_t up = 0, wp = 0;
  for (uint64_t i = 0; i < num_unaligned; i++) { /* read_dnaN_from_bits, util.cpp:350-374 */
    if (up + 2 > unaligned_bytes) return -2;
    uint16_t len;
    memcpy(&len, unaligned + up, 2);
    up += 2;
    for (uint16_t j = 0; j < len; j++) unal[wp++] = N4[(unaligned[up + j / 2] >> (4 * (j & 1))) & 15];
    up += ((uint64_t)len + 1) / 2;
  }
  uint64_t cur = 0;
  for (uint64_t i = 0; i < num_unaligned; i++) {
    const uint64_t ord = use_order ? order[num_aligned + i] : num_aligned + i;
    len_arr[ord] = lengths[num_aligned + i];
    pos_arr[ord] = cur;
    cur += lengths[num_aligned + i];
    flag_arr[ord] = 0;
  }
  /* blocks, :201-361 */
  const uint64_t half = n / 2, units = paired_end ? half : n, B = num_reads_per_block;
  const uint32_t nb = (uint32_t)((units + B - 1) / B);
  o->num_blocks = nb;
  for (int s = 0; s < RB_NSTREAMS; s++) o->off[s] = (uint64_t *)calloc((size_t)nb + 1, 8);
  for (uint32_t b = 0; b < nb; b++) {
    for (int s = 0; s < RB_NSTREAMS; s++) o->off[s][b] = o->size[s];
    const uint64_t start = (uint64_t)b * B, end = (start + B < units) ? start + B : units;
    uint64_t prevpos = 0;
#define EMIT_POS_R1(i)                                                                  \
  do {                                                                                  \
    if (preserve_order) put(o, RB_POS, &pos_arr[i], 8);                                 \
    else if ((i) == start) { put(o, RB_POS, &pos_arr[i], 8); prevpos = pos_arr[i]; }     \
    else {                                                                              \
      const uint64_t diff = pos_arr[i] - prevpos;                                       \
      uint16_t d16 = diff < 65535 ? (uint16_t)diff : 65535;                             \
      put(o, RB_POS, &d16, 2);                                                          \
      if (diff >= 65535) put(o, RB_POS, &pos_arr[i], 8);                                \
      prevpos = pos_arr[i];                                                             \
    }                                                                                   \
  } while (0)
#define EMIT_NOISE(i)                                                                   \
  do {                                                                                  \
    put(o, RB_NOISE, noise_arr + pos_in_noise[i], noise_len[i]);                        \
    put(o, RB_NOISEPOS, noisepos + pos_in_noise[i], 2 * (uint64_t)noise_len[i]);        \
    put1(o, RB_NOISE, '\n');                                                            \
  } while (0)
    for (uint64_t i = start; i < end; i++) {
      if (!paired_end) { /* :251-282 */
        put(o, RB_LENGTHS, &len_arr[i], 2);
        if (flag_arr[i]) {
          put1(o, RB_FLAG, '0');
          put1(o, RB_RC, (uint8_t)RC_arr[i]);
          EMIT_POS_R1(i);
          EMIT_NOISE(i);
        } else {
          put1(o, RB_FLAG, '2');
          put(o, RB_UNALIGNED, unal + pos_arr[i], len_arr[i]);
        }
      } else { /* :283-359 */
        const uint64_t ip = half + i;
        put(o, RB_LENGTHS, &len_arr[i], 2);
        put(o, RB_LENGTHS, &len_arr[ip], 2);
        const int64_t pos_pair = (int64_t)pos_arr[ip] - (int64_t)pos_arr[i];
        const int64_t ap = pos_pair < 0 ? -pos_pair : pos_pair; /* std::abs, :288 */
        int flag;
        if (flag_arr[i] && flag_arr[ip] && ap < 32767) flag = 0;
        else if (flag_arr[i] && flag_arr[ip]) flag = 1;
        else if (!flag_arr[i] && !flag_arr[ip]) flag = 2;
        else if (flag_arr[i]) flag = 3;
        else flag = 4;
        put1(o, RB_FLAG, (uint8_t)('0' + flag));
        if (flag == 0) {
          const int16_t pp16 = (int16_t)pos_pair;
          put(o, RB_POS_PAIR, &pp16, 2);
          put1(o, RB_RC_PAIR, RC_arr[i] != RC_arr[ip] ? '0' : '1');
        }
        if (flag == 0 || flag == 1 || flag == 3) {
          EMIT_POS_R1(i);
          EMIT_NOISE(i);
          put1(o, RB_RC, (uint8_t)RC_arr[i]);
        } else {
          put(o, RB_UNALIGNED, unal + pos_arr[i], len_arr[i]);
        }
        if (flag == 0 || flag == 1 || flag == 4) {
          EMIT_NOISE(ip);
          if (flag == 1 || flag == 4) {
            put(o, RB_POS, &pos_arr[ip], 8);
            put1(o, RB_RC, (uint8_t)RC_arr[ip]);
          }
        } else {
          put(o, RB_UNALIGNED, unal + pos_arr[ip], len_arr[ip]);
        }
      }
    }
#undef EMIT_POS_R1
#undef EMIT_NOISE
  }
  for (int s = 0; s < RB_NSTREAMS; s++) o->off[s][nb] = o->size[s];
  free(RC_arr); free(len_arr); free(flag_arr); free(pos_in_noise); free(pos_arr); free(noise_len); free(noise_arr); free(unal);
  return 0;
}

/* ---- the inverse: decompress_short's block decode, reference src/decompress.cpp:230-320 -------------------
 * (SURVEY.md 8f rank 4).  Input: the nine per-block streams in the layout of orc_blocks_out (blocks
 * concatenated + per-block byte offsets) and the consensus as ASCII; output: every read as ASCII, file 1's
 * reads in output order followed by file 2's (paired end), with their lengths.  dec_noise: decompress.cpp:664-685. */
static char DEC[128][128];
static int dec_ready = 0;
static void dec_init(void) {
  if (dec_ready) return;
  memset(DEC, 0, sizeof(DEC));
  DEC['A']['0'] = 'C'; DEC['A']['1'] = 'G'; DEC['A']['2'] = 'T'; DEC['A']['3'] = 'N';
  DEC['C']['0'] = 'A'; DEC['C']['1'] = 'G'; DEC['C']['2'] = 'T'; DEC['C']['3'] = 'N';
  DEC['G']['0'] = 'T'; DEC['G']['1'] = 'A'; DEC['G']['2'] = 'C'; DEC['G']['3'] = 'N';
  DEC['T']['0'] = 'G'; DEC['T']['1'] = 'C'; DEC['T']['2'] = 'A'; DEC['T']['3'] = 'N';
  DEC['N']['0'] = 'A'; DEC['N']['1'] = 'G'; DEC['N']['2'] = 'C'; DEC['N']['3'] = 'T';
  dec_ready = 1;
}
static char rcomp(char c) { /* util.h:23-29 */
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return 'N'; }
}

/* out_reads: capacity >= sum of all lengths; out_len[num_reads]; reads of file 1 (units 0..U-1) first, then file 2.
 * returns 0, or a negative code when a stream runs out. */
int orc_decode_blocks(const uint8_t *const data[RB_NSTREAMS], const uint64_t *const off[RB_NSTREAMS], uint32_t num_blocks,
                      const uint8_t *seq, uint64_t seq_len, uint64_t num_reads, int paired_end, int preserve_order,
                      uint32_t num_reads_per_block, uint8_t *out_reads, uint64_t *out_off, uint16_t *out_len) {
  dec_init();
  const uint64_t units = paired_end ? num_reads / 2 : num_reads, B = num_reads_per_block;
  /* lengths first: they decide where every read lands in out_reads (file 1, then file 2) */
  for (uint32_t b = 0; b < num_blocks; b++) {
    const uint64_t start = (uint64_t)b * B, end = start + B < units ? start + B : units;
    const uint8_t *pl = data[RB_LENGTHS] + off[RB_LENGTHS][b];
    for (uint64_t i = start; i < end; i++) {
      memcpy(&out_len[i], pl, 2); pl += 2;
      if (paired_end) { memcpy(&out_len[units + i], pl, 2); pl += 2; }
    }
  }
  uint64_t acc = 0;
  for (uint64_t i = 0; i < num_reads; i++) { out_off[i] = acc; acc += out_len[i]; }
  out_off[num_reads] = acc;
  for (uint32_t b = 0; b < num_blocks; b++) {
    const uint64_t start = (uint64_t)b * B, end = start + B < units ? start + B : units;
    const uint8_t *pf = data[RB_FLAG] + off[RB_FLAG][b], *pp = data[RB_POS] + off[RB_POS][b];
    const uint8_t *pn = data[RB_NOISE] + off[RB_NOISE][b], *pnp = data[RB_NOISEPOS] + off[RB_NOISEPOS][b];
    const uint8_t *prc = data[RB_RC] + off[RB_RC][b], *pu = data[RB_UNALIGNED] + off[RB_UNALIGNED][b];
    const uint8_t *ppp = data[RB_POS_PAIR] + off[RB_POS_PAIR][b], *prp = data[RB_RC_PAIR] + off[RB_RC_PAIR][b];
    int first_read_of_block = 1;
    uint64_t prevpos = 0;
    for (uint64_t i = start; i < end; i++) {
      const char flag = (char)*pf++;
      uint64_t pos_1 = 0;
      char RC_1 = 'd';
      for (int mate = 0; mate < (paired_end ? 2 : 1); mate++) {
        const uint64_t slot = mate ? units + i : i;
        const uint16_t rl = out_len[slot];
        uint8_t *dst = out_reads + out_off[slot];
        const int singleton = mate == 0 ? (flag == '2' || flag == '4') : (flag == '2' || flag == '3'); /* :234, :286 */
        if (singleton) { memcpy(dst, pu, rl); pu += rl; continue; }                                   /* :275-278 */
        uint64_t pos;
        char RC;
        if (mate == 0) { /* :236-254 */
          if (preserve_order) { memcpy(&pos, pp, 8); pp += 8; }
          else if (first_read_of_block) { first_read_of_block = 0; memcpy(&pos, pp, 8); pp += 8; prevpos = pos; }
          else {
            uint16_t d16; memcpy(&d16, pp, 2); pp += 2;
            if (d16 == 65535) { memcpy(&pos, pp, 8); pp += 8; } else pos = prevpos + d16;
            prevpos = pos;
          }
          RC = (char)*prc++;
          pos_1 = pos; RC_1 = RC;
        } else if (flag == '1' || flag == '4') { /* :290-294 */
          memcpy(&pos, pp, 8); pp += 8;
          RC = (char)*prc++;
        } else { /* :295-305 */
          int16_t pp16; memcpy(&pp16, ppp, 2); ppp += 2;
          pos = pos_1 + (int64_t)pp16;
          const char rel = (char)*prp++;
          RC = rel == '0' ? (RC_1 == 'd' ? 'r' : 'd') : (RC_1 == 'd' ? 'd' : 'r');
        }
        if (pos + rl > seq_len) return -3;
        char tmp[512];
        memcpy(tmp, seq + pos, rl);
        uint16_t prevnp = 0;
        while (*pn != '\n') { /* :258-266 */
          uint16_t np; memcpy(&np, pnp, 2); pnp += 2;
          np = (uint16_t)(np + prevnp);
          if (np >= rl) return -4;
          tmp[np] = DEC[(uint8_t)tmp[np]][*pn];
          prevnp = np;
          pn++;
        }
        pn++;
        if (RC == 'd') memcpy(dst, tmp, rl);
        else for (uint16_t k = 0; k < rl; k++) dst[k] = (uint8_t)rcomp(tmp[rl - 1 - k]); /* util.cpp:376-381 */
      }
    }
  }
  return 0;
}
