/*
 * spring_b200.h -- C ABI of libspring_b200.so: SPRING's short-read reorder + encode hot path
 * on NVIDIA B200 (sm_100a).
 *
 * The reference (shubhamchandak94/Spring) has no plugin/FFI layer; the seam this library fills
 * is the pair of C++ calls
 *     void call_reorder(const std::string &temp_dir, compression_params &cp);
 *     void call_encoder(const std::string &temp_dir, compression_params &cp);
 * (src/call_template_functions.h:9-11, invoked at src/spring.cpp:153 and :166).  Every entry
 * point below cites the reference code it replaces.  Plain pointers and sizes only; no
 * exceptions cross the boundary: every function returns 0 on success or a negative
 * SPRING_B200_E* code, with a message retrievable through spring_b200_last_error().
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * SPRING_B200_ENODEV.
 */
#ifndef SPRING_B200_H_
#define SPRING_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPRING_B200_OK 0
#define SPRING_B200_EINVAL (-1)  /* bad argument (e.g. max_readlen > 511: "Wrong bitset size.", call_template_functions.cpp:61) */
#define SPRING_B200_ENODEV (-2)  /* no usable CUDA device */
#define SPRING_B200_ECUDA (-3)   /* CUDA runtime error */
#define SPRING_B200_EIO (-4)     /* file-level entry point: cannot read/write temp_dir */
#define SPRING_B200_ELIMIT (-5)  /* internal limit hit (watchdog, contig > 10 M reads, ...) */

/* Mirror of spring::compression_params (src/util.h:30-51), the raw 64-byte struct the host
 * writes to cp.bin (src/spring.cpp:218-221).  Offsets checked with static_assert in the library. */
typedef struct spring_b200_cp {
  uint8_t paired_end, preserve_order, preserve_quality, preserve_id;
  uint8_t long_flag, qvz_flag, ill_bin_flag, bin_thr_flag;
  double qvz_ratio;
  uint32_t bin_thr_thr, bin_thr_high, bin_thr_low;
  uint32_t num_reads;
  uint32_t num_reads_clean[2];
  uint32_t max_readlen;
  uint8_t paired_id_code;
  uint8_t paired_id_match;
  int32_t num_reads_per_block;
  int32_t num_reads_per_block_long;
  int32_t num_thr;
} spring_b200_cp;

typedef struct spring_b200_ctx spring_b200_ctx; /* one per GPU; not thread-safe; owns all device + pinned memory */

/* The hot path's inputs, i.e. what preprocess leaves in temp_dir (src/preprocess.cpp:296-403),
 * already in the layout the reference's readDnaFile builds in RAM (src/reorder.h:222-244):
 * one bitset of W = (2*max_readlen-1)/64+1 uint64 words per clean read (2 bits/base, A0 G1 C2 T3,
 * base j at bits 2j..2j+1, zero beyond the read's length) + uint16 lengths. */
typedef struct spring_b200_input {
  const uint64_t *reads;      /* [num_clean * W]; host or device pointer, see each entry point */
  const uint16_t *lengths;    /* [num_clean] */
  uint32_t num_clean;         /* cp.num_reads_clean[0] + cp.num_reads_clean[1] */
  uint32_t max_readlen;       /* cp.max_readlen, 1..511 */
  const uint8_t *n_records;   /* HOST: contents of input_N.dna, 4-bit records (src/util.cpp:322-348) */
  uint64_t n_record_bytes;
  const uint32_t *order_n;    /* HOST: contents of read_order_N.bin (src/preprocess.cpp:300-301,373-378) */
  uint32_t num_n;             /* cp.num_reads - num_clean */
  uint32_t num_reads;         /* cp.num_reads */
} spring_b200_input;

/* The encoder's output streams (what src/encoder.h:386-487 leaves for reorder_compress_streams,
 * src/reorder_compress_streams.cpp:91-172), as flat arrays.  Pointers are owned by the context
 * and stay valid until the next call on it. */
typedef struct spring_b200_streams {
  const uint8_t *seq_packed;   /* consensus, 2 bits/base A0 C1 G2 T3, 4 bases/byte LSB first (src/encoder.cpp:126-141) */
  uint64_t seq_len;            /* bases; seq_packed holds ceil(seq_len/4) bytes (the reference keeps the
                                  last seq_len%4 bases as ASCII in .tail; see spring_b200_write_streams) */
  const uint64_t *pos;         /* read_pos.bin: absolute position per aligned read (src/encoder.h:473-487) */
  const uint8_t *noise;        /* read_noise.txt: substitution codes + '\n' per aligned read (src/encoder.cpp:88-97) */
  uint64_t noise_bytes;
  const uint16_t *noisepos;    /* read_noisepos.bin: delta position per noise symbol */
  uint64_t num_noise;
  const uint8_t *rev;          /* read_rev.txt: 'd' / 'r' per aligned read */
  const uint32_t *order;       /* read_order.bin: original index per read; aligned first, then unaligned */
  const uint16_t *lengths;     /* read_lengths.bin, same order */
  const uint8_t *unaligned;    /* read_unaligned.txt: 4-bit records (src/encoder.h:438-447) */
  uint64_t unaligned_bytes;
  uint64_t unaligned_len;      /* read_unaligned.txt.count: total bases */
  uint64_t num_aligned;
  uint64_t num_reads;
  uint32_t singletons_aligned; /* "N singleton reads were aligned" (src/encoder.h:490-492) */
  uint32_t n_reads_aligned;
} spring_b200_streams;

/* Output of the reorder stage alone (what the reference's reorder threads write to
 * read_order.bin.<t>, tempflag.txt.<t>, temppos.txt.<t>, read_rev.txt.<t> and
 * read_order.bin.singleton, src/reorder.h:498-512, :594-609), chain after chain. */
typedef struct spring_b200_reorder_out {
  const uint32_t *order;   /* clean-read index */
  const uint8_t *flag;     /* 0 = first read of a contig, 1 = matched */
  const int64_t *pos;
  const uint8_t *rev;      /* 'd' / 'r' */
  uint64_t num;
  const uint32_t *singleton_order;
  uint64_t num_singletons;
} spring_b200_reorder_out;

typedef struct spring_b200_stats {
  uint32_t num_chains;       /* chains actually run (reference: cp.num_thr greedy threads) */
  uint32_t unmatched;        /* "Reordering done, X were unmatched" (src/reorder.h:633-635) */
  uint64_t rounds;           /* scheduler rounds of the chain kernel */
  uint64_t lost_proposals;
  uint64_t probes_issued;    /* dictionary lookups the GPU issued (speculative across shifts) */
  uint64_t probes_seq;       /* lookups a sequential search_match would have issued (src/reorder.h:262-273) */
  uint64_t compares;         /* Hamming evaluations counted as the sequential scan would */
  uint64_t gpu_launches;     /* kernels launched by the last call (ours + CUB) */
  float ms_h2d, ms_dict, ms_chains, ms_scatter, ms_encode, ms_d2h, ms_total;
  float ms_chain_kernel;     /* k_chains alone (CUDA events around the cooperative launch) */
  uint64_t cyc_search, cyc_wait_a, cyc_commit, cyc_wait_b; /* SM cycles summed over blocks, per phase of a round */
  uint64_t slot_probes;      /* lookups that passed the L2-resident key filter and read the slot table in HBM */
  float ms_reblock;          /* device time of the last spring_b200_reblock_* call (pe_encode + re-blocking kernels) */
  float ms_exchange;         /* device time of the last spring_b200_exchange_reads (bucket + scatter kernels, NCCL group) */
  uint32_t singletons_aligned; /* "N singleton reads were aligned" / "N reads with N were aligned" (src/encoder.h:490-492) */
  uint32_t n_reads_aligned;
  uint32_t contigs;          /* contigs the chains left in the aligned stream */
  uint32_t contigs_stitched; /* ... of which the encoder laid into another contig (spring_b200_set_stitch) */
} spring_b200_stats;

/* ---- context ---------------------------------------------------------------------------- */
const char *spring_b200_version(void);
int spring_b200_device_count(void);
/* device: CUDA ordinal.  stream: a cudaStream_t to run on (e.g. torch's current stream) or NULL
 * for a stream owned by the context. */
int spring_b200_create(int device, void *stream, spring_b200_ctx **out);
/* Run all later calls on this cudaStream_t.  Unlike create(), a NULL handle here means the legacy
 * default stream (what e.g. torch.cuda.current_stream().cuda_stream is unless a side stream is set). */
int spring_b200_set_stream(spring_b200_ctx *ctx, void *stream);
void spring_b200_destroy(spring_b200_ctx *ctx);
/* The process-wide context of a device, created on first use and kept until the process ends: what the reference-side
 * drop-ins use, so that the stages of one `spring -c` run share CUDA initialisation, buffers and the streams left in
 * HBM (spring_b200_reblock_files then skips reading the stream files back).  Do not destroy it. */
int spring_b200_shared_ctx(int device, spring_b200_ctx **out);
const char *spring_b200_last_error(const spring_b200_ctx *ctx); /* ctx may be NULL: last create() error */
int spring_b200_get_stats(const spring_b200_ctx *ctx, spring_b200_stats *out);
/* Chain schedule of the reorder stage.  deterministic = 0 (default): free-running chains that claim
 * reads with atomic test-and-set, like the reference's threads under their striped locks
 * (src/reorder.h:303-309); fastest, but for more than one chain the output depends on timing, as the
 * reference's does for -t > 1.  deterministic = 1: round-synchronous chains (propose / lowest chain
 * id wins / commit); the output is a pure function of (input, num_chains).  With num_chains = 1 both
 * reproduce the reference's single-thread result. */
int spring_b200_set_schedule(spring_b200_ctx *ctx, int deterministic);
/* Lookup / compare counters of the free-running schedule (spring_b200_stats: probes_issued, probes_seq, compares,
 * slot_probes).  on = 0 (default): the production chain kernel, which leaves the counting out of its hot loops (the
 * four fields read 0); on = 1: the counting instantiation -- what the roofline's algorithmic bytes are computed from
 * (bench.py runs one counted pass outside the timed region).  The deterministic schedule always counts: its counters
 * are part of the parity tests against the oracle's (src/reorder.h:262-311 counted the same way). */
int spring_b200_set_chain_stats(spring_b200_ctx *ctx, int on);
/* Contig stitching in the encoder.  The reference's threads seed every new contig from ONE pool of reads
 * (src/reorder.h:576-592); thousands of GPU chains seed from their own slices, so neighbouring contigs overlap and the
 * overlap's consensus is stored twice.  With stitching, after the first consensus pass every contig's head (its first
 * max_readlen consensus bases) is searched in the other contigs' consensus -- the sweep that re-aligns singletons
 * (src/encoder.h:231-352), with a tighter threshold -- and a contig whose head fits is laid into that contig's coordinates
 * (flipped if it fits reversed) before the consensus is rebuilt.  Only positions / orientations of reads change.
 * mode: -1 (default) automatic = on when the chains are so many for the input that their contig starts would cost more
 * than ~1 % of the read streams (reads < 6400 * chains), free-running schedule only; 0 off (the encoder then equals the
 * reference's encoder on the same reorder stream bit for bit); 1 on. */
int spring_b200_set_stitch(spring_b200_ctx *ctx, int mode);

/* ---- the hot path ------------------------------------------------------------------------ */
/* reorder_main + encoder_main (src/reorder.h:732-786, src/encoder.h:572-633) on HOST buffers:
 * host->device copies, all kernels, device->host copies of the streams.  num_chains = 0 picks
 * the largest co-resident chain count; 1 reproduces the reference's single-thread result. */
int spring_b200_reorder_encode(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains,
                               spring_b200_streams *out);
/* Same with in->reads / in->lengths already resident in HBM (device pointers); the streams are
 * left on the device (out pointers are device pointers); scalar fields are filled. */
int spring_b200_reorder_encode_device(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains,
                                      spring_b200_streams *out);
/* Copy the streams of the last *_device call to pinned host memory. */
int spring_b200_fetch_streams(spring_b200_ctx *ctx, spring_b200_streams *out);

/* ---- stages, for parity tests ------------------------------------------------------------- */
/* constructdictionary (src/bitset_util.h:74-221) for reorder dictionary `which` (0/1): sorted
 * unique keys, CSR bin starts, read ids ascending per bin.  Host inputs, host outputs sized
 * num_clean, num_clean+1, num_clean. */
int spring_b200_build_dictionary(spring_b200_ctx *ctx, const spring_b200_input *in, int which, uint64_t *keys,
                                 uint32_t *bin_start, uint32_t *read_id, uint32_t *num_keys, uint32_t *dict_numreads);
/* reorder<>() alone (src/reorder.h:320-641), host inputs, host outputs owned by the context. */
int spring_b200_reorder(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains,
                        spring_b200_reorder_out *out);

/* The reorder stage's output of the last spring_b200_reorder_encode* call on this context (what the
 * encoder consumed), copied to host memory owned by the context: lets a test feed the very same
 * stream to an independent encoder. */
int spring_b200_fetch_reorder(spring_b200_ctx *ctx, spring_b200_reorder_out *out);

/* ---- the stages after the encoder ----------------------------------------------- */
/* pe_encode (src/pe_encode.cpp:24-84; called at src/spring.cpp:193 for -r paired input): order[i] =
 * original index of stream read i (read_order.bin) -> order_out[i] = its position in the decompressed
 * output (file-1 reads keep the stream order, mates follow at + num_reads/2).  HOST buffers. */
int spring_b200_pe_encode(spring_b200_ctx *ctx, const uint32_t *order, uint32_t num_reads, uint32_t *order_out);

/* The nine per-block streams reorder_compress_streams hands to BSC
 * (src/reorder_compress_streams.cpp:201-361): read_flag.txt.<b>, read_pos.bin.<b>, read_noise.txt.<b>,
 * read_noisepos.bin.<b>, read_rev.txt.<b>, read_unaligned.txt.<b>, read_lengths.bin.<b>,
 * read_pos_pair.bin.<b>, read_rev_pair.txt.<b> (the last two only for paired-end input). */
#define SPRING_B200_NUM_BLOCK_STREAMS 9
enum spring_b200_block_stream {
  SPRING_B200_BS_FLAG = 0, SPRING_B200_BS_POS, SPRING_B200_BS_NOISE, SPRING_B200_BS_NOISEPOS, SPRING_B200_BS_REV,
  SPRING_B200_BS_UNALIGNED, SPRING_B200_BS_LENGTHS, SPRING_B200_BS_POS_PAIR, SPRING_B200_BS_REV_PAIR
};
typedef struct spring_b200_blocks {
  uint32_t num_blocks;                                    /* ceil(units / cp.num_reads_per_block), units = reads or pairs */
  const uint8_t *data[SPRING_B200_NUM_BLOCK_STREAMS];     /* stream s, blocks concatenated; HOST, owned by the context */
  uint64_t size[SPRING_B200_NUM_BLOCK_STREAMS];           /* bytes */
  const uint64_t *off[SPRING_B200_NUM_BLOCK_STREAMS];     /* [num_blocks + 1]: block b = data[s][off[s][b] .. off[s][b+1]) */
  const uint32_t *order;   /* read_order.bin as the stage consumed it (after pe_encode), NULL when it is not used (SE -r) */
  uint64_t num_reads;
} spring_b200_blocks;
/* The re-blocking of reorder_compress_streams (src/reorder_compress_streams.cpp:83-361), preceded by
 * pe_encode when cp->paired_end && !cp->preserve_order (src/spring.cpp:193).  streams: the encoder's
 * output as HOST arrays, or NULL to use the streams the last spring_b200_reorder_encode* call on this
 * context left in HBM (no host round trip).  Uses cp->paired_end, preserve_order, num_reads_per_block. */
int spring_b200_reblock_streams(spring_b200_ctx *ctx, const spring_b200_streams *streams, const spring_b200_cp *cp,
                                spring_b200_blocks *out);
/* File-level drop-in for pe_encode + reorder_compress_streams up to (not including) the BSC calls:
 * reads read_pos.bin, read_noise.txt, read_noisepos.bin, read_rev.txt, read_order.bin, read_lengths.bin,
 * read_unaligned.txt[.count] from temp_dir, deletes them (as :174-181 does) and writes the raw block
 * files <name>.<b>; the host then runs bsc::BSC_compress(<name>.<b>, <name>.<b>.bsc) on each and removes
 * the raw file (:363-428).  See INTEGRATION.md. */
int spring_b200_reblock_files(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp);

/* ---- the decoder's mirror of the re-blocking: decompress_short's block decode ----- */
/* src/decompress.cpp:230-320: the nine per-block streams (HOST, layout of spring_b200_blocks; `order` is
 * not used) + the consensus (read_seq.bin's coding: 2 bits/base A0 C1 G2 T3, 4 bases per byte LSB first,
 * all shards concatenated, src/decompress.cpp:106-120,615-660) -> every read as ASCII.  bases: the reads
 * concatenated without separators, file 1's reads in output order followed by file 2's (paired end);
 * offsets[num_reads + 1].  HOST arrays owned by the context.  Streams that do not belong together
 * (block boundaries, line counts, positions beyond the consensus, noise beyond a read) are refused.
 * The reference decompressor remains the parity oracle of the compressor; this entry point is the GPU
 * counterpart of its inner loop. */
typedef struct spring_b200_decoded {
  const uint8_t *bases;
  const uint64_t *offsets;
  uint64_t num_reads;
} spring_b200_decoded;
int spring_b200_decode_blocks(spring_b200_ctx *ctx, const spring_b200_blocks *blocks, const uint8_t *seq_packed,
                              uint64_t seq_len, const spring_b200_cp *cp, spring_b200_decoded *out);

/* ---- full-size round trip, resident in HBM --------------------------------------------------------------- */
/* The reference's -r check (util/test_script.sh:78-82: compress, decompress, sort, compare) for the streams the
 * last spring_b200_reorder_encode* call left in HBM: re-blocking (src/reorder_compress_streams.cpp:83-361, with
 * pe_encode for paired -r input) -> block decode (src/decompress.cpp:230-320) -> every decoded read compared
 * base by base with the input read that read_order.bin names, and read_order.bin checked to be a permutation.
 * Nothing is copied to the host, so it runs at 100 M reads.  After a *_device call the caller's input arrays
 * must still be alive.  Uses cp->paired_end, preserve_order, num_reads, num_reads_per_block. */
typedef struct spring_b200_verify {
  int32_t ok;                       /* 1: every read decoded to its original */
  uint64_t num_reads, reads_checked;
  uint64_t base_mismatch_reads, length_mismatch_reads, bad_order;
  uint64_t num_blocks, block_stream_bytes, decoded_bases;
} spring_b200_verify;
int spring_b200_verify_roundtrip(spring_b200_ctx *ctx, const spring_b200_cp *cp, spring_b200_verify *out);

/* ---- the stage before the dictionaries: preprocess's read path ------------------ */
/* What preprocess does to the sequence lines of the FASTQ input (src/preprocess.cpp:196-207, :293-304,
 * :364-378; record packers src/util.cpp:269-294, :322-348): reads that contain 'N' go to input_N.dna
 * (4 bits/base) with their original index in read_order_N.bin, all others are packed 2 bits/base in
 * input order -- here straight into the in-memory layout that struct spring_b200_input describes (one bitset row per clean
 * read), so FASTQ bases can go to the hot path without the .dna files.
 *   bases   : HOST, the reads' sequence lines concatenated without separators, file 1 then file 2
 *   offsets : HOST, [num_reads + 1], start of read i in bases
 * Characters other than A, C, G, T, N are refused (the reference's tables are undefined for them);
 * a read longer than 511 fails with the reference's message (preprocess.cpp:199-206).
 * out->reads / out->lengths are DEVICE pointers when keep_on_device != 0 (feed them to
 * spring_b200_reorder_encode_device), HOST pointers otherwise; n_records / order_n are always HOST. */
typedef struct spring_b200_packed_reads {
  const uint64_t *reads;      /* [num_clean * W], W = (2*max_readlen-1)/64+1 */
  const uint16_t *lengths;    /* [num_clean] */
  uint32_t num_clean;         /* cp.num_reads_clean[0] + cp.num_reads_clean[1] */
  uint32_t num_clean_file1;   /* cp.num_reads_clean[0] */
  uint32_t max_readlen;       /* cp.max_readlen */
  const uint8_t *n_records;   /* contents of input_N.dna */
  uint64_t n_record_bytes;
  const uint32_t *order_n;    /* contents of read_order_N.bin */
  uint32_t num_n;
  uint32_t num_reads;
} spring_b200_packed_reads;
int spring_b200_pack_reads(spring_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t num_reads,
                           uint32_t num_reads_file1, int keep_on_device, spring_b200_packed_reads *out);
/* call_reorder + call_encoder's stream generation (as spring_b200_reorder_encode_files: same output files in temp_dir,
 * same use of cp) on the reads the last spring_b200_pack_reads(keep_on_device != 0) of this context left in HBM -- what
 * a preprocess that packs on the GPU hands over instead of input_clean_{1,2}.dna, input_N.dna and read_order_N.bin
 * (src/preprocess.cpp:293-304 -> src/reorder.h:222-244 without the files in between).  cp must describe those reads
 * (num_reads, num_reads_clean[2], max_readlen).  spring_b200_packed_pending: 1 while such reads wait to be consumed. */
int spring_b200_packed_pending(const spring_b200_ctx *ctx);
int spring_b200_reorder_encode_packed(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp, uint32_t num_chains);

/* ---- multi-GPU partitioning ------------------------------------------------------------------ */
/* DEVICE pointers.  bucket[i] = hash(strand-canonical 16-mer minimizer of read i) mod num_buckets:
 * the owner GPU of read i.  No reference counterpart (the reference is single-process,
 * SURVEY.md section 2.4); see DESIGN.md "multi-GPU". */
int spring_b200_bucket_reads(spring_b200_ctx *ctx, const uint64_t *reads, const uint16_t *lengths, uint32_t num_reads,
                             uint32_t max_readlen, uint32_t num_buckets, uint32_t *bucket);

/* ---- multi-GPU: one context (one process or thread) per GPU, ONE exchange, inside the library --------------- */
/* The reference has no distributed path (single process, OpenMP threads sharing one pool of reads; SURVEY.md
 * section 2.4).  Its on-disk format is already sharded -- decompress concatenates read_seq.bin.<t> for t < cp.num_thr
 * and positions are absolute in that concatenation (src/decompress.cpp:106-120, src/encoder.h:473-487) -- so a GPU plays
 * the role of one reference thread:
 *   1. every rank holds a block of the clean reads with their global ids (index among the job's reads, file 2 after file 1)
 *   2. spring_b200_exchange_reads: owner(read) = minimizer bucket mod world; one all-to-all(v) of {row, length, id}
 *      over NCCL / NVLink (fused bucket + histogram kernel, stable scatter into per-destination regions, one group of
 *      ncclSend / ncclRecv); the rank's own reads with N stay where they are
 *   3. spring_b200_reorder_encode_device on the owned reads (N reads numbered after them)
 *   4. spring_b200_finalize_shard: positions made absolute over all ranks' consensus shards, local indices replaced by
 *      global ids -- what the reference's merge of its thread files does (src/encoder.h:386-423, :473-487)
 *   5. the whole job's streams are the ranks' pieces in the order the layout names: all aligned parts in rank order,
 *      then all unaligned parts (the invariant of src/reorder_compress_streams.cpp:254-270); spring_b200_merge_shards
 *      concatenates finalized host streams that way for a single consumer.
 * NCCL is loaded with dlopen when spring_b200_comm_* is first called. */
#define SPRING_B200_COMM_ID_BYTES 128
int spring_b200_comm_unique_id(uint8_t id[SPRING_B200_COMM_ID_BYTES]);   /* on one rank; the host passes it to the others */
int spring_b200_comm_init(spring_b200_ctx *ctx, const uint8_t id[SPRING_B200_COMM_ID_BYTES], int rank, int world);
int spring_b200_comm_free(spring_b200_ctx *ctx);
typedef struct spring_b200_exchanged {
  const uint64_t *reads;     /* DEVICE, [num_reads * W]: the reads this rank owns, sources in rank order, input order inside */
  const uint16_t *lengths;   /* DEVICE */
  const uint32_t *ids;       /* DEVICE: their global ids */
  uint32_t num_reads;
  uint64_t sent_to_peers, received_from_peers;   /* reads that crossed NVLink */
} spring_b200_exchanged;
/* DEVICE pointers in; the outputs stay valid until the next exchange on this context. */
int spring_b200_exchange_reads(spring_b200_ctx *ctx, const uint64_t *reads, const uint16_t *lengths, const uint32_t *ids,
                               uint32_t num_reads, uint32_t max_readlen, spring_b200_exchanged *out);
typedef struct spring_b200_shard_layout {
  uint32_t rank, world;
  uint64_t seq_base;                 /* bases of the lower ranks' consensus shards (added to this rank's positions) */
  uint64_t aligned_before, noise_before, num_noise_before;       /* where this rank's aligned piece starts in each stream */
  uint64_t unaligned_reads_before, unaligned_bytes_before;       /* ... and its unaligned piece, after ALL aligned pieces */
  uint64_t total_seq_len, total_aligned, total_reads, total_noise_bytes, total_num_noise, total_unaligned_bytes, total_unaligned_len;
} spring_b200_shard_layout;
/* After spring_b200_reorder_encode_device on the exchanged reads.  ids: DEVICE, the exchanged ids (num_owned of them);
 * n_ids: HOST, global ids of this rank's reads with N (numbered num_owned, num_owned + 1, ... in the call).  Rewrites
 * the resident streams' pos / order in place (fetch them afterwards) and tells where the shard's pieces go. */
int spring_b200_finalize_shard(spring_b200_ctx *ctx, const uint32_t *ids, uint32_t num_owned, const uint32_t *n_ids, uint32_t num_n,
                               spring_b200_shard_layout *out);
/* Host-side concatenation of finalized shards (HOST streams, rank order) into one job: aligned pieces first, then the
 * unaligned ones, the consensus shards joined at 2 bits per base -- a merged job looks exactly like a single-GPU one
 * (write it with spring_b200_write_streams).  The result owns its memory; release it with spring_b200_free_merged. */
typedef struct spring_b200_merged {
  spring_b200_streams streams;       /* seq_packed: all shards' consensus as one 2-bit stream (shards shifted into place) */
  int num_shards;
  const uint8_t **shard_seq;         /* [num_shards] packed consensus of each shard (pointers into the inputs) */
  uint64_t *shard_seq_len;           /* [num_shards] */
  void *owner;
} spring_b200_merged;
int spring_b200_merge_shards(const spring_b200_streams *shards, int num_shards, spring_b200_merged *out);
void spring_b200_free_merged(spring_b200_merged *m);
/* read_seq.bin.<t> + .tail per shard and the other stream files of a merged job, as spring_b200_write_streams does */
int spring_b200_write_merged(const char *temp_dir, const spring_b200_merged *m);

/* One process driving several GPUs (one host thread and one shared context per GPU): the multi-GPU form of
 * spring_b200_reorder_encode_files below, i.e. of call_reorder + call_encoder (src/call_template_functions.cpp:9-143).
 * Reads the same files from temp_dir, block-distributes the reads over the GPUs, runs exchange -> reorder + encode ->
 * finalisation on every GPU, merges the shards (src/encoder.h:386-487) and writes the same output files, the consensus
 * cut into cp->num_thr pieces as in the single-GPU call -- so cp.bin and every later host stage stay as they are.
 * device_ids may be NULL (devices 0 .. num_gpus-1).  stats (may be NULL): GPU 0's, with the match counters summed.
 * Everything that can be checked is checked before the per-GPU threads start; a CUDA or NCCL failure on one GPU after
 * that point leaves the others waiting in the exchange (as a failed rank does in any NCCL job). */
int spring_b200_reorder_encode_files_multi(const char *temp_dir, const spring_b200_cp *cp, int num_gpus, const int *device_ids,
                                           uint32_t num_chains, spring_b200_stats *stats, char *err, size_t errlen);

/* ---- file-level drop-in -------------------------------------------------------------------- */
/* Replaces call_reorder + call_encoder (src/call_template_functions.cpp:9-143) on a temp_dir:
 * reads input_clean_1.dna [input_clean_2.dna], input_N.dna, read_order_N.bin, deletes them
 * (as src/reorder.h:232,241, src/encoder.h:606, src/encoder.cpp:218 do) and writes
 * read_seq.bin.<t> (2-bit packed, NOT yet BSC-compressed) + read_seq.bin.<t>.tail for
 * t < cp->num_thr, read_pos.bin, read_noise.txt, read_noisepos.bin, read_rev.txt,
 * read_order.bin, read_lengths.bin, read_unaligned.txt, read_unaligned.txt.count.
 * The host then runs bsc::BSC_compress on each read_seq.bin.<t> exactly as
 * pack_compress_seq does (src/encoder.cpp:148-153); see INTEGRATION.md. */
int spring_b200_reorder_encode_files(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp,
                                     uint32_t num_chains);
/* Write a host-side spring_b200_streams into temp_dir in that layout (num_shards = cp.num_thr). */
int spring_b200_write_streams(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_streams *s, int num_shards);

#ifdef __cplusplus
}
#endif
#endif /* SPRING_B200_H_ */
