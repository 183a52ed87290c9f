// kernels.cuh -- host-callable stages of the pipeline (each implemented in its own .cu).
#pragma once
#include "common.cuh"

namespace sb {

struct Ctx {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  BufferPool pool;
  std::string err;
  uint64_t launches = 0;  // kernels launched since the last reset (ours + CUB passes)
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;  // around the chain kernel
  cudaEvent_t ev_x0 = nullptr, ev_x1 = nullptr;  // around the multi-GPU exchange
  unsigned long long l2_policies[2] = {0, 0};  // createpolicy results (evict_last, evict_first), made once
  bool lockstep = false;  // chain schedule: deterministic round-synchronous, or free-running (default)
  int stitch = -1;  // contig stitching in the encoder: -1 auto (many chains for the input, free-running schedule), 0 off, 1 on
  bool chain_stats = false;  // free-running chains also count lookups / compares (spring_b200_set_chain_stats; the deterministic schedule always does)
};

// ---- dict.cu : constructdictionary (bitset_util.h:74-221) ---------------------------------------
struct DictBuild {
  DictView view{};
  uint32_t capacity = 0;
  // everything below stays on the device (the build never synchronises with the host)
  const uint32_t *d_counts = nullptr;      // [0] indexed reads, [1] unique keys, [2] bins dropped for lack of spare slots
  const uint64_t *sorted_keys = nullptr;   // [n] hashed keys hk = mix64(key) in sorted order, unindexed reads (~0) last
  const uint32_t *bin_start_idx = nullptr; // [unique keys] index of each bin's first entry
  const uint32_t *sorted_rids = nullptr;   // [n] read ids in (hk, id) order
};
// reads: [n][W] 2-bit packed; nflag: optional [n][W] (bit 2j set where base j is N; such reads are
// left out of the dictionary when the N falls inside the window).  tag names the pool buffers.
void build_dictionary(Ctx &c, const uint64_t *reads, const uint16_t *lens, const uint64_t *nflag, uint32_t n, int W,
                      int start, int end, const char *tag, DictBuild &out);

// ---- reorder.cu : reorder<>() (reorder.h:320-641) ------------------------------------------------
struct ReorderDev {
  // aligned stream in chain order (device)
  uint32_t *order = nullptr; uint8_t *flag = nullptr; int64_t *pos = nullptr; uint8_t *rev = nullptr;
  uint64_t num = 0;
  uint32_t *s_order = nullptr; uint64_t num_singletons = 0;
  // stats
  uint32_t num_chains = 0, unmatched = 0;
  uint64_t rounds = 0, lost = 0, probes_issued = 0, probes_seq = 0, compares = 0;
  float ms_kernel = 0;
  uint64_t slot_probes = 0;  // probes that passed the key filter and went to the slot table
  uint64_t cyc[4] = {0, 0, 0, 0};  // SM cycles summed over chains: search, wait A, commit, wait B
};
void run_reorder(Ctx &c, const uint64_t *reads, const uint16_t *lens, uint32_t n, int L, uint32_t num_chains,
                 const DictBuild dict[2], ReorderDev &out);

// ---- encode.cu : encoder_main (encoder.h:572-633) ----------------------------------------------
struct EncodeDev {
  uint8_t *seq_packed = nullptr; uint64_t seq_len = 0;
  uint64_t *pos = nullptr; uint8_t *noise = nullptr; uint64_t noise_bytes = 0;
  uint16_t *noisepos = nullptr; uint64_t num_noise = 0;
  uint8_t *rev = nullptr; uint32_t *order = nullptr; uint16_t *lengths = nullptr;
  uint8_t *unaligned = nullptr; uint64_t unaligned_bytes = 0, unaligned_len = 0;
  uint64_t num_aligned = 0, num_reads = 0;
  uint32_t singletons_aligned = 0, n_reads_aligned = 0;
  uint32_t contigs = 0, contigs_stitched = 0;  // contigs the chains left; how many of them were laid into another contig
};
struct NReads {  // reads with N: input_N.dna records uploaded as they are and unpacked by k_unpack_n
  const uint64_t *codes = nullptr;  // [num][W] 2-bit, N stored as 00
  const uint64_t *nflag = nullptr;  // [num][W] bit 2j set where base j is N
  const uint16_t *lens = nullptr;
  const uint32_t *order = nullptr;  // original index (read_order_N.bin)
  uint32_t num = 0;
};
void run_encode(Ctx &c, const uint64_t *reads, const uint16_t *lens, uint32_t n, int L, const ReorderDev &ro,
                const NReads &nr, uint32_t num_total, EncodeDev &out);

// ---- reblock.cu : pe_encode + the re-blocking of reorder_compress_streams ---------------
enum { RB_FLAG = 0, RB_POS, RB_NOISE, RB_NOISEPOS, RB_RC, RB_UNALIGNED, RB_LENGTHS, RB_POS_PAIR, RB_RC_PAIR, RB_NSTREAMS };
struct ReblockDev {
  uint32_t num_blocks = 0;
  uint8_t *data[RB_NSTREAMS] = {};           // blocks concatenated (device)
  uint64_t size[RB_NSTREAMS] = {};           // bytes
  unsigned long long *block_off = nullptr;   // [RB_NSTREAMS][num_blocks + 1] byte offsets (device)
  uint32_t *order = nullptr;                 // output slot of every stream read, or nullptr (SE -r: identity)
};
// pe_encode.cpp:24-84 on device arrays: slot[i] = position of stream read i in the decompressed output
void run_pe_encode(Ctx &c, const uint32_t *order, uint32_t n, uint32_t *slot);
// reorder_compress_streams.cpp:83-361 on the device-resident encoder streams `e`
void run_reblock(Ctx &c, const EncodeDev &e, bool paired, bool preserve, uint32_t block, ReblockDev &out);

// ---- decode.cu : decompress_short's block decode -----------------------------------
struct DecodeDev {
  uint8_t *bases = nullptr;                 // every read as ASCII, file 1's reads in output order, then file 2's (device)
  unsigned long long *offsets = nullptr;    // [num_reads + 1] (device)
  uint64_t total = 0, num_reads = 0;
};
// b: the nine block streams + block offsets on the device (layout of run_reblock's output); sizes[s]: bytes of stream s
void run_decode_blocks(Ctx &c, const ReblockDev &b, const uint64_t *sizes, const uint8_t *d_seq_packed, uint64_t seq_len,
                       uint64_t num_reads, bool paired, bool preserve, uint32_t block, DecodeDev &out);

// ---- verify.cu : re-block -> block decode -> compare with the input, all in HBM ---------------------------
struct VerifyReport {
  uint64_t num_reads = 0, reads_checked = 0;
  uint64_t base_mismatch_reads = 0, length_mismatch_reads = 0, bad_order = 0;
  uint64_t num_blocks = 0, block_stream_bytes = 0, decoded_bases = 0;
};
void run_verify(Ctx &c, const EncodeDev &e, const uint64_t *reads, const uint16_t *lens, uint32_t num_clean, int W,
                const NReads &nr, bool paired, bool preserve, uint32_t block, VerifyReport &out);

// ---- pack.cu : preprocess's read path, N split + 2-bit / 4-bit packing -----------
struct PackDev {
  uint64_t *reads = nullptr; uint16_t *lengths = nullptr;  // clean reads, input order, [num_clean][W] (device)
  uint8_t *n_records = nullptr; uint64_t n_record_bytes = 0; uint32_t *order_n = nullptr;  // input_N.dna / read_order_N.bin (device)
  uint32_t num_reads = 0, num_clean = 0, num_clean_file1 = 0, num_n = 0, max_readlen = 0;
  int W = 1;
};
// d_bases: the reads' sequence lines concatenated, file 1 then file 2; d_offsets[n + 1]: start of read i
void run_pack_reads(Ctx &c, const uint8_t *d_bases, const unsigned long long *d_offsets, uint32_t n, uint32_t n_file1, PackDev &out);
// input_N.dna records (device) at the given byte offsets -> 2-bit rows + N bit-plane + lengths (readsingletons' N half, encoder.h:556-567)
void run_unpack_n(Ctx &c, const uint8_t *d_records, const unsigned long long *d_offsets, uint32_t nn, int W, uint64_t *codes,
                  uint64_t *nflag, uint16_t *lens);

// ---- exchange.cu : the multi-GPU exchange and the shard finalisation -------------------------------
struct Comm;  // NCCL communicator of one rank (opaque; NCCL is dlopen'ed on first use)
void comm_unique_id(uint8_t id[128]);
Comm *comm_create(const uint8_t id[128], int rank, int world);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_world(const Comm *c);
struct ExchangeDev {  // the reads this rank owns after the all-to-all (device, owned by the context's pool)
  const uint64_t *reads = nullptr; const uint16_t *lens = nullptr; const uint32_t *ids = nullptr;
  uint32_t n = 0;
  uint64_t sent_to_peers = 0, received_from_peers = 0;
};
void run_exchange(Ctx &c, Comm *cm, const uint64_t *reads, const uint16_t *lens, const uint32_t *ids, uint32_t n, int L, ExchangeDev &out);
float exchange_ms(Ctx &c);
struct ShardLayout {  // where this rank's pieces sit in the whole job's streams (all ranks' aligned parts first)
  uint32_t rank = 0, world = 1;
  uint64_t seq_base = 0, aligned_before = 0, noise_before = 0, num_noise_before = 0, unaligned_reads_before = 0, unaligned_bytes_before = 0;
  uint64_t total_seq_len = 0, total_aligned = 0, total_reads = 0, total_noise_bytes = 0, total_num_noise = 0, total_unaligned_bytes = 0,
           total_unaligned_len = 0;
};
// ids[i] = original FASTQ index of clean read k0 + i (d_order_n: the job's read_order_N.bin on the device)
void run_original_ids(Ctx &c, uint32_t k0, uint32_t n, const uint32_t *d_order_n, uint32_t nn, uint32_t *ids);
// pos += sum of the lower ranks' consensus lengths; order[i] = global id (ids: device, the exchanged ids of the owned
// clean reads; h_n_ids: host, global ids of the rank's own reads with N, which were numbered after the owned reads)
void run_finalize_shard(Ctx &c, Comm *cm, EncodeDev &e, const uint32_t *ids, uint32_t n_owned, const uint32_t *h_n_ids, uint32_t n_n,
                        ShardLayout &out);

// ---- bucket.cu : multi-GPU partitioning key --------------------------------------------------
void bucket_reads(Ctx &c, const uint64_t *reads, const uint16_t *lens, uint32_t n, int L, uint32_t num_buckets, uint32_t *bucket);

}  // namespace sb
