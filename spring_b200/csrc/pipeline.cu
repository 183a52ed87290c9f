// pipeline.cu -- context, orchestration and the C ABI (include/spring_b200.h).
#include <errno.h>
#include <string.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <cstddef>
#include <fstream>
#include <memory>
#include <mutex>
#include <thread>
#include "../../include/spring_b200.h"
#include "kernels.cuh"

using namespace sb;

// cp.bin layout (reference src/util.h:30-51; offsets in SURVEY.md section 8b)
static_assert(sizeof(spring_b200_cp) == 64, "compression_params is 64 bytes");
static_assert(offsetof(spring_b200_cp, qvz_ratio) == 8, "cp layout");
static_assert(offsetof(spring_b200_cp, num_reads) == 28, "cp layout");
static_assert(offsetof(spring_b200_cp, num_reads_clean) == 32, "cp layout");
static_assert(offsetof(spring_b200_cp, max_readlen) == 40, "cp layout");
static_assert(offsetof(spring_b200_cp, paired_id_code) == 44, "cp layout");
static_assert(offsetof(spring_b200_cp, num_reads_per_block) == 48, "cp layout");
static_assert(offsetof(spring_b200_cp, num_thr) == 56, "cp layout");

struct spring_b200_ctx {
  Ctx c;
  spring_b200_stats stats{};
  EncodeDev last_enc{};
  ReorderDev last_ro{};
  bool have_enc = false;
  // the input of that call, for spring_b200_verify_roundtrip (device pointers; the caller's own for *_device)
  const uint64_t *last_reads = nullptr; const uint16_t *last_lens = nullptr; uint32_t last_num_clean = 0; int last_W = 1;
  NReads last_nr{};
  Comm *comm = nullptr;  // multi-GPU communicator (spring_b200_comm_init)
  std::string files_dir;  // temp_dir whose stream files hold exactly the resident streams (spring_b200_reorder_encode_files)
  // reads left in HBM by spring_b200_pack_reads(keep_on_device != 0) and not consumed yet (spring_b200_reorder_encode_packed)
  bool packed_pending = false;
  spring_b200_packed_reads packed{};
  cudaEvent_t ev[8]{};  // 0-5: the hot path's stages, 6-7: around the re-blocking
};

static thread_local std::string g_create_err;

namespace {

struct IoError : std::runtime_error {
  explicit IoError(const std::string &m) : std::runtime_error(m) {}
};
struct ArgError : std::runtime_error {
  explicit ArgError(const std::string &m) : std::runtime_error(m) {}
};

template <typename F> int guarded(spring_b200_ctx *ctx, F &&f) {
  if (!ctx) return SPRING_B200_EINVAL;
  try {
    SB_CUDA(cudaSetDevice(ctx->c.device));
    f();
    ctx->c.err.clear();
    return SPRING_B200_OK;
  } catch (const ArgError &e) { ctx->c.err = e.what(); return SPRING_B200_EINVAL;
  } catch (const IoError &e) { ctx->c.err = e.what(); return SPRING_B200_EIO;
  } catch (const LimitError &e) { ctx->c.err = e.what(); return SPRING_B200_ELIMIT;
  } catch (const CudaError &e) { ctx->c.err = e.what(); cudaGetLastError(); return SPRING_B200_ECUDA;
  } catch (const std::exception &e) { ctx->c.err = e.what(); return SPRING_B200_ECUDA; }
}

void check_input(const spring_b200_input *in) {
  if (!in) throw ArgError("null input");
  if (in->max_readlen < 1 || in->max_readlen > (uint32_t)kMaxReadLen)
    throw ArgError("Wrong bitset size. (max_readlen must be 1..511)");  // call_template_functions.cpp:61
  if ((uint64_t)in->num_clean + in->num_n != in->num_reads) throw ArgError("num_clean + num_n != num_reads");
  if (in->num_clean && (!in->reads || !in->lengths)) throw ArgError("null reads/lengths");
  if (in->num_n && (!in->n_records || !in->order_n)) throw ArgError("null n_records/order_n");
  // 32-bit slot indices and bins[] offsets (up to 2 n): at most 2^30 reads per GPU shard; bigger jobs are sharded over GPUs
  if (in->num_reads > (1u << 30)) throw ArgError("too many reads for one GPU shard (> 2^30)");
}

// input_N.dna records (util.cpp:322-374) -> 2-bit codes (N as 00) + N bit-plane.  The host only hops over the length
// fields (record i starts where record i - 1 ends) and checks them; the records travel as they are and are unpacked on
// the GPU (k_unpack_n) -- the per-base host loop this replaces was 47 ms of every 100 M-read pass (200 k reads with N).
NReads upload_n_reads(Ctx &c, const spring_b200_input *in, int W) {
  NReads nr{};
  nr.num = in->num_n;
  if (!in->num_n) return nr;
  const uint32_t nn = in->num_n;
  unsigned long long *h_off = c.pool.pin<unsigned long long>("n.h_off", nn);
  uint64_t off = 0;
  for (uint32_t i = 0; i < nn; i++) {
    if (off + 2 > in->n_record_bytes) throw ArgError("input_N.dna truncated");
    uint16_t len; memcpy(&len, in->n_records + off, 2);
    if (len > in->max_readlen) throw ArgError("N read longer than max_readlen");
    const uint64_t nb = ((uint64_t)len + 1) / 2;
    if (off + 2 + nb > in->n_record_bytes) throw ArgError("input_N.dna truncated");
    h_off[i] = off;
    off += 2 + nb;
    if (i && in->order_n[i] <= in->order_n[i - 1]) throw ArgError("read_order_N.bin must be strictly ascending");
  }
  uint8_t *d_rec = c.pool.dev<uint8_t>("n.rec", off + 1);
  unsigned long long *d_off = c.pool.dev<unsigned long long>("n.off", nn);
  uint64_t *d_codes = c.pool.dev<uint64_t>("n.codes", (size_t)nn * W);
  uint64_t *d_flag = c.pool.dev<uint64_t>("n.flag", (size_t)nn * W);
  uint16_t *d_len = c.pool.dev<uint16_t>("n.len", nn);
  uint32_t *d_order = c.pool.dev<uint32_t>("n.order", nn);
  SB_CUDA(cudaMemcpyAsync(d_rec, in->n_records, off, cudaMemcpyHostToDevice, c.stream));
  SB_CUDA(cudaMemcpyAsync(d_off, h_off, sizeof(unsigned long long) * nn, cudaMemcpyHostToDevice, c.stream));
  SB_CUDA(cudaMemcpyAsync(d_order, in->order_n, sizeof(uint32_t) * nn, cudaMemcpyHostToDevice, c.stream));
  run_unpack_n(c, d_rec, d_off, nn, W, d_codes, d_flag, d_len);
  nr.codes = d_codes; nr.nflag = d_flag; nr.lens = d_len; nr.order = d_order;
  return nr;
}

struct DevInput { const uint64_t *reads; const uint16_t *lens; };

DevInput upload_reads(Ctx &c, const spring_b200_input *in, int W) {
  const size_t n = in->num_clean ? in->num_clean : 1;
  uint64_t *d_reads = c.pool.dev<uint64_t>("in.reads", n * W);
  uint16_t *d_lens = c.pool.dev<uint16_t>("in.lens", n);
  if (in->num_clean) {
    SB_CUDA(cudaMemcpyAsync(d_reads, in->reads, sizeof(uint64_t) * (size_t)in->num_clean * W, cudaMemcpyHostToDevice, c.stream));
    SB_CUDA(cudaMemcpyAsync(d_lens, in->lengths, sizeof(uint16_t) * (size_t)in->num_clean, cudaMemcpyHostToDevice, c.stream));
  }
  return {d_reads, d_lens};
}

void rec(spring_b200_ctx *ctx, int i) { SB_CUDA(cudaEventRecord(ctx->ev[i], ctx->c.stream)); }
float ms(spring_b200_ctx *ctx, int i, int j) {
  float t = 0;
  SB_CUDA(cudaEventElapsedTime(&t, ctx->ev[i], ctx->ev[j]));
  return t;
}

// dictionaries + chains (reorder_main, reorder.h:732-786) on device-resident reads
void reorder_on_device(spring_b200_ctx *ctx, DevInput d, const spring_b200_input *in, uint32_t num_chains, ReorderDev &ro) {
  Ctx &c = ctx->c;
  const int L = (int)in->max_readlen, W = words_for(L);
  int s[2], e[2];
  reorder_windows(L, s, e);
  DictBuild dict[2];
  rec(ctx, 1);
  build_dictionary(c, d.reads, d.lens, nullptr, in->num_clean, W, s[0], e[0], "rd.dict0", dict[0]);
  build_dictionary(c, d.reads, d.lens, nullptr, in->num_clean, W, s[1], e[1], "rd.dict1", dict[1]);
  rec(ctx, 2);
  run_reorder(c, d.reads, d.lens, in->num_clean, L, num_chains, dict, ro);
  rec(ctx, 3);
  spring_b200_stats &st = ctx->stats;
  st.num_chains = ro.num_chains; st.unmatched = ro.unmatched; st.rounds = ro.rounds; st.lost_proposals = ro.lost;
  st.probes_issued = ro.probes_issued; st.probes_seq = ro.probes_seq; st.compares = ro.compares;
  st.ms_chain_kernel = ro.ms_kernel; st.slot_probes = ro.slot_probes;
  st.cyc_search = ro.cyc[0]; st.cyc_wait_a = ro.cyc[1]; st.cyc_commit = ro.cyc[2]; st.cyc_wait_b = ro.cyc[3];
}

void fill_scalars(const EncodeDev &e, spring_b200_streams *o) {
  o->seq_len = e.seq_len; o->noise_bytes = e.noise_bytes; o->num_noise = e.num_noise;
  o->unaligned_bytes = e.unaligned_bytes; o->unaligned_len = e.unaligned_len;
  o->num_aligned = e.num_aligned; o->num_reads = e.num_reads;
  o->singletons_aligned = e.singletons_aligned; o->n_reads_aligned = e.n_reads_aligned;
}

void fetch(spring_b200_ctx *ctx, spring_b200_streams *o) {
  Ctx &c = ctx->c;
  const EncodeDev &e = ctx->last_enc;
  fill_scalars(e, o);
  auto get = [&](const char *name, const void *dptr, size_t bytes) -> void * {
    void *h = c.pool.pinned(name, bytes ? bytes : 1);
    if (bytes) SB_CUDA(cudaMemcpyAsync(h, dptr, bytes, cudaMemcpyDeviceToHost, c.stream));
    return h;
  };
  o->seq_packed = (const uint8_t *)get("out.seq", e.seq_packed, (e.seq_len + 3) / 4);
  o->pos = (const uint64_t *)get("out.pos", e.pos, e.num_aligned * 8);
  o->noise = (const uint8_t *)get("out.noise", e.noise, e.noise_bytes);
  o->noisepos = (const uint16_t *)get("out.noisepos", e.noisepos, e.num_noise * 2);
  o->rev = (const uint8_t *)get("out.rev", e.rev, e.num_aligned);
  o->order = (const uint32_t *)get("out.order", e.order, e.num_reads * 4);
  o->lengths = (const uint16_t *)get("out.lengths", e.lengths, e.num_reads * 2);
  o->unaligned = (const uint8_t *)get("out.unaligned", e.unaligned, e.unaligned_bytes);
  SB_CUDA(cudaStreamSynchronize(c.stream));
}

void copy_reorder_out(Ctx &c, const ReorderDev &ro, spring_b200_reorder_out *out) {
  const size_t m = ro.num, s = ro.num_singletons;
  uint32_t *h_order = c.pool.pin<uint32_t>("ro.h_order", m + 1);
  uint8_t *h_flag = c.pool.pin<uint8_t>("ro.h_flag", m + 1);
  int64_t *h_pos = c.pool.pin<int64_t>("ro.h_pos", m + 1);
  uint8_t *h_rev = c.pool.pin<uint8_t>("ro.h_rev", m + 1);
  uint32_t *h_s = c.pool.pin<uint32_t>("ro.h_s", s + 1);
  if (m) {
    SB_CUDA(cudaMemcpyAsync(h_order, ro.order, m * 4, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaMemcpyAsync(h_flag, ro.flag, m, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaMemcpyAsync(h_pos, ro.pos, m * 8, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaMemcpyAsync(h_rev, ro.rev, m, cudaMemcpyDeviceToHost, c.stream));
  }
  if (s) SB_CUDA(cudaMemcpyAsync(h_s, ro.s_order, s * 4, cudaMemcpyDeviceToHost, c.stream));
  out->order = h_order; out->flag = h_flag; out->pos = h_pos; out->rev = h_rev; out->num = m;
  out->singleton_order = h_s; out->num_singletons = s;
}

void run_all(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains, bool host_in, spring_b200_streams *out) {
  check_input(in);
  if (!out) throw ArgError("null output");
  Ctx &c = ctx->c;
  c.launches = 0;
  ctx->stats = spring_b200_stats{};
  ctx->files_dir.clear();
  const int L = (int)in->max_readlen, W = words_for(L);
  rec(ctx, 0);
  DevInput d = host_in ? upload_reads(c, in, W) : DevInput{in->reads, in->lengths};
  NReads nr = upload_n_reads(c, in, W);
  ReorderDev ro;
  reorder_on_device(ctx, d, in, num_chains, ro);
  run_encode(c, d.reads, d.lens, in->num_clean, L, ro, nr, in->num_reads, ctx->last_enc);
  ctx->last_ro = ro;
  ctx->have_enc = true;
  ctx->last_reads = d.reads; ctx->last_lens = d.lens; ctx->last_num_clean = in->num_clean; ctx->last_W = W; ctx->last_nr = nr;
  rec(ctx, 4);
  memset(out, 0, sizeof(*out));
  if (host_in) {
    fetch(ctx, out);
  } else {
    const EncodeDev &e = ctx->last_enc;
    fill_scalars(e, out);
    out->seq_packed = e.seq_packed; out->pos = e.pos; out->noise = e.noise; out->noisepos = e.noisepos;
    out->rev = e.rev; out->order = e.order; out->lengths = e.lengths; out->unaligned = e.unaligned;
  }
  rec(ctx, 5);
  SB_CUDA(cudaStreamSynchronize(c.stream));
  spring_b200_stats &st = ctx->stats;
  st.ms_h2d = ms(ctx, 0, 1); st.ms_dict = ms(ctx, 1, 2); st.ms_chains = ms(ctx, 2, 3); st.ms_scatter = 0;
  st.ms_encode = ms(ctx, 3, 4); st.ms_d2h = ms(ctx, 4, 5); st.ms_total = ms(ctx, 0, 5);
  st.gpu_launches = c.launches;
  st.singletons_aligned = ctx->last_enc.singletons_aligned; st.n_reads_aligned = ctx->last_enc.n_reads_aligned;
  st.contigs = ctx->last_enc.contigs; st.contigs_stitched = ctx->last_enc.contigs_stitched;
}

// ---- files ---------------------------------------------------------------------------------------
std::vector<uint8_t> slurp(const std::string &p, bool must_exist) {
  std::ifstream f(p, std::ios::binary);
  if (!f.is_open()) { if (must_exist) throw IoError("cannot open " + p); return {}; }
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> v((size_t)n);
  if (n) f.read((char *)v.data(), n);
  return v;
}
void spill(const std::string &p, const void *data, size_t n) {
  const int fd = open(p.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) throw IoError("cannot create " + p);
  const char *c = static_cast<const char *>(data);
  size_t at = 0;
  while (at < n) {
    const ssize_t w = write(fd, c + at, n - at > ((size_t)1 << 30) ? ((size_t)1 << 30) : n - at);
    if (w < 0) { if (errno == EINTR) continue; close(fd); throw IoError("write failed: " + p); }
    at += (size_t)w;
  }
  if (close(fd) != 0) throw IoError("write failed: " + p);
}

// A whole file mapped read-only (the page cache is the only copy; the records go straight into pinned memory).
struct MappedFile {
  const uint8_t *p = nullptr; size_t n = 0; int fd = -1;
  MappedFile(const std::string &path, bool must_exist) {
    fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { if (must_exist) throw IoError("cannot open " + path); return; }
    struct stat sb;
    if (fstat(fd, &sb) != 0) { close(fd); fd = -1; throw IoError("cannot stat " + path); }
    n = (size_t)sb.st_size;
    if (n) {
      void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);  // pages are faulted in by the copy threads, in parallel
      if (m == MAP_FAILED) { close(fd); fd = -1; throw IoError("cannot map " + path); }
      p = static_cast<const uint8_t *>(m);
    }
  }
  ~MappedFile() { if (p) munmap(const_cast<uint8_t *>(p), n); if (fd >= 0) close(fd); }
  MappedFile(const MappedFile &) = delete;
  MappedFile &operator=(const MappedFile &) = delete;
};

int host_threads() {
  static const int n = [] {
    if (const char *e = getenv("SPRING_B200_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return v; }
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)(hc ? (hc > 16 ? 16 : hc) : 4);
  }();
  return n;
}
template <typename F> void parallel_ranges(size_t n, F &&f) {  // f(lo, hi) on host_threads() slices of [0, n)
  const int T = n < (1u << 16) ? 1 : host_threads();
  if (T == 1) { f((size_t)0, n); return; }
  std::vector<std::thread> th;
  std::vector<std::string> errs(T);
  for (int t = 0; t < T; t++)
    th.emplace_back([&, t] {
      try { f(n * t / T, n * (t + 1) / T); } catch (const std::exception &e) { errs[t] = e.what(); }
    });
  for (auto &x : th) x.join();
  for (auto &e : errs) if (!e.empty()) throw IoError(e);
}

// readDnaFile (reorder.h:222-244): records {u16 len; ceil(len/4) B} copied into bitset storage.  Fixed-length files
// (the usual case: every record has the same size) are cut into slices copied by several host threads; files with
// mixed lengths need the sequential walk over the 2-byte headers first, then the same parallel copy.
void parse_dna(const MappedFile &buf, uint32_t num, int W, uint32_t max_readlen, uint64_t *reads, uint16_t *lens,
               const std::string &name) {
  if (!num) return;
  if (buf.n < 2) throw IoError(name + " truncated");
  uint16_t len0; memcpy(&len0, buf.p, 2);
  const size_t rec0 = 2 + ((size_t)len0 + 3) / 4;
  bool fixed = len0 <= max_readlen && buf.n == rec0 * (size_t)num;
  if (fixed) {  // every header must say the same length; checked inside the copy
    std::vector<uint8_t> bad(1, 0);
    parallel_ranges(num, [&](size_t lo, size_t hi) {
      const size_t nb = rec0 - 2;
      for (size_t i = lo; i < hi; i++) {
        const uint8_t *r = buf.p + i * rec0;
        uint16_t len; memcpy(&len, r, 2);
        if (len != len0) { bad[0] = 1; return; }
        uint64_t *d = reads + i * (size_t)W;
        for (int w = (int)(nb >> 3); w < W; w++) d[w] = 0;  // the words the record does not fill completely
        memcpy(d, r + 2, nb);
        lens[i] = len;
      }
    });
    if (!bad[0]) return;
    fixed = false;  // same total size by coincidence: fall through to the general walk
  }
  std::vector<size_t> off((size_t)num + 1);
  size_t at = 0;
  for (uint32_t i = 0; i < num; i++) {
    if (at + 2 > buf.n) throw IoError(name + " truncated");
    uint16_t len; memcpy(&len, buf.p + at, 2);
    if (len > max_readlen) throw IoError(name + ": read longer than cp.max_readlen");
    off[i] = at;
    at += 2 + ((size_t)len + 3) / 4;
    if (at > buf.n) throw IoError(name + " truncated");
  }
  off[num] = at;
  parallel_ranges(num, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; i++) {
      const uint8_t *r = buf.p + off[i];
      uint16_t len; memcpy(&len, r, 2);
      uint64_t *d = reads + i * (size_t)W;
      for (int w = 0; w < W; w++) d[w] = 0;
      memcpy(d, r + 2, ((size_t)len + 3) / 4);
      lens[i] = len;
    }
  });
}

// the eight stream files next to read_seq.bin.<t>: every file is created at its final size and filled by pwrite() in
// 64 MiB pieces, the pieces of all files spread over the host threads (read_pos.bin alone is 800 MB at 100 M reads)
void write_stream_files(const std::string &dir, const spring_b200_streams *s) {
  struct Job { const char *name; const void *p; size_t n; int fd; };
  std::vector<Job> jobs = {{"/read_pos.bin", s->pos, (size_t)s->num_aligned * 8, -1}, {"/read_noise.txt", s->noise, (size_t)s->noise_bytes, -1},
                           {"/read_noisepos.bin", s->noisepos, (size_t)s->num_noise * 2, -1}, {"/read_rev.txt", s->rev, (size_t)s->num_aligned, -1},
                           {"/read_order.bin", s->order, (size_t)s->num_reads * 4, -1}, {"/read_lengths.bin", s->lengths, (size_t)s->num_reads * 2, -1},
                           {"/read_unaligned.txt", s->unaligned, (size_t)s->unaligned_bytes, -1},
                           {"/read_unaligned.txt.count", &s->unaligned_len, (size_t)8, -1}};
  constexpr size_t kPiece = (size_t)64 << 20;
  struct Piece { int job; size_t off, n; };
  std::vector<Piece> pieces;
  auto close_all = [&] { for (auto &j : jobs) if (j.fd >= 0) { close(j.fd); j.fd = -1; } };
  for (size_t k = 0; k < jobs.size(); k++) {
    Job &j = jobs[k];
    j.fd = open((dir + j.name).c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (j.fd < 0) { close_all(); throw IoError("cannot create " + dir + j.name); }
    for (size_t off = 0; off < j.n; off += kPiece) pieces.push_back({(int)k, off, std::min(kPiece, j.n - off)});
  }
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  const int T = (int)std::min<size_t>((size_t)host_threads(), pieces.size() ? pieces.size() : 1);
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++)
    th.emplace_back([&] {
      for (size_t i = next++; i < pieces.size(); i = next++) {
        const Piece &pc = pieces[i];
        const char *src = static_cast<const char *>(jobs[pc.job].p) + pc.off;
        size_t at = 0;
        while (at < pc.n) {
          const ssize_t w = pwrite(jobs[pc.job].fd, src + at, pc.n - at, (off_t)(pc.off + at));
          if (w < 0) { if (errno == EINTR) continue; failed = 1; return; }
          at += (size_t)w;
        }
      }
    });
  for (auto &t : th) t.join();
  bool bad = failed.load() != 0;
  for (auto &j : jobs) if (j.fd >= 0) { if (close(j.fd) != 0) bad = true; j.fd = -1; }
  if (bad) throw IoError("write failed in " + dir);
}

void write_streams(const std::string &dir, const spring_b200_streams *s, int num_shards) {
  if (num_shards < 1) throw ArgError("num_shards < 1");
  // read_seq.bin.<t> + .tail (encoder.cpp:111-156, decompress.cpp:106-120): the consensus is cut
  // into num_shards pieces at multiples of 4 bases; only the last piece can have a tail
  static const char code2char[4] = {'A', 'C', 'G', 'T'};
  const uint64_t full = s->seq_len / 4;
  for (int t = 0; t < num_shards; t++) {
    const uint64_t b0 = full * t / num_shards, b1 = full * (t + 1) / num_shards;
    const std::string base = dir + "/read_seq.bin." + std::to_string(t);
    spill(base, s->seq_packed + b0, b1 - b0);
    std::string tail;
    if (t == num_shards - 1)
      for (uint64_t x = full * 4; x < s->seq_len; x++) tail.push_back(code2char[(s->seq_packed[x / 4] >> (2 * (x & 3))) & 3]);
    spill(base + ".tail", tail.data(), tail.size());
  }
  write_stream_files(dir, s);
}

// ---- pe_encode / re-blocking -----------------------------------------------------------------
// host copies of the encoder's streams -> device buffers laid out like run_encode's output
EncodeDev upload_streams(Ctx &c, const spring_b200_streams *s) {
  if (s->num_aligned > s->num_reads) throw ArgError("streams: num_aligned > num_reads");
  if (s->num_reads >= 0x7FFFFFF0ull) throw ArgError("too many reads for one GPU shard (>= 2^31)");
  if (s->noise_bytes != s->num_noise + s->num_aligned) throw ArgError("streams: noise_bytes != num_noise + num_aligned");
  EncodeDev e{};
  auto up = [&](const char *name, const void *h, size_t bytes) -> void * {
    void *d = c.pool.device(name, bytes + 8);
    if (bytes) {
      if (!h) throw ArgError(std::string("streams: null pointer for ") + name);
      SB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c.stream));
    }
    return d;
  };
  e.pos = (uint64_t *)up("rbin.pos", s->pos, s->num_aligned * 8);
  e.noise = (uint8_t *)up("rbin.noise", s->noise, s->noise_bytes);
  e.noisepos = (uint16_t *)up("rbin.noisepos", s->noisepos, s->num_noise * 2);
  e.rev = (uint8_t *)up("rbin.rev", s->rev, s->num_aligned);
  e.order = (uint32_t *)up("rbin.order", s->order, s->num_reads * 4);
  e.lengths = (uint16_t *)up("rbin.lengths", s->lengths, s->num_reads * 2);
  e.unaligned = (uint8_t *)up("rbin.unaligned", s->unaligned, s->unaligned_bytes);
  e.noise_bytes = s->noise_bytes; e.num_noise = s->num_noise; e.unaligned_bytes = s->unaligned_bytes;
  e.unaligned_len = s->unaligned_len; e.num_aligned = s->num_aligned; e.num_reads = s->num_reads;
  return e;
}

void reblock(spring_b200_ctx *ctx, const spring_b200_streams *streams, const spring_b200_cp *cp, spring_b200_blocks *out) {
  if (!cp || !out) throw ArgError("null argument");
  if (cp->long_flag) throw ArgError("long mode has no reorder_compress_streams stage (spring.cpp:150)");
  if (cp->num_reads_per_block <= 0) throw ArgError("cp.num_reads_per_block <= 0");
  Ctx &c = ctx->c;
  c.launches = 0;
  EncodeDev e;
  if (streams) e = upload_streams(c, streams);
  else {
    if (!ctx->have_enc) throw ArgError("no streams resident on the device (run spring_b200_reorder_encode* first, or pass host streams)");
    e = ctx->last_enc;
  }
  if (e.num_reads != cp->num_reads) throw ArgError("streams hold a different number of reads than cp.num_reads");
  rec(ctx, 6);
  ReblockDev rb;
  run_reblock(c, e, cp->paired_end != 0, cp->preserve_order != 0, (uint32_t)cp->num_reads_per_block, rb);
  rec(ctx, 7);
  memset(out, 0, sizeof(*out));
  out->num_blocks = rb.num_blocks;
  out->num_reads = e.num_reads;
  static const char *names[RB_NSTREAMS] = {"rbo.flag", "rbo.pos", "rbo.noise", "rbo.noisepos", "rbo.rc", "rbo.unal", "rbo.len", "rbo.pos_pair", "rbo.rc_pair"};
  for (int s = 0; s < RB_NSTREAMS; s++) {
    uint8_t *h = c.pool.pin<uint8_t>(names[s], rb.size[s] + 1);
    if (rb.size[s]) SB_CUDA(cudaMemcpyAsync(h, rb.data[s], rb.size[s], cudaMemcpyDeviceToHost, c.stream));
    out->data[s] = h; out->size[s] = rb.size[s];
  }
  const size_t st = (size_t)rb.num_blocks + 1;
  uint64_t *h_off = c.pool.pin<uint64_t>("rbo.off", RB_NSTREAMS * st);
  SB_CUDA(cudaMemcpyAsync(h_off, rb.block_off, sizeof(uint64_t) * RB_NSTREAMS * st, cudaMemcpyDeviceToHost, c.stream));
  for (int s = 0; s < RB_NSTREAMS; s++) out->off[s] = h_off + s * st;
  if (rb.order) {
    uint32_t *h_order = c.pool.pin<uint32_t>("rbo.order", e.num_reads + 1);
    SB_CUDA(cudaMemcpyAsync(h_order, rb.order, sizeof(uint32_t) * e.num_reads, cudaMemcpyDeviceToHost, c.stream));
    out->order = h_order;
  }
  SB_CUDA(cudaStreamSynchronize(c.stream));
  ctx->stats.ms_reblock = ms(ctx, 6, 7);
  ctx->stats.gpu_launches = c.launches;
}

}  // namespace

extern "C" {

const char *spring_b200_version(void) { return "spring_b200 0.1.0 (sm_100a)"; }

int spring_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int spring_b200_create(int device, void *stream, spring_b200_ctx **out) {
  if (!out) return SPRING_B200_EINVAL;
  *out = nullptr;
  int n = spring_b200_device_count();
  if (n <= 0 || device < 0 || device >= n) {
    g_create_err = "no usable CUDA device (spring_b200 has no CPU fallback)";
    return SPRING_B200_ENODEV;
  }
  spring_b200_ctx *ctx = nullptr;
  try {
    SB_CUDA(cudaSetDevice(device));
    ctx = new spring_b200_ctx();
    ctx->c.device = device;
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, device));
    ctx->c.num_sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) throw CudaError("device lacks cooperative launch");
    if (stream) { ctx->c.stream = (cudaStream_t)stream; ctx->c.own_stream = false; }
    else { SB_CUDA(cudaStreamCreateWithFlags(&ctx->c.stream, cudaStreamNonBlocking)); ctx->c.own_stream = true; }
    for (auto &e : ctx->ev) SB_CUDA(cudaEventCreate(&e));
    if (const char *e = getenv("SPRING_B200_CHAIN_STATS")) ctx->c.chain_stats = atoi(e) != 0;  // tools: counters without an API call
    if (const char *e = getenv("SPRING_B200_STITCH")) { const int m = atoi(e); if (m >= -1 && m <= 1) ctx->c.stitch = m; }
  } catch (const std::exception &e) {
    g_create_err = e.what();
    delete ctx;
    return SPRING_B200_ECUDA;
  }
  *out = ctx;
  return SPRING_B200_OK;
}

// Process-wide context of a device for the reference-side drop-ins (call_reorder, call_encoder, reorder_compress_streams
// are separate calls of one `spring -c` run): created on first use and kept, so that CUDA initialisation, the buffer pool
// and the streams left in HBM are shared between the stages.  Never destroyed explicitly (process exit releases it).
int spring_b200_shared_ctx(int device, spring_b200_ctx **out) {
  static std::map<int, spring_b200_ctx *> shared;
  static std::mutex mu;
  if (!out) return SPRING_B200_EINVAL;
  std::lock_guard<std::mutex> lock(mu);
  auto it = shared.find(device);
  if (it != shared.end()) { *out = it->second; return SPRING_B200_OK; }
  const int rc = spring_b200_create(device, nullptr, out);
  if (rc == SPRING_B200_OK) shared[device] = *out;
  return rc;
}

void spring_b200_destroy(spring_b200_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  ctx->c.pool.release();
  if (ctx->comm) { try { comm_destroy(ctx->comm); } catch (...) {} ctx->comm = nullptr; }
  if (ctx->c.ev_x0) { cudaEventDestroy(ctx->c.ev_x0); cudaEventDestroy(ctx->c.ev_x1); }
  if (ctx->c.ev_k0) { cudaEventDestroy(ctx->c.ev_k0); cudaEventDestroy(ctx->c.ev_k1); }
  for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
  if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
  delete ctx;
}

const char *spring_b200_last_error(const spring_b200_ctx *ctx) { return ctx ? ctx->c.err.c_str() : g_create_err.c_str(); }

int spring_b200_get_stats(const spring_b200_ctx *ctx, spring_b200_stats *out) {
  if (!ctx || !out) return SPRING_B200_EINVAL;
  *out = ctx->stats;
  return SPRING_B200_OK;
}

int spring_b200_set_stream(spring_b200_ctx *ctx, void *stream) {
  if (!ctx) return SPRING_B200_EINVAL;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  if (ctx->c.own_stream) cudaStreamDestroy(ctx->c.stream);
  ctx->c.stream = (cudaStream_t)stream;
  ctx->c.own_stream = false;
  return SPRING_B200_OK;
}

int spring_b200_set_schedule(spring_b200_ctx *ctx, int deterministic) {
  if (!ctx) return SPRING_B200_EINVAL;
  ctx->c.lockstep = deterministic != 0;
  return SPRING_B200_OK;
}

int spring_b200_set_chain_stats(spring_b200_ctx *ctx, int on) {
  if (!ctx) return SPRING_B200_EINVAL;
  ctx->c.chain_stats = on != 0;
  return SPRING_B200_OK;
}

int spring_b200_set_stitch(spring_b200_ctx *ctx, int mode) {
  if (!ctx || mode < -1 || mode > 1) return SPRING_B200_EINVAL;
  ctx->c.stitch = mode;
  return SPRING_B200_OK;
}

int spring_b200_reorder_encode(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains, spring_b200_streams *out) {
  return guarded(ctx, [&] { run_all(ctx, in, num_chains, true, out); });
}

int spring_b200_reorder_encode_device(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains,
                                      spring_b200_streams *out) {
  return guarded(ctx, [&] { run_all(ctx, in, num_chains, false, out); });
}

int spring_b200_fetch_streams(spring_b200_ctx *ctx, spring_b200_streams *out) {
  return guarded(ctx, [&] {
    if (!ctx->have_enc || !out) throw ArgError("no streams to fetch");
    fetch(ctx, out);
  });
}

int spring_b200_build_dictionary(spring_b200_ctx *ctx, const spring_b200_input *in, int which, uint64_t *keys,
                                 uint32_t *bin_start, uint32_t *read_id, uint32_t *num_keys, uint32_t *dict_numreads) {
  return guarded(ctx, [&] {
    check_input(in);
    if (which < 0 || which > 1 || !keys || !bin_start || !read_id || !num_keys || !dict_numreads) throw ArgError("bad argument");
    Ctx &c = ctx->c;
    c.launches = 0;
    const int L = (int)in->max_readlen, W = words_for(L);
    DevInput d = upload_reads(c, in, W);
    int s[2], e[2];
    reorder_windows(L, s, e);
    DictBuild db;
    build_dictionary(c, d.reads, d.lens, nullptr, in->num_clean, W, s[which], e[which], "rd.dict0", db);
    // The build keeps its bins in hashed-key order; the reference's CSR is in key order (bitset_util.h:118-131):
    // recover every bin's key from its first read and sort the bins by it (test entry point: host loop).
    uint32_t cnts[4];
    SB_CUDA(cudaMemcpyAsync(cnts, db.d_counts, sizeof(cnts), cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    const uint32_t nv = cnts[0], nk = cnts[1];
    if (cnts[2]) throw LimitError("dictionary: bins dropped (probe chain longer than the spare slots)");
    std::vector<uint32_t> bsi(nk + 1), rids(nv);
    if (nv) SB_CUDA(cudaMemcpyAsync(rids.data(), db.sorted_rids, sizeof(uint32_t) * nv, cudaMemcpyDeviceToHost, c.stream));
    if (nk) SB_CUDA(cudaMemcpyAsync(bsi.data(), db.bin_start_idx, sizeof(uint32_t) * nk, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    bsi[nk] = nv;
    const int kstart = s[which], kend = e[which], nbits = 2 * (kend - kstart + 1);
    auto key_of = [&](uint32_t rid) {
      const uint64_t *r = in->reads + (size_t)rid * W;
      const int pos = 2 * kstart, k = pos >> 6, bs = pos & 63;
      uint64_t v = r[k] >> bs;
      if (bs && k + 1 < W) v |= r[k + 1] << (64 - bs);
      return nbits < 64 ? v & ((1ull << nbits) - 1ull) : v;
    };
    std::vector<std::pair<uint64_t, uint32_t>> order(nk);
    for (uint32_t k = 0; k < nk; k++) order[k] = {key_of(rids[bsi[k]]), k};
    std::sort(order.begin(), order.end());
    uint32_t at = 0;
    for (uint32_t j = 0; j < nk; j++) {
      const uint32_t k = order[j].second;
      keys[j] = order[j].first;
      bin_start[j] = at;
      for (uint32_t i = bsi[k]; i < bsi[k + 1]; i++) read_id[at++] = rids[i];
    }
    bin_start[nk] = at;
    *num_keys = nk;
    *dict_numreads = nv;
    ctx->stats = spring_b200_stats{};
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_reorder(spring_b200_ctx *ctx, const spring_b200_input *in, uint32_t num_chains, spring_b200_reorder_out *out) {
  return guarded(ctx, [&] {
    check_input(in);
    if (!out) throw ArgError("null output");
    Ctx &c = ctx->c;
    c.launches = 0;
    ctx->stats = spring_b200_stats{};
    const int W = words_for((int)in->max_readlen);
    rec(ctx, 0);
    DevInput d = upload_reads(c, in, W);
    ReorderDev ro;
    reorder_on_device(ctx, d, in, num_chains, ro);
    copy_reorder_out(c, ro, out);
    rec(ctx, 4);
    SB_CUDA(cudaStreamSynchronize(c.stream));
    ctx->stats.ms_h2d = ms(ctx, 0, 1); ctx->stats.ms_dict = ms(ctx, 1, 2); ctx->stats.ms_chains = ms(ctx, 2, 3);
    ctx->stats.ms_total = ms(ctx, 0, 4);
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_fetch_reorder(spring_b200_ctx *ctx, spring_b200_reorder_out *out) {
  return guarded(ctx, [&] {
    if (!ctx->have_enc || !out) throw ArgError("no reorder output to fetch");
    copy_reorder_out(ctx->c, ctx->last_ro, out);
    SB_CUDA(cudaStreamSynchronize(ctx->c.stream));
  });
}


int spring_b200_decode_blocks(spring_b200_ctx *ctx, const spring_b200_blocks *blocks, const uint8_t *seq_packed,
                              uint64_t seq_len, const spring_b200_cp *cp, spring_b200_decoded *out) {
  return guarded(ctx, [&] {
    if (!blocks || !cp || !out || (seq_len && !seq_packed)) throw ArgError("null argument");
    if (cp->long_flag) throw ArgError("long mode archives hold no block streams of this kind (spring.cpp:150)");
    if (cp->num_reads_per_block <= 0) throw ArgError("cp.num_reads_per_block <= 0");
    if (cp->num_reads >= 0x7FFFFFF0u) throw ArgError("too many reads for one GPU shard (>= 2^31)");
    Ctx &c = ctx->c;
    c.launches = 0;
    static const char *names[RB_NSTREAMS] = {"dcin.flag", "dcin.pos", "dcin.noise", "dcin.noisepos", "dcin.rc", "dcin.unal", "dcin.len",
                                             "dcin.pos_pair", "dcin.rc_pair"};
    ReblockDev rb;
    rb.num_blocks = blocks->num_blocks;
    const size_t st = (size_t)blocks->num_blocks + 1;
    std::vector<unsigned long long> h_off(RB_NSTREAMS * st);
    for (int s = 0; s < RB_NSTREAMS; s++) {
      if (!blocks->off[s]) throw ArgError("blocks: null offsets");
      for (size_t b = 0; b < st; b++) {
        h_off[s * st + b] = blocks->off[s][b];
        if ((b && blocks->off[s][b] < blocks->off[s][b - 1]) || blocks->off[s][b] > blocks->size[s]) throw ArgError("blocks: offsets out of order");
      }
      if (blocks->off[s][0] != 0 || blocks->off[s][st - 1] != blocks->size[s]) throw ArgError("blocks: offsets do not span the stream");
      rb.size[s] = blocks->size[s];
      rb.data[s] = c.pool.dev<uint8_t>(names[s], blocks->size[s] + 16);
      if (blocks->size[s]) {
        if (!blocks->data[s]) throw ArgError("blocks: null stream");
        SB_CUDA(cudaMemcpyAsync(rb.data[s], blocks->data[s], blocks->size[s], cudaMemcpyHostToDevice, c.stream));
      }
    }
    rb.block_off = c.pool.dev<unsigned long long>("dcin.off", RB_NSTREAMS * st);
    SB_CUDA(cudaMemcpyAsync(rb.block_off, h_off.data(), sizeof(unsigned long long) * RB_NSTREAMS * st, cudaMemcpyHostToDevice, c.stream));
    uint8_t *d_seq = c.pool.dev<uint8_t>("dcin.seq", (seq_len + 3) / 4 + 16);
    if (seq_len) SB_CUDA(cudaMemcpyAsync(d_seq, seq_packed, (seq_len + 3) / 4, cudaMemcpyHostToDevice, c.stream));
    DecodeDev dd;
    try {
      run_decode_blocks(c, rb, rb.size, d_seq, seq_len, cp->num_reads, cp->paired_end != 0, cp->preserve_order != 0,
                        (uint32_t)cp->num_reads_per_block, dd);
    } catch (const LimitError &e) { throw ArgError(e.what()); }  // bad input, not an internal limit
    uint8_t *h_bases = c.pool.pin<uint8_t>("dc.h_bases", dd.total + 1);
    uint64_t *h_offs = c.pool.pin<uint64_t>("dc.h_offsets", dd.num_reads + 1);
    if (dd.total) SB_CUDA(cudaMemcpyAsync(h_bases, dd.bases, dd.total, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaMemcpyAsync(h_offs, dd.offsets, sizeof(uint64_t) * (dd.num_reads + 1), cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    out->bases = h_bases; out->offsets = h_offs; out->num_reads = dd.num_reads;
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_pack_reads(spring_b200_ctx *ctx, const uint8_t *bases, const uint64_t *offsets, uint32_t num_reads,
                           uint32_t num_reads_file1, int keep_on_device, spring_b200_packed_reads *out) {
  return guarded(ctx, [&] {
    if (!out || !offsets || (num_reads && !bases && offsets[num_reads])) throw ArgError("null pointer");
    if (num_reads_file1 > num_reads) throw ArgError("num_reads_file1 > num_reads");
    if (num_reads >= 0x7FFFFFF0u) throw ArgError("too many reads for one GPU shard (>= 2^31)");
    Ctx &c = ctx->c;
    c.launches = 0;
    const uint64_t nbytes = offsets[num_reads];
    uint8_t *d_bases = c.pool.dev<uint8_t>("pk.bases", nbytes + 1);
    unsigned long long *d_off = c.pool.dev<unsigned long long>("pk.offsets", (size_t)num_reads + 1);
    if (nbytes) SB_CUDA(cudaMemcpyAsync(d_bases, bases, nbytes, cudaMemcpyHostToDevice, c.stream));
    SB_CUDA(cudaMemcpyAsync(d_off, offsets, sizeof(uint64_t) * ((size_t)num_reads + 1), cudaMemcpyHostToDevice, c.stream));
    PackDev pk;
    try { run_pack_reads(c, d_bases, d_off, num_reads, num_reads_file1, pk); }
    catch (const LimitError &e) { throw ArgError(e.what()); }  // bad input, not an internal limit
    memset(out, 0, sizeof(*out));
    out->num_clean = pk.num_clean; out->num_clean_file1 = pk.num_clean_file1; out->max_readlen = pk.max_readlen;
    out->n_record_bytes = pk.n_record_bytes; out->num_n = pk.num_n; out->num_reads = pk.num_reads;
    uint8_t *h_nrec = c.pool.pin<uint8_t>("pk.h_nrec", pk.n_record_bytes + 1);
    uint32_t *h_on = c.pool.pin<uint32_t>("pk.h_order_n", (size_t)pk.num_n + 1);
    if (pk.n_record_bytes) SB_CUDA(cudaMemcpyAsync(h_nrec, pk.n_records, pk.n_record_bytes, cudaMemcpyDeviceToHost, c.stream));
    if (pk.num_n) SB_CUDA(cudaMemcpyAsync(h_on, pk.order_n, sizeof(uint32_t) * pk.num_n, cudaMemcpyDeviceToHost, c.stream));
    out->n_records = h_nrec; out->order_n = h_on;
    ctx->packed_pending = false;
    if (keep_on_device) { out->reads = pk.reads; out->lengths = pk.lengths; ctx->packed = *out; ctx->packed_pending = true; }
    else {
      uint64_t *h_reads = c.pool.pin<uint64_t>("pk.h_reads", (size_t)pk.num_clean * pk.W + 1);
      uint16_t *h_len = c.pool.pin<uint16_t>("pk.h_len", (size_t)pk.num_clean + 1);
      if (pk.num_clean) {
        SB_CUDA(cudaMemcpyAsync(h_reads, pk.reads, sizeof(uint64_t) * (size_t)pk.num_clean * pk.W, cudaMemcpyDeviceToHost, c.stream));
        SB_CUDA(cudaMemcpyAsync(h_len, pk.lengths, sizeof(uint16_t) * pk.num_clean, cudaMemcpyDeviceToHost, c.stream));
      }
      out->reads = h_reads; out->lengths = h_len;
    }
    SB_CUDA(cudaStreamSynchronize(c.stream));
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_pe_encode(spring_b200_ctx *ctx, const uint32_t *order, uint32_t num_reads, uint32_t *order_out) {
  return guarded(ctx, [&] {
    if (num_reads && (!order || !order_out)) throw ArgError("null pointer");
    if (num_reads & 1) throw ArgError("pe_encode: odd number of reads");
    if (!num_reads) return;
    Ctx &c = ctx->c;
    c.launches = 0;
    uint32_t *d_in = c.pool.dev<uint32_t>("pe.in", num_reads), *d_out = c.pool.dev<uint32_t>("pe.out", num_reads);
    SB_CUDA(cudaMemcpyAsync(d_in, order, sizeof(uint32_t) * num_reads, cudaMemcpyHostToDevice, c.stream));
    try { run_pe_encode(c, d_in, num_reads, d_out); }  // checks on the device that order is a permutation
    catch (const LimitError &e) { throw ArgError(e.what()); }
    SB_CUDA(cudaMemcpyAsync(order_out, d_out, sizeof(uint32_t) * num_reads, cudaMemcpyDeviceToHost, c.stream));
    SB_CUDA(cudaStreamSynchronize(c.stream));
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_reblock_streams(spring_b200_ctx *ctx, const spring_b200_streams *streams, const spring_b200_cp *cp,
                                spring_b200_blocks *out) {
  return guarded(ctx, [&] { reblock(ctx, streams, cp, out); });
}

int spring_b200_verify_roundtrip(spring_b200_ctx *ctx, const spring_b200_cp *cp, spring_b200_verify *out) {
  return guarded(ctx, [&] {
    if (!cp || !out) throw ArgError("null argument");
    if (!ctx->have_enc) throw ArgError("no streams resident on the device (run spring_b200_reorder_encode* first)");
    if (cp->num_reads_per_block <= 0) throw ArgError("cp.num_reads_per_block <= 0");
    if (ctx->last_enc.num_reads != cp->num_reads) throw ArgError("streams hold a different number of reads than cp.num_reads");
    Ctx &c = ctx->c;
    c.launches = 0;
    VerifyReport r;
    run_verify(c, ctx->last_enc, ctx->last_reads, ctx->last_lens, ctx->last_num_clean, ctx->last_W, ctx->last_nr,
               cp->paired_end != 0, cp->preserve_order != 0, (uint32_t)cp->num_reads_per_block, r);
    memset(out, 0, sizeof(*out));
    out->num_reads = r.num_reads; out->reads_checked = r.reads_checked; out->base_mismatch_reads = r.base_mismatch_reads;
    out->length_mismatch_reads = r.length_mismatch_reads; out->bad_order = r.bad_order; out->num_blocks = r.num_blocks;
    out->block_stream_bytes = r.block_stream_bytes; out->decoded_bases = r.decoded_bases;
    out->ok = (r.reads_checked == r.num_reads && !r.base_mismatch_reads && !r.length_mismatch_reads && !r.bad_order) ? 1 : 0;
    ctx->stats.gpu_launches = c.launches;
  });
}

int spring_b200_reblock_files(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp) {
  return guarded(ctx, [&] {
    if (!temp_dir || !cp) throw ArgError("null argument");
    const std::string dir(temp_dir);
    static const char *files[SPRING_B200_NUM_BLOCK_STREAMS] = {"read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin",
        "read_rev.txt", "read_unaligned.txt", "read_lengths.bin", "read_pos_pair.bin", "read_rev_pair.txt"};
    auto finish = [&](const spring_b200_blocks &b) {
      for (const char *f : {"read_noise.txt", "read_noisepos.bin", "read_rev.txt", "read_order.bin", "read_lengths.bin",
                            "read_unaligned.txt", "read_pos.bin", "read_unaligned.txt.count"})
        unlink((dir + "/" + f).c_str());  // :153, :174-181
      const int ns = cp->paired_end ? SPRING_B200_NUM_BLOCK_STREAMS : SPRING_B200_NUM_BLOCK_STREAMS - 2;
      parallel_ranges(b.num_blocks, [&](size_t lo, size_t hi) {
        for (size_t blk = lo; blk < hi; blk++)
          for (int st = 0; st < ns; st++)
            spill(dir + "/" + files[st] + "." + std::to_string(blk), b.data[st] + b.off[st][blk], b.off[st][blk + 1] - b.off[st][blk]);
      });
    };
    // The streams spring_b200_reorder_encode_files wrote into this temp_dir are still resident in HBM (same context,
    // nothing ran in between): no need to read the files back and upload them.  SPRING_B200_REBLOCK_FROM_FILES=1 forces
    // the file path.
    if (ctx->have_enc && ctx->files_dir == dir && ctx->last_enc.num_reads == cp->num_reads && !getenv("SPRING_B200_REBLOCK_FROM_FILES")) {
      spring_b200_blocks b{};
      reblock(ctx, nullptr, cp, &b);
      finish(b);
      return;
    }
    // the encoder's stream files (reorder_compress_streams.cpp:91-172)
    std::vector<uint8_t> f_pos = slurp(dir + "/read_pos.bin", true), f_noise = slurp(dir + "/read_noise.txt", true);
    std::vector<uint8_t> f_np = slurp(dir + "/read_noisepos.bin", true), f_rev = slurp(dir + "/read_rev.txt", true);
    std::vector<uint8_t> f_order = slurp(dir + "/read_order.bin", true), f_len = slurp(dir + "/read_lengths.bin", true);
    std::vector<uint8_t> f_un = slurp(dir + "/read_unaligned.txt", true), f_cnt = slurp(dir + "/read_unaligned.txt.count", true);
    if (f_cnt.size() != 8 || f_pos.size() != 8 * f_rev.size() || f_order.size() != 4 * (size_t)cp->num_reads ||
        f_len.size() != 2 * (size_t)cp->num_reads)
      throw IoError("reblock: stream files in " + dir + " are inconsistent with cp.num_reads");
    spring_b200_streams s{};
    s.pos = (const uint64_t *)f_pos.data(); s.noise = f_noise.data(); s.noise_bytes = f_noise.size();
    s.noisepos = (const uint16_t *)f_np.data(); s.num_noise = f_np.size() / 2; s.rev = f_rev.data();
    s.order = (const uint32_t *)f_order.data(); s.lengths = (const uint16_t *)f_len.data();
    s.unaligned = f_un.data(); s.unaligned_bytes = f_un.size(); memcpy(&s.unaligned_len, f_cnt.data(), 8);
    s.num_aligned = f_rev.size(); s.num_reads = cp->num_reads;
    spring_b200_blocks b{};
    reblock(ctx, &s, cp, &b);
    finish(b);
  });
}

int spring_b200_bucket_reads(spring_b200_ctx *ctx, const uint64_t *reads, const uint16_t *lengths, uint32_t num_reads,
                             uint32_t max_readlen, uint32_t num_buckets, uint32_t *bucket) {
  return guarded(ctx, [&] {
    if (max_readlen < 1 || max_readlen > (uint32_t)kMaxReadLen || !num_buckets) throw ArgError("bad argument");
    if (num_reads && (!reads || !lengths || !bucket)) throw ArgError("null pointer");
    bucket_reads(ctx->c, reads, lengths, num_reads, (int)max_readlen, num_buckets, bucket);
  });
}

// ---- multi-GPU ------------------------------------------------------------------------------------------------
int spring_b200_comm_unique_id(uint8_t *id) {
  if (!id) return SPRING_B200_EINVAL;
  try { comm_unique_id(id); return SPRING_B200_OK; }
  catch (const std::exception &e) { g_create_err = e.what(); return SPRING_B200_ECUDA; }
}

int spring_b200_comm_init(spring_b200_ctx *ctx, const uint8_t *id, int rank, int world) {
  return guarded(ctx, [&] {
    if (!id) throw ArgError("null id");
    if (ctx->comm) { comm_destroy(ctx->comm); ctx->comm = nullptr; }
    ctx->comm = comm_create(id, rank, world);
  });
}

int spring_b200_comm_free(spring_b200_ctx *ctx) {
  return guarded(ctx, [&] {
    if (ctx->comm) { SB_CUDA(cudaStreamSynchronize(ctx->c.stream)); comm_destroy(ctx->comm); ctx->comm = nullptr; }
  });
}

int spring_b200_exchange_reads(spring_b200_ctx *ctx, const uint64_t *reads, const uint16_t *lengths, const uint32_t *ids,
                               uint32_t num_reads, uint32_t max_readlen, spring_b200_exchanged *out) {
  return guarded(ctx, [&] {
    if (!out) throw ArgError("null output");
    if (max_readlen < 1 || max_readlen > (uint32_t)kMaxReadLen) throw ArgError("Wrong bitset size. (max_readlen must be 1..511)");
    if (num_reads && (!reads || !lengths || !ids)) throw ArgError("null pointer");
    if (num_reads >= 0x7FFFFFF0u) throw ArgError("too many reads for one GPU shard (>= 2^31)");
    ExchangeDev x;
    run_exchange(ctx->c, ctx->comm, reads, lengths, ids, num_reads, (int)max_readlen, x);
    memset(out, 0, sizeof(*out));
    out->reads = x.reads; out->lengths = x.lens; out->ids = x.ids; out->num_reads = x.n;
    out->sent_to_peers = x.sent_to_peers; out->received_from_peers = x.received_from_peers;
    ctx->stats.ms_exchange = 0;  // read lazily: the exchange is still in flight on the stream
  });
}

int spring_b200_finalize_shard(spring_b200_ctx *ctx, const uint32_t *ids, uint32_t num_owned, const uint32_t *n_ids, uint32_t num_n,
                               spring_b200_shard_layout *out) {
  return guarded(ctx, [&] {
    if (!out) throw ArgError("null output");
    if (!ctx->have_enc) throw ArgError("no streams resident on the device (run spring_b200_reorder_encode_device first)");
    if ((num_owned && !ids) || (num_n && !n_ids)) throw ArgError("null pointer");
    ShardLayout l;
    run_finalize_shard(ctx->c, ctx->comm, ctx->last_enc, ids, num_owned, n_ids, num_n, l);
    out->rank = l.rank; out->world = l.world; out->seq_base = l.seq_base; out->aligned_before = l.aligned_before;
    out->noise_before = l.noise_before; out->num_noise_before = l.num_noise_before;
    out->unaligned_reads_before = l.unaligned_reads_before; out->unaligned_bytes_before = l.unaligned_bytes_before;
    out->total_seq_len = l.total_seq_len; out->total_aligned = l.total_aligned; out->total_reads = l.total_reads;
    out->total_noise_bytes = l.total_noise_bytes; out->total_num_noise = l.total_num_noise;
    out->total_unaligned_bytes = l.total_unaligned_bytes; out->total_unaligned_len = l.total_unaligned_len;
    ctx->stats.ms_exchange = exchange_ms(ctx->c);
  });
}

namespace {
struct MergedOwner {
  std::vector<uint64_t> pos; std::vector<uint8_t> noise; std::vector<uint16_t> noisepos; std::vector<uint8_t> rev;
  std::vector<uint32_t> order; std::vector<uint16_t> lengths; std::vector<uint8_t> unaligned;
  std::vector<const uint8_t *> shard_seq; std::vector<uint64_t> shard_seq_len;
  std::vector<uint8_t> seq;  // all shards' consensus as ONE 2-bit stream (a shard need not end on a byte boundary)
};
}  // namespace

// What the reference's merge of its per-thread files does (src/encoder.h:386-453): every shard's aligned part in shard
// order, then every shard's unaligned part.  The shards are finalized (absolute positions, global ids), so this is pure
// concatenation; the big copies run on one host thread per stream.
int spring_b200_merge_shards(const spring_b200_streams *shards, int n, spring_b200_merged *out) {
  if (!shards || n < 1 || !out) return SPRING_B200_EINVAL;
  try {
    std::unique_ptr<MergedOwner> owner(new MergedOwner());
    MergedOwner *o = owner.get();
    uint64_t na = 0, nr = 0, nb = 0, nn = 0, ub = 0, ul = 0, sl = 0;
    for (int i = 0; i < n; i++) {
      const spring_b200_streams &s = shards[i];
      if (s.num_aligned > s.num_reads || s.noise_bytes != s.num_noise + s.num_aligned) return SPRING_B200_EINVAL;
      na += s.num_aligned; nr += s.num_reads; nb += s.noise_bytes; nn += s.num_noise; ub += s.unaligned_bytes; ul += s.unaligned_len;
      sl += s.seq_len;
      o->shard_seq.push_back(s.seq_packed); o->shard_seq_len.push_back(s.seq_len);
    }
    o->pos.resize(na); o->noise.resize(nb); o->noisepos.resize(nn); o->rev.resize(na); o->order.resize(nr); o->lengths.resize(nr);
    o->unaligned.resize(ub);
    std::vector<std::thread> th;
    auto cat = [&](auto *dst, auto field, auto count) {
      th.emplace_back([=] {
        size_t at = 0;
        for (int i = 0; i < n; i++) {
          const size_t c = count(shards[i]);
          if (c) memcpy(dst + at, field(shards[i]), c * sizeof(*dst));
          at += c;
        }
      });
    };
    cat(o->pos.data(), [](const spring_b200_streams &s) { return s.pos; }, [](const spring_b200_streams &s) { return (size_t)s.num_aligned; });
    cat(o->noise.data(), [](const spring_b200_streams &s) { return s.noise; }, [](const spring_b200_streams &s) { return (size_t)s.noise_bytes; });
    cat(o->noisepos.data(), [](const spring_b200_streams &s) { return s.noisepos; }, [](const spring_b200_streams &s) { return (size_t)s.num_noise; });
    cat(o->rev.data(), [](const spring_b200_streams &s) { return s.rev; }, [](const spring_b200_streams &s) { return (size_t)s.num_aligned; });
    cat(o->unaligned.data(), [](const spring_b200_streams &s) { return s.unaligned; }, [](const spring_b200_streams &s) { return (size_t)s.unaligned_bytes; });
    // order / lengths: aligned parts of all shards, then the unaligned parts
    th.emplace_back([=] {
      size_t at = 0;
      for (int i = 0; i < n; i++) { const size_t c = shards[i].num_aligned; if (c) { memcpy(o->order.data() + at, shards[i].order, c * 4); memcpy(o->lengths.data() + at, shards[i].lengths, c * 2); } at += c; }
      for (int i = 0; i < n; i++) {
        const size_t a0 = shards[i].num_aligned, c = shards[i].num_reads - a0;
        if (c) { memcpy(o->order.data() + at, shards[i].order + a0, c * 4); memcpy(o->lengths.data() + at, shards[i].lengths + a0, c * 2); }
        at += c;
      }
    });
    // the consensus: shard after shard at 2 bits per base; a shard that starts inside a byte is shifted into place
    o->seq.assign((size_t)((sl + 3) / 4) + 1, 0);
    th.emplace_back([=] {
      uint64_t at = 0;  // bases written
      for (int i = 0; i < n; i++) {
        const uint64_t len = shards[i].seq_len, nb = (len + 3) / 4;
        const uint8_t *src = shards[i].seq_packed;
        const int sh = 2 * (int)(at & 3);
        uint8_t *dst = o->seq.data() + at / 4;
        if (!len) continue;
        if (sh == 0) memcpy(dst, src, nb);
        else {
          for (uint64_t b = 0; b < nb; b++) {
            uint8_t v = src[b];
            if (b == nb - 1 && (len & 3)) v &= (uint8_t)((1u << (2 * (len & 3))) - 1u);  // bases beyond the shard's end
            dst[b] |= (uint8_t)(v << sh);
            dst[b + 1] |= (uint8_t)(v >> (8 - sh));
          }
        }
        at += len;
        if (at & 3) o->seq[at / 4] &= (uint8_t)((1u << (2 * (at & 3))) - 1u);
      }
    });
    for (auto &t : th) t.join();
    memset(out, 0, sizeof(*out));
    spring_b200_streams &m = out->streams;
    m.seq_packed = o->seq.data(); m.seq_len = sl; m.pos = o->pos.data(); m.noise = o->noise.data(); m.noise_bytes = nb;
    m.noisepos = o->noisepos.data(); m.num_noise = nn; m.rev = o->rev.data(); m.order = o->order.data(); m.lengths = o->lengths.data();
    m.unaligned = o->unaligned.data(); m.unaligned_bytes = ub; m.unaligned_len = ul; m.num_aligned = na; m.num_reads = nr;
    for (int i = 0; i < n; i++) { m.singletons_aligned += shards[i].singletons_aligned; m.n_reads_aligned += shards[i].n_reads_aligned; }
    out->num_shards = n; out->shard_seq = o->shard_seq.data(); out->shard_seq_len = o->shard_seq_len.data(); out->owner = owner.release();
    return SPRING_B200_OK;
  } catch (const std::exception &e) { g_create_err = e.what(); return SPRING_B200_ECUDA; }
}

void spring_b200_free_merged(spring_b200_merged *m) {
  if (!m || !m->owner) return;
  delete static_cast<MergedOwner *>(m->owner);
  memset(m, 0, sizeof(*m));
}

int spring_b200_write_merged(const char *temp_dir, const spring_b200_merged *m) {
  if (!temp_dir || !m || m->num_shards < 1) return SPRING_B200_EINVAL;
  try {
    const std::string dir(temp_dir);
    static const char code2char[4] = {'A', 'C', 'G', 'T'};
    for (int t = 0; t < m->num_shards; t++) {  // one read_seq.bin.<t> per shard, its last len % 4 bases as ASCII in .tail
      const uint64_t len = m->shard_seq_len[t], full = len / 4;
      const std::string base = dir + "/read_seq.bin." + std::to_string(t);
      spill(base, m->shard_seq[t], full);
      std::string tail;
      for (uint64_t x = full * 4; x < len; x++) tail.push_back(code2char[(m->shard_seq[t][x / 4] >> (2 * (x & 3))) & 3]);
      spill(base + ".tail", tail.data(), tail.size());
    }
    write_stream_files(dir, &m->streams);
    return SPRING_B200_OK;
  } catch (const IoError &e) { g_create_err = e.what(); return SPRING_B200_EIO;
  } catch (const std::exception &e) { g_create_err = e.what(); return SPRING_B200_ECUDA; }
}

int spring_b200_write_streams(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_streams *s, int num_shards) {
  return guarded(ctx, [&] {
    if (!temp_dir || !s) throw ArgError("null argument");
    write_streams(temp_dir, s, num_shards);
  });
}

int spring_b200_reorder_encode_files(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp, uint32_t num_chains) {
  return guarded(ctx, [&] {
    if (!temp_dir || !cp) throw ArgError("null argument");
    if (cp->long_flag) throw ArgError("long mode has no reorder/encode stage (spring.cpp:150)");
    const std::string dir(temp_dir);
    const uint32_t L = cp->max_readlen;
    if (L < 1 || L > (uint32_t)kMaxReadLen) throw ArgError("Wrong bitset size.");
    const int W = words_for((int)L);
    const uint32_t n0 = cp->num_reads_clean[0], n1 = cp->num_reads_clean[1], n = n0 + n1;
    Ctx &c = ctx->c;
    uint64_t *h_reads = c.pool.pin<uint64_t>("file.reads", (size_t)(n ? n : 1) * W);
    uint16_t *h_lens = c.pool.pin<uint16_t>("file.lens", n ? n : 1);
    const std::string f1 = dir + "/input_clean_1.dna", f2 = dir + "/input_clean_2.dna";
    const std::string fn = dir + "/input_N.dna", fo = dir + "/read_order_N.bin";
    { MappedFile b(f1, true); parse_dna(b, n0, W, L, h_reads, h_lens, "input_clean_1.dna"); }
    if (cp->paired_end) { MappedFile b(f2, true); parse_dna(b, n1, W, L, h_reads + (size_t)n0 * W, h_lens + n0, "input_clean_2.dna"); }
    std::vector<uint8_t> nrec = slurp(fn, false), nord = slurp(fo, false);
    spring_b200_input in{};
    in.reads = h_reads; in.lengths = h_lens; in.num_clean = n; in.max_readlen = L;
    in.n_records = nrec.data(); in.n_record_bytes = nrec.size();
    in.order_n = (const uint32_t *)nord.data(); in.num_n = cp->num_reads - n; in.num_reads = cp->num_reads;
    if (nord.size() != (size_t)in.num_n * 4) throw IoError("read_order_N.bin size does not match cp.num_reads");
    spring_b200_streams s{};
    run_all(ctx, &in, num_chains, true, &s);
    // inputs are consumed, as in the reference (reorder.h:232,241; encoder.h:606; encoder.cpp:218)
    unlink(f1.c_str()); unlink(f2.c_str()); unlink(fn.c_str()); unlink(fo.c_str());
    write_streams(dir, &s, cp->num_thr > 0 ? cp->num_thr : 1);
    ctx->files_dir = dir;
  });
}

int spring_b200_packed_pending(const spring_b200_ctx *ctx) { return ctx && ctx->packed_pending ? 1 : 0; }

// call_reorder behind a preprocess that packed the reads on the GPU (csrc/host/preprocess_b200.cpp): no input_clean_*.dna /
// input_N.dna / read_order_N.bin round trip -- the hot path starts from the rows spring_b200_pack_reads left in HBM and
// writes the same stream files as spring_b200_reorder_encode_files.
int spring_b200_reorder_encode_packed(spring_b200_ctx *ctx, const char *temp_dir, const spring_b200_cp *cp, uint32_t num_chains) {
  return guarded(ctx, [&] {
    if (!temp_dir || !cp) throw ArgError("null argument");
    if (!ctx->packed_pending) throw ArgError("no packed reads resident on the device (spring_b200_pack_reads with keep_on_device first)");
    const spring_b200_packed_reads &pk = ctx->packed;
    if (cp->long_flag) throw ArgError("long mode has no reorder/encode stage (spring.cpp:150)");
    if (cp->num_reads != pk.num_reads || cp->max_readlen != pk.max_readlen || cp->num_reads_clean[0] != pk.num_clean_file1 ||
        cp->num_reads_clean[0] + cp->num_reads_clean[1] != pk.num_clean)
      throw ArgError("cp does not describe the packed reads on the device");
    if (cp->max_readlen < 1 || cp->max_readlen > (uint32_t)kMaxReadLen) throw ArgError("Wrong bitset size.");
    spring_b200_input in{};
    in.reads = pk.reads; in.lengths = pk.lengths; in.num_clean = pk.num_clean; in.max_readlen = pk.max_readlen;
    in.n_records = pk.n_records; in.n_record_bytes = pk.n_record_bytes; in.order_n = pk.order_n; in.num_n = pk.num_n;
    in.num_reads = pk.num_reads;
    spring_b200_streams dev{}, s{};
    run_all(ctx, &in, num_chains, false, &dev);
    ctx->packed_pending = false;
    fetch(ctx, &s);
    write_streams(std::string(temp_dir), &s, cp->num_thr > 0 ? cp->num_thr : 1);
    ctx->files_dir = temp_dir;
  });
}

// ---- one process, several GPUs: the reference-facing form of the multi-GPU path (num_gpus / device_ids) ----
namespace {
struct NSplit { std::vector<size_t> off; };  // byte offset of every input_N.dna record (+ the end)
NSplit split_n_records(const uint8_t *rec, size_t bytes, uint32_t nn) {
  NSplit s;
  s.off.resize((size_t)nn + 1);
  size_t at = 0;
  for (uint32_t i = 0; i < nn; i++) {
    if (at + 2 > bytes) throw IoError("input_N.dna truncated");
    uint16_t len; memcpy(&len, rec + at, 2);
    s.off[i] = at;
    at += 2 + ((size_t)len + 1) / 2;
    if (at > bytes) throw IoError("input_N.dna truncated");
  }
  s.off[nn] = at;
  return s;
}
}  // namespace

int spring_b200_reorder_encode_files_multi(const char *temp_dir, const spring_b200_cp *cp, int num_gpus, const int *device_ids,
                                           uint32_t num_chains, spring_b200_stats *stats, char *err, size_t errlen) {
  auto fail = [&](int code, const std::string &m) {
    if (err && errlen) { strncpy(err, m.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return code;
  };
  if (!temp_dir || !cp || num_gpus < 1 || num_gpus > 64) return fail(SPRING_B200_EINVAL, "bad argument");
  if (cp->long_flag) return fail(SPRING_B200_EINVAL, "long mode has no reorder/encode stage (spring.cpp:150)");
  const uint32_t L = cp->max_readlen;
  if (L < 1 || L > (uint32_t)kMaxReadLen) return fail(SPRING_B200_EINVAL, "Wrong bitset size.");
  try {
    const int G = num_gpus, W = words_for((int)L);
    std::vector<spring_b200_ctx *> ctx(G, nullptr);
    for (int g = 0; g < G; g++)
      if (spring_b200_shared_ctx(device_ids ? device_ids[g] : g, &ctx[g]) != SPRING_B200_OK)
        return fail(SPRING_B200_ENODEV, std::string("spring_b200: ") + spring_b200_last_error(nullptr));
    if (G == 1) {
      const int rc = spring_b200_reorder_encode_files(ctx[0], temp_dir, cp, num_chains);
      if (rc != SPRING_B200_OK) return fail(rc, spring_b200_last_error(ctx[0]));
      if (stats) spring_b200_get_stats(ctx[0], stats);
      return SPRING_B200_OK;
    }
    const std::string dir(temp_dir);
    const uint32_t n0 = cp->num_reads_clean[0], n1 = cp->num_reads_clean[1], n = n0 + n1, nn = cp->num_reads - n;
    // the job's input, once, in pinned host memory of rank 0's context
    SB_CUDA(cudaSetDevice(ctx[0]->c.device));
    uint64_t *h_reads = ctx[0]->c.pool.pin<uint64_t>("file.reads", (size_t)(n ? n : 1) * W);
    uint16_t *h_lens = ctx[0]->c.pool.pin<uint16_t>("file.lens", n ? n : 1);
    const std::string f1 = dir + "/input_clean_1.dna", f2 = dir + "/input_clean_2.dna";
    const std::string fn = dir + "/input_N.dna", fo = dir + "/read_order_N.bin";
    { MappedFile b(f1, true); parse_dna(b, n0, W, L, h_reads, h_lens, "input_clean_1.dna"); }
    if (cp->paired_end) { MappedFile b(f2, true); parse_dna(b, n1, W, L, h_reads + (size_t)n0 * W, h_lens + n0, "input_clean_2.dna"); }
    std::vector<uint8_t> nrec = slurp(fn, false), nord = slurp(fo, false);
    if (nord.size() != (size_t)nn * 4) throw IoError("read_order_N.bin size does not match cp.num_reads");
    const uint32_t *order_n = reinterpret_cast<const uint32_t *>(nord.data());
    const NSplit ns = split_n_records(nrec.data(), nrec.size(), nn);
    uint8_t id[SPRING_B200_COMM_ID_BYTES];
    comm_unique_id(id);
    std::vector<std::string> errs(G);
    std::vector<spring_b200_streams> shard(G);
    std::vector<spring_b200_stats> st(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
      th.emplace_back([&, g] {
        try {
          spring_b200_ctx *x = ctx[g];
          SB_CUDA(cudaSetDevice(x->c.device));
          if (spring_b200_comm_init(x, id, g, G) != SPRING_B200_OK) throw CudaError(spring_b200_last_error(x));
          Ctx &c = x->c;
          // this rank's block of the clean reads and of the reads with N
          const uint32_t k0 = (uint32_t)((uint64_t)n * g / G), k1 = (uint32_t)((uint64_t)n * (g + 1) / G), nl = k1 - k0;
          const uint32_t j0 = (uint32_t)((uint64_t)nn * g / G), j1 = (uint32_t)((uint64_t)nn * (g + 1) / G), njl = j1 - j0;
          uint64_t *d_reads = c.pool.dev<uint64_t>("mg.reads", (size_t)(nl ? nl : 1) * W);
          uint16_t *d_lens = c.pool.dev<uint16_t>("mg.lens", nl ? nl : 1);
          uint32_t *d_ids = c.pool.dev<uint32_t>("mg.ids", nl ? nl : 1);
          uint32_t *d_on = c.pool.dev<uint32_t>("mg.order_n", (size_t)nn + 1);
          if (nl) {
            SB_CUDA(cudaMemcpyAsync(d_reads, h_reads + (size_t)k0 * W, sizeof(uint64_t) * (size_t)nl * W, cudaMemcpyHostToDevice, c.stream));
            SB_CUDA(cudaMemcpyAsync(d_lens, h_lens + k0, sizeof(uint16_t) * nl, cudaMemcpyHostToDevice, c.stream));
          }
          if (nn) SB_CUDA(cudaMemcpyAsync(d_on, order_n, sizeof(uint32_t) * nn, cudaMemcpyHostToDevice, c.stream));
          run_original_ids(c, k0, nl, d_on, nn, d_ids);
          spring_b200_exchanged ex{};
          if (spring_b200_exchange_reads(x, d_reads, d_lens, d_ids, nl, L, &ex) != SPRING_B200_OK) throw CudaError(spring_b200_last_error(x));
          std::vector<uint32_t> on_local(njl);
          for (uint32_t j = 0; j < njl; j++) on_local[j] = ex.num_reads + j;  // the rank's N reads are numbered after the reads it owns
          spring_b200_input in{};
          in.reads = ex.reads; in.lengths = ex.lengths; in.num_clean = ex.num_reads; in.max_readlen = L;
          in.n_records = nrec.data() + ns.off[j0]; in.n_record_bytes = ns.off[j1] - ns.off[j0];
          in.order_n = on_local.data(); in.num_n = njl; in.num_reads = ex.num_reads + njl;
          spring_b200_streams dev_s{};
          if (spring_b200_reorder_encode_device(x, &in, num_chains, &dev_s) != SPRING_B200_OK) throw CudaError(spring_b200_last_error(x));
          spring_b200_get_stats(x, &st[g]);
          spring_b200_shard_layout lay{};
          if (spring_b200_finalize_shard(x, ex.ids, ex.num_reads, order_n + j0, njl, &lay) != SPRING_B200_OK) throw CudaError(spring_b200_last_error(x));
          if (spring_b200_fetch_streams(x, &shard[g]) != SPRING_B200_OK) throw CudaError(spring_b200_last_error(x));
        } catch (const std::exception &e) { errs[g] = e.what(); }
      });
    for (auto &t : th) t.join();
    for (int g = 0; g < G; g++) if (!errs[g].empty()) return fail(SPRING_B200_ECUDA, "GPU " + std::to_string(g) + ": " + errs[g]);
    spring_b200_merged m{};
    if (spring_b200_merge_shards(shard.data(), G, &m) != SPRING_B200_OK) return fail(SPRING_B200_ECUDA, "merge failed");
    unlink(f1.c_str()); unlink(f2.c_str()); unlink(fn.c_str()); unlink(fo.c_str());
    try { write_streams(dir, &m.streams, cp->num_thr > 0 ? cp->num_thr : 1); }
    catch (...) { spring_b200_free_merged(&m); throw; }
    spring_b200_free_merged(&m);
    for (int g = 0; g < G; g++) ctx[g]->files_dir.clear();  // a shard's resident streams are not the job's files
    if (stats) {
      *stats = st[0];
      for (int g = 1; g < G; g++) { stats->unmatched += st[g].unmatched; stats->singletons_aligned += st[g].singletons_aligned; stats->n_reads_aligned += st[g].n_reads_aligned; stats->num_chains += st[g].num_chains; }
    }
    return SPRING_B200_OK;
  } catch (const IoError &e) { return fail(SPRING_B200_EIO, e.what());
  } catch (const LimitError &e) { return fail(SPRING_B200_ELIMIT, e.what());
  } catch (const std::exception &e) { return fail(SPRING_B200_ECUDA, e.what()); }
}

}  // extern "C"
