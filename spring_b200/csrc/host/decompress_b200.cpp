// decompress_b200.cpp -- spring::decompress_short with the read reconstruction on the GPU.
//
// Drop-in for the `decompress_short` symbol of shubhamchandak94/Spring (src/decompress.h:22-26, called at
// src/spring.cpp:366).  Per step of num_thr blocks the reference BSC-decodes the block streams and then rebuilds
// every read with a per-read loop (src/decompress.cpp:230-320: position deltas, consensus substring, noise
// substitution through dec_noise, reverse complement, mate position / strand).  Here that loop is ONE call of
// spring_b200_decode_blocks per step; BSC, ids, qualities and the FASTQ writer stay the reference's own host code
// (bsc::BSC_decompress, decompress_id_block, bsc::BSC_str_array_decompress, write_fastq_block).  The consensus is
// kept 2 bits per base (the reference expands it to one char per base, src/decompress.cpp:615-660).
//
// Build: compile the reference's src/decompress.cpp with -Ddecompress_short=decompress_short_reference
// (decompress_long / decompress_unpack_seq are still used from it) and add this file; see oracle/Makefile target
// splice3 and INTEGRATION.md 1c.  SPRING_B200_DECODE=0 hands the call to the reference's loop.
#include <omp.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "decompress.h"
#include "libbsc/bsc.h"
#include "spring_b200.h"
#include "util.h"

namespace spring {

// the reference's own implementation (src/decompress.cpp compiled with -Ddecompress_short=decompress_short_reference)
void decompress_short_reference(const std::string &temp_dir, const std::string &outfile_1, const std::string &outfile_2,
                                const compression_params &cp, const int &num_thr, const uint64_t &start_num, const uint64_t &end_num,
                                const bool &gzip_flag, const int &gzip_level);

namespace {

std::vector<uint8_t> file_bytes(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("cannot open " + path);
  f.seekg(0, std::ios::end);
  const std::streamoff n = f.tellg();
  f.seekg(0);
  std::vector<uint8_t> v((size_t)n);
  if (n) f.read(reinterpret_cast<char *>(v.data()), n);
  return v;
}

// <raw>.bsc -> bytes; both files are removed, as the reference removes them (decompress.cpp:155-196, :331-350)
std::vector<uint8_t> bsc_file_bytes(const std::string &raw) {
  const std::string packed = raw + ".bsc";
  bsc::BSC_decompress(packed.c_str(), raw.c_str());
  remove(packed.c_str());
  std::vector<uint8_t> v = file_bytes(raw);
  remove(raw.c_str());
  return v;
}

// read_seq.bin.<t> shards -> one stream, 2 bits per base A0 C1 G2 T3, 4 bases per byte LSB first.  A shard is
// whole bytes + a tail of up to 3 ASCII bases (encoder.cpp:126-141), so a later shard can start inside a byte.
struct Consensus {
  std::vector<uint8_t> packed;
  uint64_t len = 0;
  void push_base(unsigned code) {
    if ((len & 3) == 0) packed.push_back(0);
    packed.back() |= (uint8_t)(code << (2 * (len & 3)));
    len++;
  }
  void append(const std::vector<uint8_t> &bytes, const std::vector<uint8_t> &tail) {
    if ((len & 3) == 0) {
      packed.insert(packed.end(), bytes.begin(), bytes.end());
      len += 4ull * bytes.size();
    } else {
      const int sh = 2 * (int)(len & 3);
      for (uint8_t b : bytes) {
        packed.back() |= (uint8_t)(b << sh);
        packed.push_back((uint8_t)(b >> (8 - sh)));
        len += 4;
      }
    }
    for (uint8_t c : tail) {
      if (c == '\n' || c == '\r') continue;
      const char *p = strchr("ACGT", c);
      if (!p || !c) throw std::runtime_error("read_seq tail holds a character other than A, C, G, T");
      push_base((unsigned)(p - "ACGT"));
    }
  }
};

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

void decompress_short(const std::string &temp_dir, const std::string &outfile_1, const std::string &outfile_2,
                      const compression_params &cp, const int &num_thr, const uint64_t &start_num, const uint64_t &end_num,
                      const bool &gzip_flag, const int &gzip_level) {
  if (env_int("SPRING_B200_DECODE", 1) == 0) {
    decompress_short_reference(temp_dir, outfile_1, outfile_2, cp, num_thr, start_num, end_num, gzip_flag, gzip_level);
    return;
  }
  spring_b200_ctx *ctx = nullptr;
  if (spring_b200_shared_ctx(env_int("SPRING_B200_DEVICE", 0), &ctx) != SPRING_B200_OK)
    throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(nullptr));

  // stream files of a block, in the order of enum spring_b200_block_stream
  const char *stream_name[SPRING_B200_NUM_BLOCK_STREAMS] = {"read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin",
                                                             "read_rev.txt", "read_unaligned.txt", "read_lengths.bin",
                                                             "read_pos_pair.bin", "read_rev_pair.txt"};
  const int nstreams = cp.paired_end ? 9 : 7;
  const int nfiles = cp.paired_end ? 2 : 1;
  const std::string q_path[2] = {temp_dir + "/quality_1", temp_dir + "/quality_2"};
  const std::string id_path[2] = {temp_dir + "/id_1", temp_dir + "/id_2"};
  const std::string out_path[2] = {outfile_1, outfile_2};
  const uint32_t per_block = (uint32_t)cp.num_reads_per_block;
  const uint64_t units = cp.paired_end ? cp.num_reads / 2 : cp.num_reads;  // reads, or pairs

  std::ofstream fout[2];
  for (int j = 0; j < nfiles; j++) {
    if (gzip_flag) fout[j].open(out_path[j], std::ios::binary); else fout[j].open(out_path[j]);
    if (!fout[j].is_open()) throw std::runtime_error("Error opening output file");
  }
  omp_set_num_threads(num_thr);

  // ---- consensus: BSC-decode every shard (in parallel, like decompress_unpack_seq), keep it packed -----------------
  Consensus cons;
  {
    const int shards = cp.num_thr;
    std::vector<std::vector<uint8_t>> body((size_t)shards), tail((size_t)shards);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < shards; t++) {
      const std::string base = temp_dir + "/read_seq.bin." + std::to_string(t);
      body[t] = bsc_file_bytes(base);
      tail[t] = file_bytes(base + ".tail");
      remove((base + ".tail").c_str());
    }
    for (int t = 0; t < shards; t++) cons.append(body[t], tail[t]);
  }

  const uint64_t per_step = std::min<uint64_t>((uint64_t)num_thr * per_block, units);
  std::vector<std::string> reads[2], ids(per_step), quals(cp.preserve_quality ? per_step : 0);
  std::vector<uint32_t> lens[2];
  for (int j = 0; j < nfiles; j++) { reads[j].resize(per_step); lens[j].resize(per_step); }

  const uint32_t first_block = (uint32_t)(start_num / per_block);
  uint32_t blocks_done = first_block;
  uint64_t done_units = (uint64_t)first_block * per_block;
  bool finished = false;
  while (!finished) {
    const uint64_t cur = std::min<uint64_t>(per_step, units - done_units);
    if (cur == 0) break;
    const uint32_t nblk = (uint32_t)((cur + per_block - 1) / per_block);

    // ---- this step's block streams: BSC on the host threads, then one GPU decode ----------------------------------
    std::vector<std::vector<uint8_t>> piece((size_t)nblk * SPRING_B200_NUM_BLOCK_STREAMS);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t k = 0; k < (int64_t)nblk * nstreams; k++) {
      const uint32_t b = (uint32_t)(k / nstreams);
      const int s = (int)(k % nstreams);
      piece[(size_t)b * SPRING_B200_NUM_BLOCK_STREAMS + s] =
          bsc_file_bytes(temp_dir + "/" + stream_name[s] + "." + std::to_string(blocks_done + b));
    }
    std::vector<uint8_t> data[SPRING_B200_NUM_BLOCK_STREAMS];
    std::vector<uint64_t> off[SPRING_B200_NUM_BLOCK_STREAMS];
    spring_b200_blocks blk;
    memset(&blk, 0, sizeof(blk));
    blk.num_blocks = nblk;
    for (int s = 0; s < SPRING_B200_NUM_BLOCK_STREAMS; s++) {
      off[s].assign((size_t)nblk + 1, 0);
      for (uint32_t b = 0; b < nblk; b++) {
        const std::vector<uint8_t> &p = piece[(size_t)b * SPRING_B200_NUM_BLOCK_STREAMS + s];
        data[s].insert(data[s].end(), p.begin(), p.end());
        off[s][b + 1] = data[s].size();
      }
      blk.data[s] = data[s].data(); blk.size[s] = data[s].size(); blk.off[s] = off[s].data();
    }
    blk.num_reads = cur * nfiles;
    spring_b200_cp step_cp;
    static_assert(sizeof(step_cp) == sizeof(cp), "cp.bin layout mismatch");
    memcpy(&step_cp, &cp, sizeof(cp));
    step_cp.num_reads = (uint32_t)(cur * nfiles);
    spring_b200_decoded dec;
    if (spring_b200_decode_blocks(ctx, &blk, cons.packed.data(), cons.len, &step_cp, &dec) != SPRING_B200_OK)
      throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(ctx));
    // file 1's reads of the step, then file 2's
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)cur; i++)
      for (int j = 0; j < nfiles; j++) {
        const uint64_t r = (uint64_t)j * cur + (uint64_t)i, a = dec.offsets[r], e = dec.offsets[r + 1];
        reads[j][i].assign(reinterpret_cast<const char *>(dec.bases) + a, (size_t)(e - a));
        lens[j][i] = (uint32_t)(e - a);
      }

    for (int j = 0; j < nfiles; j++) {
      // ---- qualities and ids of the step's blocks (decompress.cpp:353-392) ---------------------------------------
#pragma omp parallel for schedule(static, 1)
      for (int64_t b = 0; b < (int64_t)nblk; b++) {
        const uint64_t i0 = (uint64_t)b * per_block, cnt = std::min<uint64_t>(cur, i0 + per_block) - i0;
        const std::string suffix = "." + std::to_string(blocks_done + (uint32_t)b);
        if (cp.preserve_quality) {
          bsc::BSC_str_array_decompress((q_path[j] + suffix).c_str(), quals.data() + i0, (uint32_t)cnt, lens[j].data() + i0);
          remove((q_path[j] + suffix).c_str());
        }
        if (!cp.preserve_id) {
          for (uint64_t i = i0; i < i0 + cnt; i++) ids[i] = "@" + std::to_string(done_units + i + 1) + "/" + std::to_string(j + 1);
        } else if (j == 1 && cp.paired_id_match) {
          for (uint64_t i = i0; i < i0 + cnt; i++) modify_id(ids[i], cp.paired_id_code);
        } else {
          decompress_id_block((id_path[j] + suffix).c_str(), ids.data() + i0, (uint32_t)cnt);
          remove((id_path[j] + suffix).c_str());
        }
      }
      // ---- the requested range of records (decompress.cpp:395-414) ------------------------------------------------
      uint64_t out_n = cur;
      if (done_units + out_n >= end_num) { out_n = end_num - done_units; finished = true; }
      const uint64_t skip = blocks_done == first_block ? start_num % per_block : 0;
      write_fastq_block(fout[j], ids.data() + skip, reads[j].data() + skip, cp.preserve_quality ? quals.data() + skip : nullptr,
                        (uint32_t)(out_n - skip), cp.preserve_quality, num_thr, gzip_flag, gzip_level);
    }
    done_units += cur;
    blocks_done += (uint32_t)num_thr;
  }
  for (int j = 0; j < nfiles; j++) fout[j].close();
}

}  // namespace spring
