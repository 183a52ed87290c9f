// preprocess_b200.cpp -- spring::preprocess for short reads with the read path on the GPU.
//
// Drop-in for the `preprocess` symbol of shubhamchandak94/Spring (src/preprocess.h:22-24, called at
// src/spring.cpp:118).  The reference's preprocess does three things per block of FASTQ records: (1) ids and
// qualities (binning, qvz, per-block compression in order-preserving mode, plain files otherwise), (2) checks, and
// (3) the read path: split reads with N from clean ones and pack them 4 / 2 bits per base into input_N.dna /
// input_clean_{1,2}.dna (src/preprocess.cpp:293-304, src/util.cpp:269-348).  (1) and (2) stay host work and use the
// reference's own helpers (read_fastq_block, compress_id_block, quantize_quality, bsc::BSC_str_array_compress, ...);
// (3) becomes ONE call of spring_b200_pack_reads on the sequence lines, and the packed rows stay in HBM for
// call_reorder (spring_b200_reorder_encode_packed): the .dna files are never written.
//
// Build: compile the reference's src/preprocess.cpp with -Dpreprocess=preprocess_reference (its long-read mode is
// still used through that name) and add this file; see oracle/Makefile target splice3 and INTEGRATION.md 1c.
// With SPRING_B200_GPUS > 1 (or SPRING_B200_PREPROCESS_FILES=1) the packed records are written to the .dna files
// instead, for the multi-GPU file entry point.
#include <omp.h>
#include <algorithm>
#include <boost/iostreams/copy.hpp>
#include <boost/iostreams/filter/gzip.hpp>
#include <boost/iostreams/filtering_streambuf.hpp>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "libbsc/bsc.h"
#include "params.h"
#include "preprocess.h"
#include "spring_b200.h"
#include "util.h"

namespace spring {

// the reference's own implementation (src/preprocess.cpp compiled with -Dpreprocess=preprocess_reference)
void preprocess_reference(const std::string &infile_1, const std::string &infile_2, const std::string &temp_dir,
                          compression_params &cp, const bool &gzip_flag, const bool &fasta_flag);

namespace {

namespace io = boost::iostreams;

// one FASTQ / FASTA input, plain or gzip'd (src/preprocess.cpp:71-82)
struct Input {
  std::ifstream file;
  std::unique_ptr<io::filtering_streambuf<io::input>> gz;
  std::unique_ptr<std::istream> gz_stream;
  std::string path;
  bool gzip = false;
  void open(const std::string &p, bool gzip_flag) {
    path = p; gzip = gzip_flag;
    rewind();
  }
  void rewind() {
    gz_stream.reset(); gz.reset();
    if (file.is_open()) file.close();
    file.clear();
    if (gzip) {
      file.open(path, std::ios_base::binary);
      gz.reset(new io::filtering_streambuf<io::input>);
      gz->push(io::gzip_decompressor());
      gz->push(file);
      gz_stream.reset(new std::istream(gz.get()));
    } else {
      file.open(path);
    }
  }
  std::istream *stream() { return gzip ? gz_stream.get() : &file; }
};

// the sequence lines of one input file, concatenated (what spring_b200_pack_reads takes)
struct Bases {
  std::vector<uint8_t> text;
  std::vector<uint64_t> start;  // start[i] of read i; the end is appended when the file is done
  void add(const std::string &r) { start.push_back(text.size()); text.insert(text.end(), r.begin(), r.end()); }
};

int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

// input_clean_<j>.dna records {u16 len; ceil(len / 4) bytes} (src/util.cpp:269-294): a packed row IS the record's
// payload (base k sits in bits 2 (k % 4) of byte k / 4 in both), so the file is the rows cut to their lengths
void write_clean_records(const std::string &path, const uint64_t *rows, const uint16_t *lens, uint32_t n, int W) {
  std::ofstream f(path, std::ios::binary);
  if (!f.is_open()) throw std::runtime_error("cannot create " + path);
  for (uint32_t i = 0; i < n; i++) {
    f.write(reinterpret_cast<const char *>(&lens[i]), 2);
    f.write(reinterpret_cast<const char *>(rows + (size_t)i * W), (lens[i] + 3) / 4);
  }
}

}  // namespace

void preprocess(const std::string &infile_1, const std::string &infile_2, const std::string &temp_dir, compression_params &cp,
                const bool &gzip_flag, const bool &fasta_flag) {
  if (cp.long_flag) {  // long reads have no reorder / encode stage (spring.cpp:150): nothing of this path is involved
    preprocess_reference(infile_1, infile_2, temp_dir, cp, gzip_flag, fasta_flag);
    return;
  }
  const int nfiles = cp.paired_end ? 2 : 1;
  const std::string in_path[2] = {infile_1, infile_2};
  const std::string id_path[2] = {temp_dir + "/id_1", temp_dir + "/id_2"};
  const std::string q_path[2] = {temp_dir + "/quality_1", temp_dir + "/quality_2"};
  const bool flat_side_files = !cp.preserve_order;  // ids / qualities as text, re-ordered and compressed later

  Input in[2];
  std::ofstream id_out[2], q_out[2];
  for (int j = 0; j < nfiles; j++) {
    in[j].open(in_path[j], gzip_flag);
    if (!in[j].file.is_open()) throw std::runtime_error("Error opening input file");
    if (flat_side_files) {
      if (cp.preserve_id) id_out[j].open(id_path[j]);
      if (cp.preserve_quality) q_out[j].open(q_path[j]);
    }
  }

  // paired ids that differ only by a fixed pattern are stored once (src/preprocess.cpp:115-139)
  uint8_t id_code = 0;
  bool id_match = false;
  if (cp.paired_end && cp.preserve_id) {
    std::string a, b;
    std::getline(*in[0].stream(), a);
    std::getline(*in[1].stream(), b);
    id_code = find_id_pattern(a, b);
    id_match = id_code != 0;
    in[0].rewind(); in[1].rewind();
  }

  std::vector<char> bin_table(128);
  if (cp.ill_bin_flag) generate_illumina_binning_table(bin_table.data());
  if (cp.bin_thr_flag) generate_binary_binning_table(bin_table.data(), cp.bin_thr_thr, cp.bin_thr_high, cp.bin_thr_low);

  const uint32_t per_block = (uint32_t)cp.num_reads_per_block;
  const uint64_t per_step = (uint64_t)cp.num_thr * per_block;
  std::vector<std::string> reads(per_step), ids_1(per_step), ids_2(per_step), quals(per_step);
  std::vector<uint32_t> lens(per_step);
  std::vector<char> thread_id_match((size_t)cp.num_thr);
  omp_set_num_threads(cp.num_thr);

  Bases seq[2];
  uint64_t num_reads[2] = {0, 0};
  uint32_t blocks_done = 0;
  for (bool more = true; more; blocks_done += (uint32_t)cp.num_thr) {
    more = false;
    for (int j = 0; j < nfiles; j++) {
      std::vector<std::string> &ids = j == 0 ? ids_1 : ids_2;
      const uint32_t got = read_fastq_block(in[j].stream(), ids.data(), reads.data(), quals.data(), (uint32_t)per_step, fasta_flag);
      if (got == per_step) more = true;
      if (got == 0) continue;
      if (num_reads[0] + num_reads[1] + got > MAX_NUM_READS) {
        std::cerr << "Max number of reads allowed is " << MAX_NUM_READS << "\n";
        throw std::runtime_error("Too many reads.");
      }
      std::string failure;  // exceptions must not leave an OpenMP region: first message wins
#pragma omp parallel
      {
        const uint64_t t = (uint64_t)omp_get_thread_num();
        const uint64_t b0 = t * per_block, b1 = std::min<uint64_t>(got, b0 + per_block);
        bool match = id_match;
        std::string err;
        if (b0 < got) {
          const uint32_t cnt = (uint32_t)(b1 - b0);
          for (uint64_t i = b0; i < b1 && err.empty(); i++) {
            const size_t len = reads[i].size();
            if (len > MAX_READ_LEN) {
              std::cerr << "Max read length without long mode is " << MAX_READ_LEN << ", but found read of length " << len << "\n";
              err = "Too long read length (please try --long/-l flag).";
            } else if (cp.preserve_quality && quals[i].size() != len) {
              err = "Read length does not match quality length.";
            }
            lens[i] = (uint32_t)len;
            if (j == 1 && match) match = check_id_pattern(ids_1[i], ids_2[i], id_code);
          }
          if (err.empty()) {
            if (cp.preserve_quality && (cp.ill_bin_flag || cp.bin_thr_flag)) quantize_quality(quals.data() + b0, cnt, bin_table.data());
            if (cp.preserve_quality && cp.qvz_flag && cp.preserve_order) quantize_quality_qvz(quals.data() + b0, cnt, lens.data() + b0, cp.qvz_ratio);
            if (!flat_side_files) {  // order kept: ids and qualities are final, compress them block by block now
              const std::string blk = "." + std::to_string(blocks_done + t);
              if (cp.preserve_id) compress_id_block((id_path[j] + blk).c_str(), ids.data() + b0, cnt);
              if (cp.preserve_quality) bsc::BSC_str_array_compress((q_path[j] + blk).c_str(), quals.data() + b0, cnt, lens.data() + b0);
            }
          }
        }
        thread_id_match[t] = match;
        if (!err.empty()) {
#pragma omp critical
          if (failure.empty()) failure = err;
        }
      }
      if (!failure.empty()) throw std::runtime_error(failure);
      if (j == 1) {
        for (int t = 0; t < cp.num_thr && id_match; t++) id_match = thread_id_match[t] != 0;
        if (!id_match) id_code = 0;
      }
      // the read path: the sequence lines are only collected here; N split and packing happen on the GPU below
      for (uint32_t i = 0; i < got; i++) seq[j].add(reads[i]);
      if (flat_side_files) {
        if (cp.preserve_quality) for (uint32_t i = 0; i < got; i++) q_out[j] << quals[i] << "\n";
        if (cp.preserve_id) for (uint32_t i = 0; i < got; i++) id_out[j] << ids[i] << "\n";
      }
      num_reads[j] += got;
    }
    if (cp.paired_end && num_reads[0] != num_reads[1]) throw std::runtime_error("Number of reads in paired files do not match.");
  }
  for (int j = 0; j < nfiles; j++) {
    if (id_out[j].is_open()) id_out[j].close();
    if (q_out[j].is_open()) q_out[j].close();
  }
  if (num_reads[0] == 0) throw std::runtime_error("No reads found.");

  // ---- N split + packing on the GPU (preprocess.cpp:293-304, :364-378 in one call) ----------------------------------
  const uint64_t total = num_reads[0] + num_reads[1];
  std::vector<uint64_t> offsets;
  offsets.reserve(total + 1);
  offsets = std::move(seq[0].start);
  const uint64_t bytes_1 = seq[0].text.size();
  for (uint64_t s : seq[1].start) offsets.push_back(bytes_1 + s);
  offsets.push_back(bytes_1 + seq[1].text.size());
  seq[0].text.insert(seq[0].text.end(), seq[1].text.begin(), seq[1].text.end());
  std::vector<uint8_t>().swap(seq[1].text);
  std::vector<uint64_t>().swap(seq[1].start);

  spring_b200_ctx *ctx = nullptr;
  if (spring_b200_shared_ctx(env_int("SPRING_B200_DEVICE", 0), &ctx) != SPRING_B200_OK)
    throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(nullptr));
  const bool to_files = env_int("SPRING_B200_GPUS", 1) > 1 || env_int("SPRING_B200_PREPROCESS_FILES", 0) != 0;
  spring_b200_packed_reads pk;
  if (spring_b200_pack_reads(ctx, seq[0].text.data(), offsets.data(), (uint32_t)total, (uint32_t)num_reads[0], to_files ? 0 : 1, &pk) !=
      SPRING_B200_OK)
    throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(ctx));
  if (to_files) {
    const int W = (2 * (int)(pk.max_readlen ? pk.max_readlen : 1) - 1) / 64 + 1;
    write_clean_records(temp_dir + "/input_clean_1.dna", pk.reads, pk.lengths, pk.num_clean_file1, W);
    if (cp.paired_end)
      write_clean_records(temp_dir + "/input_clean_2.dna", pk.reads + (size_t)pk.num_clean_file1 * W, pk.lengths + pk.num_clean_file1,
                          pk.num_clean - pk.num_clean_file1, W);
    std::ofstream fn(temp_dir + "/input_N.dna", std::ios::binary), fo(temp_dir + "/read_order_N.bin", std::ios::binary);
    fn.write(reinterpret_cast<const char *>(pk.n_records), (std::streamsize)pk.n_record_bytes);
    fo.write(reinterpret_cast<const char *>(pk.order_n), (std::streamsize)pk.num_n * 4);
  }

  if (cp.paired_end && id_match) {  // the second file's ids follow from the first's (preprocess.cpp:381-393)
    if (flat_side_files) {
      remove(id_path[1].c_str());
    } else {
      const uint32_t nblocks = 1 + (uint32_t)((num_reads[0] - 1) / per_block);
      for (uint32_t b = 0; b < nblocks; b++) remove((id_path[1] + "." + std::to_string(b)).c_str());
    }
  }
  cp.paired_id_code = id_code;
  cp.paired_id_match = id_match;
  cp.num_reads = (uint32_t)total;
  cp.num_reads_clean[0] = pk.num_clean_file1;
  cp.num_reads_clean[1] = pk.num_clean - pk.num_clean_file1;
  cp.max_readlen = pk.max_readlen;

  std::cout << "Max Read length: " << cp.max_readlen << "\n";
  std::cout << "Total number of reads: " << cp.num_reads << "\n";
  std::cout << "Total number of reads without N: " << cp.num_reads_clean[0] + cp.num_reads_clean[1] << "\n";
  if (cp.preserve_id && cp.paired_end) std::cout << "Paired id match code: " << (int)cp.paired_id_code << "\n";
}

}  // namespace spring
