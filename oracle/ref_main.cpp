// Driver around the UNMODIFIED reference sources (compiled in place from
// $(REF)/src by oracle/Makefile; nothing is copied into this repo).
// TEST INFRASTRUCTURE ONLY.
//
// Replaces the reference's src/main.cpp (boost::program_options is not in this
// image) with a plain argv parser that accepts the same flags for the subset
// we use, and adds two modes that expose the hot path on its own:
//
//   spring_ref -c -i A [B] -o OUT [-r] [-t N] [--no-quality] [--no-ids] [-w DIR]
//   spring_ref -d -i IN -o OUT [OUT2] [-t N] [-w DIR]
//   spring_ref --preprocess -i A [B] --temp DIR [-r] [-t N] [--no-quality] [--no-ids]
//        runs spring::preprocess only, leaves its files + cp_in.bin in DIR
//   spring_ref --hotpath --temp DIR [-t N] [--unbsc]
//        reads DIR/cp_in.bin, runs call_reorder + call_encoder
//        (spring.cpp:153,166), prints "HOTPATH_SECONDS reorder encode";
//        --unbsc also BSC-decodes read_seq.bin.<t>.bsc to read_seq.bin.<t>
//   spring_ref --reblock --temp DIR [-t N]
//        reads DIR/cp_in.bin, runs pe_encode (spring.cpp:193, only for -r paired input) and
//        reorder_compress_streams (spring.cpp:206) on the encoder streams in DIR, then BSC-decodes
//        every per-block file it wrote (X.<b>.bsc -> X.<b>) so the raw block streams can be compared
//   spring_ref --bsc-decode -i F.bsc [G.bsc ...]
//        bsc::BSC_decompress of each file into the same name without ".bsc" (to look inside an archive)
//
// When built with -DSPRING_B200_SPLICE the two call_* symbols come from
// spring_b200/csrc/host/call_template_functions_b200.cpp (our CUDA library
// behind the reference's own interface) instead of the reference's
// call_template_functions.cpp.
#include <omp.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>
#include <boost/filesystem.hpp>
#include "call_template_functions.h"
#include "libbsc/bsc.h"
#include "pe_encode.h"
#include "preprocess.h"
#include "reorder_compress_streams.h"
#include "spring.h"
#include "util.h"

static double now_s() {
  return std::chrono::duration<double>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}

int main(int argc, char **argv) {
  bool compress_flag = false, decompress_flag = false, pairing_only = false,
       no_quality = false, no_ids = false, pre_flag = false, hot_flag = false,
       unbsc = false, reblock_flag = false, bscdec_flag = false;
  std::vector<std::string> in_vec, out_vec, quality_opts;
  std::vector<uint64_t> range_vec;
  std::string working_dir = ".", temp_given;
  int num_thr = 8;
  std::vector<std::string> *cur = NULL;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-c" || a == "--compress") { compress_flag = true; cur = NULL; }
    else if (a == "-d" || a == "--decompress") { decompress_flag = true; cur = NULL; }
    else if (a == "-r" || a == "--allow-read-reordering") { pairing_only = true; cur = NULL; }
    else if (a == "--no-quality") { no_quality = true; cur = NULL; }
    else if (a == "--no-ids") { no_ids = true; cur = NULL; }
    else if (a == "--preprocess") { pre_flag = true; cur = NULL; }
    else if (a == "--hotpath") { hot_flag = true; cur = NULL; }
    else if (a == "--unbsc") { unbsc = true; cur = NULL; }
    else if (a == "--reblock") { reblock_flag = true; cur = NULL; }
    else if (a == "--bsc-decode") { bscdec_flag = true; cur = NULL; }
    else if (a == "-i" || a == "--input-file") cur = &in_vec;
    else if (a == "-o" || a == "--output-file") cur = &out_vec;
    else if ((a == "-t" || a == "--num-threads") && i + 1 < argc) { num_thr = atoi(argv[++i]); cur = NULL; }
    else if ((a == "-w" || a == "--working-dir") && i + 1 < argc) { working_dir = argv[++i]; cur = NULL; }
    else if (a == "--temp" && i + 1 < argc) { temp_given = argv[++i]; cur = NULL; }
    else if (cur) cur->push_back(a);
    else { std::cerr << "unknown argument " << a << "\n"; return 2; }
  }
  try {
    if (pre_flag) {
      spring::compression_params cp;
      memset(&cp, 0, sizeof(cp));
      omp_set_dynamic(0);
      cp.paired_end = in_vec.size() == 2;
      cp.preserve_order = !pairing_only;
      cp.preserve_id = !no_ids;
      cp.preserve_quality = !no_quality;
      cp.long_flag = false;
      cp.num_reads_per_block = spring::NUM_READS_PER_BLOCK;
      cp.num_reads_per_block_long = spring::NUM_READS_PER_BLOCK_LONG;
      cp.num_thr = num_thr;
      spring::preprocess(in_vec[0], cp.paired_end ? in_vec[1] : std::string(),
                         temp_given, cp, false, false);
      std::ofstream f(temp_given + "/cp_in.bin", std::ios::binary);
      f.write((char *)&cp, sizeof(cp));
      return 0;
    }
    if (hot_flag) {
      spring::compression_params cp;
      std::ifstream f(temp_given + "/cp_in.bin", std::ios::binary);
      f.read((char *)&cp, sizeof(cp));
      if (!f.good()) throw std::runtime_error("cannot read cp_in.bin");
      f.close();
      remove((temp_given + "/cp_in.bin").c_str());
      cp.num_thr = num_thr;
      omp_set_dynamic(0);
      double t0 = now_s();
      spring::call_reorder(temp_given, cp);
      double t1 = now_s();
      spring::call_encoder(temp_given, cp);
      double t2 = now_s();
      printf("HOTPATH_SECONDS %.6f %.6f\n", t1 - t0, t2 - t1);
      if (unbsc)
        for (int t = 0; t < cp.num_thr; t++) {
          std::string b = temp_given + "/read_seq.bin." + std::to_string(t);
          spring::bsc::BSC_decompress((b + ".bsc").c_str(), b.c_str());
        }
      return 0;
    }
    if (bscdec_flag) {
      for (const std::string &f : in_vec) {
        if (f.size() < 5 || f.substr(f.size() - 4) != ".bsc") throw std::runtime_error("not a .bsc file: " + f);
        spring::bsc::BSC_decompress(f.c_str(), f.substr(0, f.size() - 4).c_str());
      }
      return 0;
    }
    if (reblock_flag) {
      spring::compression_params cp;
      std::ifstream f(temp_given + "/cp_in.bin", std::ios::binary);
      f.read((char *)&cp, sizeof(cp));
      if (!f.good()) throw std::runtime_error("cannot read cp_in.bin");
      f.close();
      remove((temp_given + "/cp_in.bin").c_str());
      cp.num_thr = num_thr;
      omp_set_dynamic(0);
      if (!cp.preserve_order && cp.paired_end) spring::pe_encode(temp_given, cp);
      spring::reorder_compress_streams(temp_given, cp);
      const uint64_t units = cp.paired_end ? cp.num_reads / 2 : cp.num_reads;
      const uint64_t nb = (units + cp.num_reads_per_block - 1) / cp.num_reads_per_block;
      const char *names[] = {"read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin", "read_rev.txt",
                             "read_unaligned.txt", "read_lengths.bin", "read_pos_pair.bin", "read_rev_pair.txt"};
      for (uint64_t b = 0; b < nb; b++)
        for (int s = 0; s < (cp.paired_end ? 9 : 7); s++) {
          std::string x = temp_given + "/" + names[s] + "." + std::to_string(b);
          spring::bsc::BSC_decompress((x + ".bsc").c_str(), x.c_str());
        }
      return 0;
    }
    if (compress_flag == decompress_flag) {
      std::cerr << "Exactly one of -c / -d / --preprocess / --hotpath\n";
      return 1;
    }
    std::string temp_dir;
    while (true) {
      temp_dir = working_dir + "/tmp." + spring::random_string(10) + "/";
      if (!boost::filesystem::exists(temp_dir) &&
          boost::filesystem::create_directory(temp_dir))
        break;
    }
    try {
      if (compress_flag)
        spring::compress(temp_dir, in_vec, out_vec, num_thr, pairing_only,
                         no_quality, no_ids, quality_opts, false, false, false);
      else
        spring::decompress(temp_dir, in_vec, out_vec, num_thr, range_vec,
                           false, 6);
    } catch (...) {
      boost::filesystem::remove_all(temp_dir);
      throw;
    }
    boost::filesystem::remove_all(temp_dir);
  } catch (std::runtime_error &e) {
    std::cout << "Program terminated unexpectedly with error: " << e.what()
              << "\n";
    return 1;
  }
  return 0;
}
