// reorder_compress_streams_b200.cpp -- the reference's pe_encode and reorder_compress_streams,
// backed by libspring_b200.so.
//
// Drop-in replacement for src/pe_encode.cpp and src/reorder_compress_streams.cpp of
// shubhamchandak94/Spring: same symbols, same signatures (src/pe_encode.h, src/reorder_compress_streams.h),
// called from the same places (src/spring.cpp:193 and :206).  Compile this file INSTEAD of those two
// and link libspring_b200.so.
//
//   pe_encode                 : nothing left to do on the host.  It only rewrites read_order.bin, whose one
//                               remaining consumer is reorder_compress_streams (reorder_compress_quality_id
//                               runs before it, src/spring.cpp:179); the GPU re-blocking applies the same
//                               mapping on the device (spring_b200_reblock_files does pe_encode first when
//                               cp.paired_end && !cp.preserve_order).
//   reorder_compress_streams  : the GPU writes the raw per-block streams (what the reference writes to its
//                               temporary files a.<b> ... g.<b>, :201-361); the host runs bsc::BSC_compress on
//                               each, blocks spread over cp.num_thr threads, as :363-428 does.
#include <omp.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "libbsc/bsc.h"  // reference libbsc (stays on the host)
#include "pe_encode.h"
#include "reorder_compress_streams.h"
#include "spring_b200.h"
#include "util.h"

namespace spring {

static_assert(sizeof(compression_params) == sizeof(spring_b200_cp), "cp.bin layout mismatch");

void pe_encode(const std::string &, const compression_params &) {}

void reorder_compress_streams(const std::string &temp_dir, const compression_params &cp) {
  spring_b200_ctx *ctx = nullptr;
  const char *dev = std::getenv("SPRING_B200_DEVICE");
  if (spring_b200_shared_ctx(dev ? std::atoi(dev) : 0, &ctx) != SPRING_B200_OK)  // the streams call_reorder left in HBM are reused
    throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(nullptr));
  spring_b200_cp c;
  std::memcpy(&c, &cp, sizeof(c));
  const int rc = spring_b200_reblock_files(ctx, temp_dir.c_str(), &c);
  if (rc != SPRING_B200_OK) throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(ctx));

  static const char *const files[] = {"read_flag.txt", "read_pos.bin", "read_noise.txt", "read_noisepos.bin", "read_rev.txt",
                                      "read_unaligned.txt", "read_lengths.bin", "read_pos_pair.bin", "read_rev_pair.txt"};
  const int num_streams = cp.paired_end ? 9 : 7;
  const uint64_t units = cp.paired_end ? cp.num_reads / 2 : cp.num_reads;
  const int64_t num_blocks = (int64_t)((units + cp.num_reads_per_block - 1) / cp.num_reads_per_block);
  omp_set_num_threads(cp.num_thr);
#pragma omp parallel for schedule(static, 1)  // block b on thread b % num_thr, :204-207,:430
  for (int64_t b = 0; b < num_blocks; b++)
    for (int s = 0; s < num_streams; s++) {
      const std::string raw = temp_dir + "/" + files[s] + "." + std::to_string(b);
      bsc::BSC_compress(raw.c_str(), (raw + ".bsc").c_str());
      std::remove(raw.c_str());
    }
}

}  // namespace spring
