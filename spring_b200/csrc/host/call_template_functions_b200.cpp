// call_template_functions_b200.cpp -- the reference's own interface for this path, backed by
// libspring_b200.so.
//
// Drop-in replacement for src/call_template_functions.cpp of shubhamchandak94/Spring: same two
// symbols, same signatures (src/call_template_functions.h:9-11), called from the same two places
// (src/spring.cpp:153 and :166).  Compile this file INSTEAD of call_template_functions.cpp and link
// libspring_b200.so; nothing else in the reference changes.
//
//   call_reorder : the whole GPU job (reorder_main + the encoder's stream generation).  The
//                  reorder -> encoder hand-off files of the reference (temp.dna.<t>, temppos.txt.<t>,
//                  tempflag.txt.<t>, read_order.bin.<t>, ...) are private to the two stages, so that
//                  hand-off stays in HBM and never touches the disk.
//   call_encoder : what is left of encoder_main on the host: pack_compress_seq's BSC step
//                  (src/encoder.cpp:148-153) on every read_seq.bin.<t>, in parallel like the
//                  reference (#pragma omp parallel, src/encoder.cpp:112).
//
// Errors: the C ABI returns codes; they are turned into std::runtime_error here so that the
// reference's catch blocks in main (src/main.cpp:151-166) clean up the temp dir as before.
#include <omp.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "call_template_functions.h"  // reference header: declares the two functions + compression_params
#include "libbsc/bsc.h"               // reference libbsc (stays on the host)
#include "spring_b200.h"

namespace spring {

static_assert(sizeof(compression_params) == sizeof(spring_b200_cp), "cp.bin layout mismatch");

namespace {
spring_b200_stats g_stats{};  // of the call_reorder that just ran (call_encoder prints the encoder's lines from it)
int env_int(const char *name, int dflt) {
  const char *v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
}  // namespace

void call_reorder(const std::string &temp_dir, compression_params &cp) {
  if (cp.max_readlen > 511) throw std::runtime_error("Wrong bitset size.");  // call_template_functions.cpp:61
  // the process-wide context: CUDA initialisation and buffers are shared with call_encoder and
  // reorder_compress_streams, and the streams stay in HBM for the re-blocking stage
  spring_b200_ctx *ctx = nullptr;
  const int device = env_int("SPRING_B200_DEVICE", 0);
  if (spring_b200_shared_ctx(device, &ctx) != SPRING_B200_OK)
    throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(nullptr));
  spring_b200_cp c;
  std::memcpy(&c, &cp, sizeof(c));
  spring_b200_stats st;
  const int gpus = env_int("SPRING_B200_GPUS", 1);
  if (gpus > 1) {  // SPRING_B200_GPUS=N: devices 0..N-1, one exchange over NVLink, shards merged before the files are written
    char err[512] = "";
    if (spring_b200_reorder_encode_files_multi(temp_dir.c_str(), &c, gpus, nullptr, (uint32_t)env_int("SPRING_B200_CHAINS", 0), &st, err,
                                               sizeof(err)) != SPRING_B200_OK)
      throw std::runtime_error(std::string("spring_b200: ") + err);
    g_stats = st;
  } else {
    // reads packed on the GPU by preprocess_b200.cpp are still in HBM: no .dna files to read
    const uint32_t chains = (uint32_t)env_int("SPRING_B200_CHAINS", 0);
    const int rc = spring_b200_packed_pending(ctx) ? spring_b200_reorder_encode_packed(ctx, temp_dir.c_str(), &c, chains)
                                                   : spring_b200_reorder_encode_files(ctx, temp_dir.c_str(), &c, chains);
    if (rc != SPRING_B200_OK) throw std::runtime_error(std::string("spring_b200: ") + spring_b200_last_error(ctx));
    spring_b200_get_stats(ctx, &st);
    g_stats = st;
  }
  std::printf("Reordering done, %u were unmatched\n", st.unmatched);  // reorder.h:633-635
}

void call_encoder(const std::string &temp_dir, compression_params &cp) {
  // pack_compress_seq (encoder.cpp:111-156): read_seq.bin.<t> is already packed 2 bits/base and
  // read_seq.bin.<t>.tail written; BSC-compress and remove the packed file.
  omp_set_num_threads(cp.num_thr);
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < cp.num_thr; t++) {
    const std::string base = temp_dir + "/read_seq.bin." + std::to_string(t);
    bsc::BSC_compress(base.c_str(), (base + ".bsc").c_str());
    std::remove(base.c_str());
  }
  // encoder.h:490-492
  std::printf("Encoding done:\n%u singleton reads were aligned\n%u reads with N were aligned\n", g_stats.singletons_aligned,
              g_stats.n_reads_aligned);
}

}  // namespace spring
