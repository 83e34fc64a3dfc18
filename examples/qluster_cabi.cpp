// qluster_cabi.cpp — the C ABI used from C++, the reference's own language (level 1).
//
// Reads a packed sample (as written by tests/test_gpu_qluster.py::test_cxx_caller_matches_python_api:
// u32 nReads, u64 offsets[nReads + 1], bases, quals), runs the collapse with the `--illumina` default
// iterations (genIlluminaDefaultPars(100) == etc/SeekDeepParameterFiles/illumina_lkmer1:2-23) and prints
// one line per final cluster: count, chimeric flag, consensus sequence.
//
//   g++ -std=c++17 -Iinclude examples/qluster_cabi.cpp -Lseekdeep_b200 -lsdgpu -Wl,-rpath,$PWD/seekdeep_b200
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "sdgpu_qluster.h"

static void die(const char* what, const char* why) {
  std::fprintf(stderr, "%s: %s\n", what, why ? why : "?");
  std::exit(1);
}

int main(int argc, char** argv) {
  if (argc < 2) die("usage", "qluster_cabi sample.bin");
  std::FILE* f = std::fopen(argv[1], "rb");
  if (!f) die("open", argv[1]);
  uint32_t n = 0;
  if (std::fread(&n, 4, 1, f) != 1) die("read", "header");
  std::vector<uint64_t> offsets(size_t(n) + 1);
  if (std::fread(offsets.data(), 8, offsets.size(), f) != offsets.size()) die("read", "offsets");
  std::vector<uint8_t> bases(offsets[n]), quals(offsets[n]);
  if (std::fread(bases.data(), 1, bases.size(), f) != bases.size()) die("read", "bases");
  if (std::fread(quals.data(), 1, quals.size(), f) != quals.size()) die("read", "quals");
  std::fclose(f);

  // == aligner(maxSize, gapInfo_, scoring_, kMaps, qScorePars_, countEndGaps_, weighHomopolymers_),
  //    clusterDown.cpp:420-426, with qluster's defaults (clusterDownSetUp.cpp:285-286, 402-404)
  static sdgpu_params gp;  // 64 KB score matrix: keep it off the stack
  gp.gaps = {5, 1, 5, 1, 0, 0, 5, 1, 0, 0};
  for (int a = 0; a < SDGPU_MAT_DIM; ++a)
    for (int b = 0; b < SDGPU_MAT_DIM; ++b) gp.scoreMat[a * SDGPU_MAT_DIM + b] = (a == b) ? 2 : -2;
  gp.primaryQual = 25;
  gp.secondaryQual = 20;
  gp.qualThresWindow = 2;
  gp.countEndGaps = 0;
  gp.weighHomopolymers = 0;
  sdgpu_ctx* ctx = nullptr;
  if (sdgpu_create(0, &gp, &ctx)) die("sdgpu_create", sdgpu_last_error(nullptr));

  // illumina_lkmer1:2-23 — stopCheck:smallCutoff:1baseIndel:2baseIndel:>2baseIndel:HQ:LQ:LK
  std::vector<sdq_iter_par> iters;
  for (uint32_t small : {3u, 0u}) {
    iters.push_back({100, small, 1, 0, 0, 0, 0, 1, 0, 0, 0});
    for (int lq : {0, 1, 2, 3, 4, 5, 6, 7, 8, 8}) iters.push_back({100, small, 2, 0, 0, 0, double(lq), 1, 0, 0, 0});
  }
  sdq_opts o{};
  o.kLen = 9, o.runCutOff = 1, o.kmersByPosition = 1, o.expandKmerSize = 5, o.checkKmers = 1, o.smallReadSize = 18;
  o.markChimeras = 1, o.chiRunCutOff = 1, o.chiOverlapSizeCutoff = 5, o.chiPosSpacing = 1, o.chiParentFreqs = 2.0;
  o.chiOverlap = {0, 0, 2.0, 1.0, 0.99, 0.0, 0.0, 1.0, 0, 0, 0.0};

  sdq_result* res = nullptr;
  if (sdq_run(ctx, &o, iters.data(), uint32_t(iters.size()), iters.data(), uint32_t(iters.size()), bases.data(), quals.data(),
              offsets.data(), n, &res))
    die("sdq_run", sdq_last_error());
  const uint32_t nc = sdq_num_clusters(res);
  for (uint32_t c = 0; c < nc; ++c) {
    std::string seq(sdq_cluster_len(res, c), '\0');
    sdq_cluster_seq(res, c, reinterpret_cast<uint8_t*>(&seq[0]), nullptr);
    std::printf("%.0f\t%d\t%s\n", sdq_cluster_cnt(res, c), sdq_cluster_chimeric(res, c, nullptr, nullptr), seq.c_str());
  }
  sdq_free(res);
  sdgpu_destroy(ctx);
  return 0;
}
