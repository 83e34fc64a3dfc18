// gapped_strings_cabi.cpp — INTEGRATION.md level 2, compiled: serving njhseq's stateful
// aligner::alignCacheGlobal(ref, read) -> alignObjectA_ / alignObjectB_ from a GPU batch.
//
// The C ABI returns an alignment as its gap runs (sdgpu_gap = njhseq's gapInfo(pos_, size_, gapInA_), in
// traceback order = descending positions); rearrangeFromGapList() below is njhseq's rearrangeGlobal:
// insert '-' (quality 0) at every recorded run.  The program aligns every (ref, read) pair given on the
// command line as  ref1 read1 ref2 read2 ...  and prints, per pair: score, gapped ref, gapped read.
//
//   g++ -std=c++17 -Iinclude examples/gapped_strings_cabi.cpp -Lseekdeep_b200 -lsdgpu -Wl,-rpath,$PWD/seekdeep_b200
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "sdgpu.h"

// alnInfoGlobal::gapInfos_ -> alignObjectA_.seqBase_ / alignObjectB_.seqBase_
static void rearrangeFromGapList(const std::string& ref, const std::string& read, const std::vector<uint8_t>& refQual,
                                 const std::vector<uint8_t>& readQual, const sdgpu_gap* gaps, uint32_t nGapRuns, std::string& alnA,
                                 std::string& alnB, std::vector<uint8_t>& qualA, std::vector<uint8_t>& qualB) {
  alnA = ref, alnB = read, qualA = refQual, qualB = readQual;
  for (uint32_t g = 0; g < nGapRuns; ++g) {  // descending positions: earlier insertions do not shift later ones
    std::string& s = gaps[g].gapInA ? alnA : alnB;
    std::vector<uint8_t>& q = gaps[g].gapInA ? qualA : qualB;
    s.insert(gaps[g].pos, gaps[g].size, '-');
    q.insert(q.begin() + gaps[g].pos, gaps[g].size, uint8_t(0));
  }
}

int main(int argc, char** argv) {
  if (argc < 3 || (argc - 1) % 2) {
    std::fprintf(stderr, "usage: gapped_strings_cabi ref1 read1 [ref2 read2 ...]\n");
    return 2;
  }
  static sdgpu_params gp;  // qluster's aligner: gap 5,1 / left 5,1 / right 0,0, match 2 / mismatch -2
  gp.gaps = {5, 1, 5, 1, 0, 0, 5, 1, 0, 0};
  for (int a = 0; a < SDGPU_MAT_DIM; ++a)
    for (int b = 0; b < SDGPU_MAT_DIM; ++b) gp.scoreMat[a * SDGPU_MAT_DIM + b] = (a == b) ? 2 : -2;
  gp.primaryQual = 25, gp.secondaryQual = 20, gp.qualThresWindow = 2;
  sdgpu_ctx* ctx = nullptr;
  if (sdgpu_create(0, &gp, &ctx)) {
    std::fprintf(stderr, "sdgpu_create: %s\n", sdgpu_last_error(nullptr));
    return 1;
  }
  std::vector<std::string> seqs(argv + 1, argv + argc);
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> offsets{0};
  for (const std::string& s : seqs) {
    bases.insert(bases.end(), s.begin(), s.end());
    quals.insert(quals.end(), s.size(), uint8_t(30));
    offsets.push_back(bases.size());
  }
  const uint32_t nPairs = uint32_t(seqs.size() / 2), gapCap = 16;
  std::vector<uint32_t> refIdx(nPairs), readIdx(nPairs);
  for (uint32_t p = 0; p < nPairs; ++p) refIdx[p] = 2 * p, readIdx[p] = 2 * p + 1;
  std::vector<sdgpu_comparison> comp(nPairs);
  std::vector<sdgpu_gap> gaps(size_t(nPairs) * gapCap);
  if (sdgpu_upload_seqs(ctx, bases.data(), quals.data(), offsets.data(), nullptr, uint32_t(seqs.size())) ||
      sdgpu_align_profile(ctx, refIdx.data(), readIdx.data(), nPairs, SDGPU_USING_QUALITY, comp.data(), gaps.data(), gapCap)) {
    std::fprintf(stderr, "sdgpu: %s\n", sdgpu_last_error(ctx));
    return 1;
  }
  for (uint32_t p = 0; p < nPairs; ++p) {
    if (comp[p].status & SDGPU_ST_GAPLIST_TRUNCATED) {
      std::fprintf(stderr, "pair %u: more than %u gap runs\n", p, gapCap);
      return 1;
    }
    std::string alnA, alnB;
    std::vector<uint8_t> qa, qb;
    rearrangeFromGapList(seqs[2 * p], seqs[2 * p + 1], std::vector<uint8_t>(seqs[2 * p].size(), 30),
                         std::vector<uint8_t>(seqs[2 * p + 1].size(), 30), &gaps[size_t(p) * gapCap], comp[p].nGapRuns, alnA, alnB, qa, qb);
    std::printf("%d\t%s\t%s\n", comp[p].alnScore, alnA.c_str(), alnB.c_str());
  }
  sdgpu_destroy(ctx);
  return 0;
}
