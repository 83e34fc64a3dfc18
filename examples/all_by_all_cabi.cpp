// all_by_all_cabi.cpp — qluster's --writeOutFinalAllByAllComparison (clusterDown.cpp:727-805) on the C ABI, compiled:
// every final cluster against every larger one, the summary row of allByAllCompCounts.tab.txt and the per-event rows of
// allByAllComp.tab.txt (distances_.mismatches_, lowKmerMismatches_, alignmentGaps_) -- without an aligner object: one
// sdgpu_align_events call returns the comparisons and the event lists of all pairs.
//
// The program takes the clusters on the command line as  name1 seq1 name2 seq2 ...  (largest first, like the sorted
// cluster vector) and prints the two tables to stdout, separated by a line holding "--".
//
//   g++ -std=c++17 -Iinclude examples/all_by_all_cabi.cpp -Lseekdeep_b200 -lsdgpu -Wl,-rpath,$PWD/seekdeep_b200
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "sdgpu.h"

int main(int argc, char** argv) {
  if (argc < 5 || (argc - 1) % 2) {
    std::fprintf(stderr, "usage: all_by_all_cabi name1 seq1 name2 seq2 [...]\n");
    return 2;
  }
  static sdgpu_params gp;  // qluster's aligner: gap 5,1 / left 5,1 / right 0,0, match 2 / mismatch -2, Illumina quality thresholds
  gp.gaps = {5, 1, 5, 1, 0, 0, 5, 1, 0, 0};
  for (int a = 0; a < SDGPU_MAT_DIM; ++a)
    for (int b = 0; b < SDGPU_MAT_DIM; ++b) gp.scoreMat[a * SDGPU_MAT_DIM + b] = (a == b) ? 2 : -2;
  gp.primaryQual = 25, gp.secondaryQual = 20, gp.qualThresWindow = 2;
  sdgpu_ctx* ctx = nullptr;
  if (sdgpu_create(0, &gp, &ctx)) {
    std::fprintf(stderr, "sdgpu_create: %s\n", sdgpu_last_error(nullptr));
    return 1;
  }
  std::vector<std::string> names, seqs;
  for (int a = 1; a + 1 < argc; a += 2) {
    names.emplace_back(argv[a]);
    seqs.emplace_back(argv[a + 1]);
  }
  const uint32_t n = uint32_t(seqs.size());
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> offsets{0};
  std::vector<double> cnts;
  for (uint32_t s = 0; s < n; ++s) {
    bases.insert(bases.end(), seqs[s].begin(), seqs[s].end());
    for (size_t i = 0; i < seqs[s].size(); ++i) quals.push_back(uint8_t(i % 7 == 3 ? 12 : 35));  // (a low-quality base every 7th column)
    offsets.push_back(bases.size());
    cnts.push_back(double(10 * (n - s)));
  }
  // for (seq : reversed(clusters)) for (otherSeq : clusters) { if same: break; alignCacheGlobal(otherSeq, seq); profileAlignment(otherSeq, seq, true, true, false); }
  std::vector<uint32_t> refIdx, readIdx;
  for (uint32_t s = n; s-- > 0;)
    for (uint32_t o = 0; o < s; ++o) {
      refIdx.push_back(o);
      readIdx.push_back(s);
    }
  const uint32_t nPairs = uint32_t(refIdx.size()), cap = 64;
  std::vector<sdgpu_comparison> comp(nPairs);
  std::vector<sdgpu_event> events(size_t(nPairs) * cap);
  std::vector<uint32_t> nEvents(nPairs);
  if (sdgpu_upload_seqs(ctx, bases.data(), quals.data(), offsets.data(), cnts.data(), n) ||
      sdgpu_build_kmer_index(ctx, nullptr, 0, 9, 1, 1, 0, 0) ||
      sdgpu_align_events(ctx, refIdx.data(), readIdx.data(), nPairs, SDGPU_CHECK_KMER | SDGPU_USING_QUALITY, comp.data(), events.data(), cap,
                         nEvents.data())) {
    std::fprintf(stderr, "sdgpu: %s\n", sdgpu_last_error(ctx));
    return 1;
  }
  std::printf("seq1\tseq2\talnScore\tscore\tscoreHq\t1BaseIndel\t2BaseIndel\t>2BaseIndel\tlqMismatch\thqMismatch\tlowKmerMismatch\n");
  for (uint32_t p = 0; p < nPairs; ++p) {
    const sdgpu_comparison& c = comp[p];
    std::printf("%s\t%s\t%d\t%.17g\t%.17g\t%.17g\t%.17g\t%.17g\t%u\t%u\t%u\n", names[refIdx[p]].c_str(), names[readIdx[p]].c_str(), c.alnScore,
                c.eventBasedIdentity, c.eventBasedIdentityHq, c.oneBaseIndel, c.twoBaseIndel, c.largeBaseIndel, c.lqMismatches, c.hqMismatches,
                c.lowKmerMismatches);
  }
  std::printf("--\nseq1\tseq2\terror\tvalue1\tpos1\tquality1\tvalue2\tpos2\tquality2\tkMerFreqByPos\tisHighQual\n");
  for (uint32_t p = 0; p < nPairs; ++p) {
    if (nEvents[p] > cap) {
      std::fprintf(stderr, "pair %u has %u events, more than this example's %u slots\n", p, nEvents[p], cap);
      return 1;
    }
    const char *ref = names[refIdx[p]].c_str(), *read = names[readIdx[p]].c_str();
    const std::string &refSeq = seqs[refIdx[p]], &readSeq = seqs[readIdx[p]];
    for (int pass = 0; pass < 3; ++pass)  // the reference writes mismatches_, then lowKmerMismatches_, then alignmentGaps_
      for (uint32_t e = 0; e < nEvents[p]; ++e) {
        const sdgpu_event& ev = events[size_t(p) * cap + e];
        const bool lowK = ev.kind == SDGPU_EV_MISMATCH_LOWKMER;
        const bool mism = ev.kind == SDGPU_EV_MISMATCH_HQ || ev.kind == SDGPU_EV_MISMATCH_LQ;
        if ((pass == 0 && mism) || (pass == 1 && lowK)) {
          std::printf("%s\t%s\t%s\t%c\t%u\t%u\t%c\t%u\t%u\t%u\t%s\n", ref, read, lowK ? "lowKmerMismatch" : "mismatch", ev.refBase, ev.refPos, ev.refQual,
                      ev.seqBase, ev.seqPos, ev.seqQual, ev.kmerFreqSeq, ev.kind == SDGPU_EV_MISMATCH_LQ ? "false" : "true");
        } else if (pass == 2 && ev.kind >= SDGPU_EV_GAP_IN_REF && !(ev.kind & SDGPU_EV_END_GAP)) {  // alignmentGaps_ holds the counted gaps
          const bool inRef = (ev.kind & 7u) == SDGPU_EV_GAP_IN_REF;  // gap::ref_: the bases sit in the read
          // gapedSequence_ / qualities_: the `size` bases of the other sequence from seqPos (gap in the ref) or refPos (gap in the read)
          const uint32_t from = inRef ? ev.seqPos : ev.refPos;
          const std::string gapped = (inRef ? readSeq : refSeq).substr(from, ev.size);
          const uint64_t q0 = offsets[inRef ? readIdx[p] : refIdx[p]] + from;
          std::string qs;
          for (uint32_t x = 0; x < ev.size; ++x) qs += (x ? "," : "") + std::to_string(int(quals[q0 + x]));
          // (the reference's own column order for gap rows, clusterDown.cpp:789-800)
          std::printf("%s\t%s\tgap-%s\t%s\t%s\t%u\t%s\t%s\t%u\t\t\n", ref, read, inRef ? "insertion" : "deletion", inRef ? "" : gapped.c_str(),
                      inRef ? "" : qs.c_str(), ev.refPos, inRef ? gapped.c_str() : "", inRef ? qs.c_str() : "", ev.refPos);
        }
      }
  }
  sdgpu_destroy(ctx);
  return 0;
}
