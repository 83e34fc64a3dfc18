// sd_oracle.cpp — CPU oracle for the qluster hot path.  TEST INFRASTRUCTURE ONLY (see sd_oracle.h).
//
// PARITY UNPINNED (no golden vectors in the reference; njhseq un-vendored).  Every function
// cites (a) the reference call site that fixes its interface and (b) the njhseq routine whose
// published algorithm it restates.  The restated rules are listed in DESIGN.md.
//
// Written for clarity, not speed: a full (lenA+1)x(lenB+1) matrix of 3 scores + 3 pointer
// chars per cell, pointer chars 'U','L','D','B' exactly as njhseq's alignCalc stores them.
// The CUDA path uses a different formulation (outgoing arg-max bits, wavefront strips), so
// agreement between the two is a real check.
#include "sd_oracle.h"

#include <algorithm>
#include <cmath>
#include <climits>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace sdo {

// ---------------------------------------------------------------- alignment (row a3)
struct Cell {
  int32_t upInherit, leftInherit, diagInherit;
  char upInheritPtr, leftInheritPtr, diagInheritPtr;
};

struct GapRun {
  uint32_t pos, size;
  bool gapInA;
};

struct AlnResult {
  int32_t score = 0;
  std::vector<GapRun> gaps;  // traceback order
};

// njhseq alignCalc::needleMaximum: 'B' marks an up/left tie; gaps win ties against diagonal.
static inline int32_t needleMaximum(int32_t u, int32_t l, int32_t d, char& p) {
  if (u == l && u == d) {
    p = 'B';
    return u;
  } else if (u >= l && u >= d) {
    if (u == l) {
      p = 'B';
    } else {
      p = 'U';
    }
    return u;
  } else if (l >= u && l >= d) {
    if (l == u) {
      p = 'B';
    } else {
      p = 'L';
    }
    return l;
  } else {
    p = 'D';
    return d;
  }
}

// njhseq alignCalc::runNeedleSave (reached from aligner::alignCacheGlobal; reference call sites
// clusterDown.cpp:694,739, ControlBencher.hpp:117).  A = ref = rows, B = read = columns.
static void runNeedleSave(const sdgpu_params& P, const uint8_t* A, uint32_t sizeA, const uint8_t* B,
                          uint32_t sizeB, std::vector<Cell>& m, AlnResult& out) {
  const sdgpu_gap_pars& g = P.gaps;
  const uint32_t lenA = sizeA + 1, lenB = sizeB + 1;
  m.resize(size_t(lenA) * lenB);
  auto at = [&](uint32_t i, uint32_t j) -> Cell& { return m[size_t(i) * lenB + j]; };
  at(0, 0) = Cell{0, 0, 0, '\0', '\0', '\0'};
  for (uint32_t i = 1; i < lenA; ++i) {  // first column: gaps in the query (read) on the left
    Cell& c = at(i, 0);
    c.upInherit = (i == 1) ? -g.gapLeftQueryOpen : at(i - 1, 0).upInherit - g.gapLeftQueryExtend;
    c.leftInherit = 0;
    c.diagInherit = 0;
    c.upInheritPtr = 'U';
    c.leftInheritPtr = '\0';
    c.diagInheritPtr = '\0';
  }
  for (uint32_t j = 1; j < lenB; ++j) {  // first row: gaps in the ref on the left
    Cell& c = at(0, j);
    c.upInherit = 0;
    c.leftInherit = (j == 1) ? -g.gapLeftRefOpen : at(0, j - 1).leftInherit - g.gapLeftRefExtend;
    c.diagInherit = 0;
    c.upInheritPtr = '\0';
    c.leftInheritPtr = 'L';
    c.diagInheritPtr = '\0';
  }
  for (uint32_t i = 1; i < lenA; ++i) {
    for (uint32_t j = 1; j < lenB; ++j) {
      char ptrFlag;
      Cell& c = at(i, j);
      const Cell& up = at(i - 1, j);
      const Cell& left = at(i, j - 1);
      const Cell& dg = at(i - 1, j - 1);
      const bool lastCol = (j == lenB - 1), lastRow = (i == lenA - 1);
      if (i == 1) {
        c.upInherit = up.leftInherit - (lastCol ? g.gapRightQueryOpen : g.gapOpen);
        c.upInheritPtr = 'L';
      } else {
        const int32_t go = lastCol ? g.gapRightQueryOpen : g.gapOpen;
        const int32_t ge = lastCol ? g.gapRightQueryExtend : g.gapExtend;
        c.upInherit = needleMaximum(up.upInherit - ge, up.leftInherit - go, up.diagInherit - go, ptrFlag);
        c.upInheritPtr = ptrFlag;
      }
      if (j == 1) {
        c.leftInherit = left.upInherit - (lastRow ? g.gapRightRefOpen : g.gapOpen);
        c.leftInheritPtr = 'U';
      } else {
        const int32_t go = lastRow ? g.gapRightRefOpen : g.gapOpen;
        const int32_t ge = lastRow ? g.gapRightRefExtend : g.gapExtend;
        c.leftInherit = needleMaximum(left.upInherit - go, left.leftInherit - ge, left.diagInherit - go, ptrFlag);
        c.leftInheritPtr = ptrFlag;
      }
      const int32_t match = P.scoreMat[size_t(A[i - 1] & 127) * SDGPU_MAT_DIM + (B[j - 1] & 127)];
      if (i == 1) {
        c.diagInherit = dg.leftInherit + match;
        c.diagInheritPtr = 'L';
      } else if (j == 1) {
        c.diagInherit = dg.upInherit + match;
        c.diagInheritPtr = 'U';
      } else {
        c.diagInherit = match + needleMaximum(dg.upInherit, dg.leftInherit, dg.diagInherit, ptrFlag);
        c.diagInheritPtr = ptrFlag;
      }
    }
  }
  // traceback
  int64_t icursor = lenA - 1, jcursor = lenB - 1;
  char tracerNext = ' ';
  const Cell& last = at(uint32_t(icursor), uint32_t(jcursor));
  out.score = needleMaximum(last.upInherit, last.leftInherit, last.diagInherit, tracerNext);
  out.gaps.clear();
  uint32_t gapBSize = 0, gapASize = 0;
  while (icursor != 0 || jcursor != 0) {
    const Cell& c = at(uint32_t(icursor), uint32_t(jcursor));
    if (tracerNext == 'U' || tracerNext == 'B') {  // up/left ambiguity resolves to 'up'
      ++gapBSize;
      tracerNext = c.upInheritPtr;
      if (tracerNext != 'U' && tracerNext != 'B') {
        out.gaps.push_back(GapRun{uint32_t(jcursor), gapBSize, false});
        gapBSize = 0;
      }
      --icursor;
    } else if (tracerNext == 'L') {
      ++gapASize;
      tracerNext = c.leftInheritPtr;
      if (tracerNext != 'L') {
        out.gaps.push_back(GapRun{uint32_t(icursor), gapASize, true});
        gapASize = 0;
      }
      --jcursor;
    } else {  // 'D'
      tracerNext = c.diagInheritPtr;
      --icursor;
      --jcursor;
    }
  }
  if ((tracerNext == 'U' || tracerNext == 'B') && gapBSize != 0) {
    out.gaps.push_back(GapRun{uint32_t(jcursor), gapBSize, false});
  } else if (tracerNext == 'L' && gapASize != 0) {
    out.gaps.push_back(GapRun{uint32_t(icursor), gapASize, true});
  }
}

struct Gapped {
  std::string seq;
  std::vector<uint8_t> qual;
};

// njhseq aligner::rearrangeGlobal / alnInfoGlobal: insert '-' (quality 0) at each recorded
// gap; the list is in descending position order so earlier insertions do not shift later ones.
static void rearrangeGlobal(const uint8_t* A, const uint8_t* qA, uint32_t la, const uint8_t* B,
                            const uint8_t* qB, uint32_t lb, const AlnResult& r, Gapped& ga, Gapped& gb) {
  ga.seq.assign(reinterpret_cast<const char*>(A), la);
  gb.seq.assign(reinterpret_cast<const char*>(B), lb);
  ga.qual.assign(la, 0);
  gb.qual.assign(lb, 0);
  if (qA) ga.qual.assign(qA, qA + la);
  if (qB) gb.qual.assign(qB, qB + lb);
  for (const GapRun& g : r.gaps) {
    Gapped& t = g.gapInA ? ga : gb;
    t.seq.insert(g.pos, g.size, '-');
    t.qual.insert(t.qual.begin() + g.pos, g.size, uint8_t(0));
  }
}

// ---------------------------------------------------------------- k-mers (rows a1, a2)
struct KMaps {
  uint32_t kLen = 0, runCutOff = 0;
  bool byPos = false;
  std::unordered_map<std::string, uint32_t> noPos;
  std::vector<std::unordered_map<std::string, uint32_t>> pos;
  uint32_t count(const std::string& kmer, uint32_t p) const {
    if (byPos) {
      if (p >= pos.size()) return 0;
      auto it = pos[p].find(kmer);
      return it == pos[p].end() ? 0 : it->second;
    }
    auto it = noPos.find(kmer);
    return it == noPos.end() ? 0 : it->second;
  }
};

// njhseq indexKmers (reference call sites clusterDown.cpp:412-417, 505-510).
static KMaps* indexKmers(const uint8_t* bases, const uint64_t* offsets, const double* cnts,
                         const uint32_t* seqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                         bool byPos, bool expandPos, uint32_t expandSize) {
  KMaps* km = new KMaps();
  km->kLen = kLen;
  km->runCutOff = runCutOff;
  km->byPos = byPos;
  for (uint32_t n = 0; n < nIdx; ++n) {
    const uint32_t s = seqIdx ? seqIdx[n] : n;
    const uint8_t* seq = bases + offsets[s];
    const uint32_t len = uint32_t(offsets[s + 1] - offsets[s]);
    const uint32_t w = uint32_t(std::llround(cnts ? cnts[s] : 1.0));
    for (uint32_t cursor = 0; cursor + kLen <= len; ++cursor) {
      std::string kmer(reinterpret_cast<const char*>(seq + cursor), kLen);
      km->noPos[kmer] += w;
      if (byPos) {
        uint32_t lo = cursor, hi = cursor;
        if (expandPos) {
          lo = cursor > expandSize ? cursor - expandSize : 0;
          hi = cursor + expandSize;
        }
        if (km->pos.size() <= hi) km->pos.resize(hi + 1);
        for (uint32_t p = lo; p <= hi; ++p) km->pos[p][kmer] += w;
      }
    }
  }
  return km;
}

// Start of the k-mer that covers base `p`, centred and clamped to the sequence.
static inline bool kmerStartFor(uint32_t p, uint32_t kLen, uint32_t len, uint32_t& start) {
  if (len < kLen || kLen == 0) return false;
  start = p >= kLen / 2 ? p - kLen / 2 : 0;
  if (start + kLen > len) start = len - kLen;
  return true;
}

static bool isLowFreqAt(const KMaps* km, const uint8_t* seq, uint32_t len, uint32_t p) {
  uint32_t start;
  if (!km || !kmerStartFor(p, km->kLen, len, start)) return false;
  std::string kmer(reinterpret_cast<const char*>(seq + start), km->kLen);
  return km->count(kmer, start) <= km->runCutOff;
}

// njhseq kmerInfo::compareKmers (reference call site extractorPairedEnd.cpp:820-824):
// shared = sum over k-mers of min(count in a, count in b); frac = shared / (min(len) - k + 1).
static void compareKmers(const uint8_t* a, uint32_t la, const uint8_t* b, uint32_t lb, uint32_t k,
                         uint32_t& shared, double& frac) {
  std::unordered_map<std::string, uint32_t> ka, kb;
  for (uint32_t i = 0; i + k <= la; ++i) ++ka[std::string(reinterpret_cast<const char*>(a + i), k)];
  for (uint32_t i = 0; i + k <= lb; ++i) ++kb[std::string(reinterpret_cast<const char*>(b + i), k)];
  shared = 0;
  for (const auto& kv : ka) {
    auto it = kb.find(kv.first);
    if (it != kb.end()) shared += std::min(kv.second, it->second);
  }
  const uint32_t minLen = std::min(la, lb);
  frac = (minLen >= k) ? shared / static_cast<double>(minLen - k + 1) : 0.0;
}

// ---------------------------------------------------------------- profileAlignment (rows a4, a5)
// njhseq seqInfo::checkQual: base strictly above primaryQual and every neighbour within
// qualThresWindow strictly above secondaryQual.  The trailing window reproduces njhseq's
// bound (the last base of the read is not inspected when the window reaches the end).
static bool checkQual(const uint8_t* q, uint32_t len, uint32_t pos, const sdgpu_params& P) {
  if (!SDGPU_QUAL_ABOVE(q[pos], P.primaryQual)) return false;
  const uint32_t out = P.qualThresWindow;
  const uint32_t lower = pos > out ? pos - out : 0;
  for (uint32_t i = lower; i < pos; ++i)
    if (!SDGPU_QUAL_ABOVE(q[i], P.secondaryQual)) return false;
  const uint32_t higher = SDGPU_QUAL_WINDOW_END(pos, out, len);
  for (uint32_t i = pos + 1; i < higher; ++i)
    if (!SDGPU_QUAL_ABOVE(q[i], P.secondaryQual)) return false;
  return true;
}

struct GapEvent {
  uint32_t startPos, size;
  std::string gapped;
};

static bool isHomopolymer(const std::string& s) {
  for (char c : s)
    if (c != s[0]) return false;
  return true;
}

// njhseq aligner::handleGapCountingInA / InB.  `withGap` is the aligned string holding the
// '-' run, `withBases` the other one.
static void handleGapCounting(const GapEvent& g, const std::string& withGap, const std::string& withBases,
                              bool weighHomopolymers, sdgpu_comparison& c) {
  double w = 1.0;
  if (weighHomopolymers && isHomopolymer(g.gapped)) {
    const char base = g.gapped[0];
    auto flank = [&](const std::string& s) {
      double n = 0;
      for (size_t cur = size_t(g.startPos) + g.size; cur < s.size() && s[cur] == base; ++cur) ++n;
      for (size_t cur = 1; cur <= g.startPos && s[g.startPos - cur] == base; ++cur) ++n;
      return n;
    };
    const double gapSide = flank(withGap);
    const double baseSide = flank(withBases) + g.size;
    w = std::fabs(gapSide - baseSide) / (gapSide + baseSide);
  }
  if (g.size >= 3) {
    c.largeBaseIndel += w;
  } else if (g.size == 2) {
    c.twoBaseIndel += w;
  } else {
    c.oneBaseIndel += w;
  }
}

// njhseq aligner::profileAlignment (reference call sites clusterDown.cpp:696,741;
// ControlBencher.hpp:118; PairedReadProcessor.cpp:336) + comparison::setEventBaseIdentity.
// The (start, stop) column window and the per-mismatch records (`mismatches_`: alignment column,
// read base position, class) are what collapser::markChimeras asks of it.
struct MismatchEv {
  uint32_t alnPos, seqBasePos;
  int cls;  // 1 = high quality, 2 = low quality, 3 = low k-mer
};
static void profileAlignment(const sdgpu_params& P, const KMaps* km, const uint8_t* A, const uint8_t* qA,
                             uint32_t la, const uint8_t* B, const uint8_t* qB, uint32_t lb, const Gapped& ga,
                             const Gapped& gb, uint32_t flags, sdgpu_comparison& c, uint32_t start = 0,
                             uint32_t stop = UINT32_MAX, std::vector<MismatchEv>* events = nullptr) {
  const bool checkKmer = flags & SDGPU_CHECK_KMER, usingQuality = flags & SDGPU_USING_QUALITY,
             matchQuality = flags & SDGPU_MATCH_QUALITY;
  const std::string& sa = ga.seq;
  const std::string& sb = gb.seq;
  const uint32_t len = uint32_t(sa.size());
  uint32_t firstOffset = 0, secondOffset = 0;
  uint32_t endA = 0, regA = 0, endB = 0, regB = 0;
  std::vector<uint8_t> zerosA, zerosB;
  if (!qA) { zerosA.assign(la, 0); qA = zerosA.data(); }
  if (!qB) { zerosB.assign(lb, 0); qB = zerosB.data(); }
  for (uint32_t i = 0; i < len; ++i) {
    if (sa[i] == '-' || sb[i] == '-') {
      const bool inA = sa[i] == '-';
      const std::string& gs = inA ? sa : sb;
      const std::string& os = inA ? sb : sa;
      uint32_t& off = inA ? firstOffset : secondOffset;
      GapEvent g{i, 1, std::string(1, os[i])};
      ++off;
      while (i + 1 < len && gs[i + 1] == '-') {
        g.gapped.push_back(os[i + 1]);
        ++g.size;
        ++i;
        ++off;
      }
      if (g.startPos < start || g.startPos >= stop) continue;
      const bool endGap = (g.startPos + g.size >= len) || g.startPos == 0;
      if (endGap) {
        (inA ? endA : endB) += g.size;
        if (P.countEndGaps) {
          handleGapCounting(g, gs, os, P.weighHomopolymers, c);
          ++c.numGaps;
        }
      } else {
        (inA ? regA : regB) += g.size;
        handleGapCounting(g, gs, os, P.weighHomopolymers, c);
        ++c.numGaps;
      }
      continue;
    }
    if (i < start || i >= stop) continue;
    const uint32_t pa = i - firstOffset, pb = i - secondOffset;
    if (P.scoreMat[size_t(uint8_t(sa[i]) & 127) * SDGPU_MAT_DIM + (uint8_t(sb[i]) & 127)] < 0) {
      bool hq = true;
      if (usingQuality) hq = checkQual(qA, la, pa, P) && checkQual(qB, lb, pb, P);
      int cls;
      if (!hq) {
        ++c.lqMismatches;
        cls = 2;
      } else if (checkKmer && ((SDGPU_POLICY_LOWKMER_EITHER_SIDE && isLowFreqAt(km, A, la, pa)) || isLowFreqAt(km, B, lb, pb))) {
        ++c.lowKmerMismatches;
        cls = 3;
      } else {
        ++c.hqMismatches;
        cls = 1;
      }
      if (events) events->push_back(MismatchEv{i, pb, cls});
    } else {
      if (matchQuality && !(checkQual(qA, la, pa, P) && checkQual(qB, lb, pb, P))) {
        ++c.lowQualityMatches;
      } else {
        ++c.highQualityMatches;
      }
    }
  }
  const double coverageArea = double(len) - endA - endB;
  c.alnLen = len;
  c.refCoverage = la ? (coverageArea - regA) / double(la) : 0.0;
  c.queryCoverage = lb ? (coverageArea - regB) / double(lb) : 0.0;
  c.basesInAln = uint32_t(len - endA - endB - regA - regB);
  c.overLappingEvents = c.highQualityMatches + c.lowQualityMatches + c.hqMismatches + c.lqMismatches +
                        c.lowKmerMismatches + c.numGaps;
  if (c.overLappingEvents == 0) {
    c.eventBasedIdentity = 0;
    c.eventBasedIdentityHq = 0;
  } else {
    c.eventBasedIdentity = (c.highQualityMatches + c.lowQualityMatches) / double(c.overLappingEvents);
    c.eventBasedIdentityHq =
        (c.highQualityMatches + c.lowQualityMatches + c.lqMismatches + c.lowKmerMismatches) /
        double(c.overLappingEvents);
  }
}

static void alignProfile(const sdgpu_params& P, const KMaps* km, const uint8_t* A, const uint8_t* qA,
                         uint32_t la, const uint8_t* B, const uint8_t* qB, uint32_t lb, uint32_t flags,
                         std::vector<Cell>& mat, AlnResult& aln, Gapped& ga, Gapped& gb, sdgpu_comparison& c) {
  std::memset(&c, 0, sizeof(c));
  runNeedleSave(P, A, la, B, lb, mat, aln);
  rearrangeGlobal(A, qA, la, B, qB, lb, aln, ga, gb);
  c.alnScore = aln.score;
  c.nGapRuns = uint32_t(aln.gaps.size());
  profileAlignment(P, km, A, qA, la, B, qB, lb, ga, gb, flags, c);
}

// ---------------------------------------------------------------- qluster host loop (rows a6, a7)
struct Uniq {  // one exact-unique input sequence (clusterDown.cpp:272-288)
  std::string seq;
  std::vector<uint8_t> qual;
  double cnt = 0;
  std::vector<uint32_t> inputIdx;
};

struct Clus {
  std::string seq;
  std::vector<uint8_t> qual;
  double cnt = 0;
  uint32_t firstId = 0;  // stands for cluster::firstReadName_ (key of previousErrorChecks_)
  std::vector<uint32_t> members;  // unique ids (cluster::reads_)
  bool remove = false, needCons = false;
  double avgErr = 0;
  std::unordered_map<uint32_t, sdgpu_comparison> prev;
  bool chimeric = false;  // seqBase_.markAsChimeric()
  uint32_t chiParent[2] = {0, 0}, chiPos[2] = {0, 0};
};

static double averageErrorRate(const std::vector<uint8_t>& q) {
  if (q.empty()) return 0;
  double s = 0;
  for (uint8_t v : q) s += std::pow(10.0, -(double(v) / 10.0));
  return s / double(q.size());
}

// readVecSorter "totalCount": count desc, then average error asc; made total with seq, firstId.
static bool clusLess(const Clus& a, const Clus& b) {
  if (a.cnt != b.cnt) return a.cnt > b.cnt;
  if (a.avgErr != b.avgErr) return a.avgErr < b.avgErr;
  if (a.seq != b.seq) return a.seq < b.seq;
  return a.firstId < b.firstId;
}

// comparison::passErrorProfile (reference: ControlBencher.hpp:120; par columns
// etc/SeekDeepParameterFiles/illumina_lkmer1:1) and the OTU identity test (clusterDownSetUp.cpp:289-312).
static bool passErrorCheck(const sdo_iter_par& it, const sdgpu_comparison& g) {
  if (it.onPerId) {
    return (it.onHqPerId ? g.eventBasedIdentityHq : g.eventBasedIdentity) >= it.perId;
  }
  return it.oneBaseIndel >= g.oneBaseIndel && it.twoBaseIndel >= g.twoBaseIndel &&
         it.largeBaseIndel >= g.largeBaseIndel && it.hqMismatches >= double(g.hqMismatches) &&
         it.lqMismatches >= double(g.lqMismatches) && it.lowKmerMismatches >= double(g.lowKmerMismatches);
}

// The per-(parent, candidate) step of collapser::markChimeras (call site clusterDown.cpp:576,
// options clusterDownSetUp.cpp:372-381): global alignment, errorProfile with mismatch records,
// low-quality mismatches dropped unless keepLowQaulityMismatches_, then the columns before the
// first and after the last remaining mismatch re-profiled and judged against chiOverlap_.
static void chimeraProbe(const sdgpu_params& P, const KMaps* km, const uint8_t* A, const uint8_t* qA, uint32_t la,
                         const uint8_t* B, const uint8_t* qB, uint32_t lb, uint32_t flags,
                         const sdo_iter_par& chiOverlap, bool keepLq, std::vector<Cell>& mat,
                         sdgpu_chimera_probe& out) {
  AlnResult aln;
  Gapped ga, gb;
  sdgpu_comparison c;
  std::memset(&c, 0, sizeof(c));
  std::memset(&out, 0, sizeof(out));
  runNeedleSave(P, A, la, B, lb, mat, aln);
  rearrangeGlobal(A, qA, la, B, qB, lb, aln, ga, gb);
  std::vector<MismatchEv> ev;
  profileAlignment(P, km, A, qA, la, B, qB, lb, ga, gb, flags, c, 0, UINT32_MAX, &ev);
  if (!keepLq) ev.erase(std::remove_if(ev.begin(), ev.end(), [](const MismatchEv& e) { return e.cls == 2; }), ev.end());
  out.alnScore = aln.score;
  out.keptMismatches = uint32_t(ev.size());
  out.numGaps = c.numGaps;
  if (ev.empty()) return;
  out.firstPos = ev.front().seqBasePos;
  out.lastPos = ev.back().seqBasePos;
  out.firstAlnPos = ev.front().alnPos;
  out.lastAlnPos = ev.back().alnPos;
  sdo_iter_par t = chiOverlap;
  t.onPerId = 0;
  sdgpu_comparison w;
  std::memset(&w, 0, sizeof(w));
  profileAlignment(P, km, A, qA, la, B, qB, lb, ga, gb, flags, w, 0, out.firstAlnPos);
  if (passErrorCheck(t, w)) out.flags |= SDGPU_CHI_FRONT_PASS;
  std::memset(&w, 0, sizeof(w));
  profileAlignment(P, km, A, qA, la, B, qB, lb, ga, gb, flags, w, out.lastAlnPos + 1, UINT32_MAX);
  if (passErrorCheck(t, w)) out.flags |= SDGPU_CHI_BACK_PASS;
}

struct CharCounter {  // njhseq charCounter restricted to what consensus building uses
  uint32_t cnt[128] = {0};
  uint32_t qual[128] = {0};
  void add(char b, uint32_t q, uint32_t w) {
    cnt[uint8_t(b) & 127] += w;
    qual[uint8_t(b) & 127] += q * w;
  }
  uint32_t total() const {
    uint32_t t = 0;
    for (uint32_t v : cnt) t += v;
    return t;
  }
  void best(char& b, uint32_t& q) const {  // first strictly-greatest in ASCII order
    uint32_t bc = 0;
    b = ' ';
    q = 0;
    for (int ch = 0; ch < 128; ++ch)
      if (cnt[ch] > bc) {
        bc = cnt[ch];
        b = char(ch);
        q = qual[ch];
      }
  }
};

struct Engine {
  const sdgpu_params& P;
  const sdo_qluster_opts& O;
  const KMaps* km = nullptr;
  std::vector<Uniq> uniq;
  std::vector<Cell> mat;
  uint64_t nAlign = 0, nCells = 0;
  std::unordered_map<std::string, AlnResult> cache;  // aligner::alnHolder_ (alignCacheGlobal)

  Engine(const sdgpu_params& p, const sdo_qluster_opts& o) : P(p), O(o) {}

  const AlnResult& alignCacheGlobal(const std::string& a, const std::string& b) {
    std::string key;
    key.reserve(a.size() + b.size() + 1);
    key.append(a).push_back('|');
    key.append(b);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    AlnResult r;
    runNeedleSave(P, reinterpret_cast<const uint8_t*>(a.data()), uint32_t(a.size()),
                  reinterpret_cast<const uint8_t*>(b.data()), uint32_t(b.size()), mat, r);
    ++nAlign;
    nCells += uint64_t(a.size()) * b.size();
    return cache.emplace(std::move(key), std::move(r)).first->second;
  }

  sdgpu_comparison compare(const Clus& ref, const Clus& read) {
    sdgpu_comparison c;
    std::memset(&c, 0, sizeof(c));
    const AlnResult& aln = alignCacheGlobal(ref.seq, read.seq);
    Gapped ga, gb;
    const uint8_t* A = reinterpret_cast<const uint8_t*>(ref.seq.data());
    const uint8_t* B = reinterpret_cast<const uint8_t*>(read.seq.data());
    rearrangeGlobal(A, ref.qual.data(), uint32_t(ref.seq.size()), B, read.qual.data(), uint32_t(read.seq.size()),
                    aln, ga, gb);
    c.alnScore = aln.score;
    c.nGapRuns = uint32_t(aln.gaps.size());
    const uint32_t flags = (O.checkKmers ? SDGPU_CHECK_KMER : 0u) | SDGPU_USING_QUALITY;
    profileAlignment(P, km, A, ref.qual.data(), uint32_t(ref.seq.size()), B, read.qual.data(),
                     uint32_t(read.seq.size()), ga, gb, flags, c);
    return c;
  }

  // njhseq cluster::compare: memoised on the pair of firstReadName_s.
  bool clusterCompare(Clus& clus, Clus& read, const sdo_iter_par& it) {
    auto a = clus.prev.find(read.firstId);
    auto b = read.prev.find(clus.firstId);
    if (a != clus.prev.end() && b != read.prev.end()) return passErrorCheck(it, a->second);
    sdgpu_comparison c = compare(clus, read);
    clus.prev[read.firstId] = c;
    read.prev[clus.firstId] = c;
    return passErrorCheck(it, c);
  }

  // njhseq baseCluster::calculateConsensus / ConsensusHelper::{increaseCounters,buildConsensus}.
  void calculateConsensus(Clus& cl) {
    cl.needCons = false;
    if (cl.members.size() <= 1) return;
    double totalLen = 0, totalCnt = 0;
    for (uint32_t u : cl.members) {
      totalLen += uniq[u].seq.size() * uniq[u].cnt;
      totalCnt += uniq[u].cnt;
    }
    const double avgLen = totalLen / totalCnt;
    if (std::fabs(double(cl.seq.size()) - avgLen) > 0.1 * avgLen) {  // restart from the longest member
      size_t biggest = 0;
      for (uint32_t u : cl.members)
        if (uniq[u].seq.size() > biggest) {
          biggest = uniq[u].seq.size();
          cl.seq = uniq[u].seq;
          cl.qual = uniq[u].qual;
        }
    }
    std::vector<CharCounter> counters(cl.seq.size());
    std::map<uint32_t, std::map<uint32_t, CharCounter>> insertions;
    std::map<uint32_t, CharCounter> beginningGap;
    for (uint32_t u : cl.members) {
      const Uniq& r = uniq[u];
      const uint32_t w = uint32_t(std::llround(r.cnt));
      const AlnResult& aln = alignCacheGlobal(cl.seq, r.seq);
      Gapped ga, gb;
      rearrangeGlobal(reinterpret_cast<const uint8_t*>(cl.seq.data()), cl.qual.data(), uint32_t(cl.seq.size()),
                      reinterpret_cast<const uint8_t*>(r.seq.data()), r.qual.data(), uint32_t(r.seq.size()), aln,
                      ga, gb);
      uint32_t offSet = 0, currentOffset = 1, start = 0;
      const uint32_t len = uint32_t(ga.seq.size());
      if (len && ga.seq[0] == '-') {
        while (start < len && ga.seq[start] == '-') ++start;
        for (uint32_t i = 0; i < start; ++i) beginningGap[i].add(gb.seq[i], gb.qual[i], w);
        offSet = start;
      }
      for (uint32_t i = start; i < len; ++i) {
        if (ga.seq[i] == '-') {
          insertions[i - offSet][currentOffset].add(gb.seq[i], gb.qual[i], w);
          ++currentOffset;
          ++offSet;
          continue;
        }
        currentOffset = 1;
        counters[i - offSet].add(gb.seq[i], gb.qual[i], w);
      }
    }
    std::string ns;
    std::vector<uint8_t> nq;
    const double fortyPercent = 0.40 * cl.cnt;
    const uint32_t size = uint32_t(std::llround(cl.cnt));
    for (const auto& bg : beginningGap) {
      char b;
      uint32_t q;
      bg.second.best(b, q);
      const uint32_t tot = bg.second.total();
      if (b == '-' || tot < fortyPercent) continue;
      ns.push_back(b);
      nq.push_back(uint8_t(q / tot));
    }
    for (uint32_t pos = 0; pos < counters.size(); ++pos) {
      auto ins = insertions.find(pos);
      if (ins != insertions.end()) {
        for (const auto& ci : ins->second) {
          char b;
          uint32_t q;
          ci.second.best(b, q);
          const uint32_t tot = ci.second.total();
          const uint32_t bc = ci.second.cnt[uint8_t(b) & 127];
          const uint32_t nonInserting = size > tot ? size - tot : 0;
          if (b == ' ' || bc <= nonInserting) continue;
          ns.push_back(b);
          nq.push_back(uint8_t(q / bc));
        }
      }
      char b;
      uint32_t q;
      counters[pos].best(b, q);
      const uint32_t tot = counters[pos].total();
      if (b == '-' || b == ' ' || tot < fortyPercent) continue;
      ns.push_back(b);
      nq.push_back(uint8_t(q / tot));
    }
    if (ns != cl.seq || nq != cl.qual) {
      cl.seq.swap(ns);
      cl.qual.swap(nq);
    }
    cl.avgErr = averageErrorRate(cl.qual);
    cl.prev.clear();
  }

  // njhseq collapser::findMatch + collapseWithParameters.
  void collapseWithParameters(std::vector<Clus>& cs, const sdo_iter_par& it) {
    size_t alive = 0;
    for (const Clus& c : cs) alive += !c.remove;
    if (alive < 2) return;
    size_t amountAdded = 0;
    for (size_t rp = cs.size(); rp-- > 0;) {
      Clus& read = cs[rp];
      if (read.remove) continue;
      uint32_t count = 0;
      for (size_t cp = 0; cp < cs.size(); ++cp) {
        Clus& clus = cs[cp];
        if (clus.remove) continue;
        if (clus.cnt <= double(it.smallCutoff)) continue;
        if (cp == rp) continue;
        ++count;
        if (count > it.stopCheck) break;
        if (clusterCompare(clus, read, it)) {
          clus.cnt += read.cnt;
          clus.members.insert(clus.members.end(), read.members.begin(), read.members.end());
          clus.needCons = true;
          read.remove = true;
          ++amountAdded;
          break;
        }
      }
    }
    if (amountAdded > 0) {
      for (Clus& c : cs)
        if (!c.remove && c.needCons) calculateConsensus(c);
    }
  }

  // njhseq collapser::runFullClustering (reference call sites clusterDown.cpp:488-518).
  // collapser::markChimeras(clusters, alignerObj, chiOpts_) (clusterDown.cpp:566-577).  `cs` is
  // sorted by total count; candidates start at the third cluster; a parent must be at least
  // parentFreqs_ times as abundant; a candidate is chimeric when one parent explains its front
  // up to a mismatch that lies behind the last mismatch against another parent explaining its back.
  void markChimeras(std::vector<Clus>& cs) {
    if (cs.size() <= 2) return;
    const uint32_t flags = SDGPU_USING_QUALITY;  // profileAlignment(parent, sub, false, true, false)
    for (size_t sub = 2; sub < cs.size(); ++sub) {
      Clus& cand = cs[sub];
      if (cand.cnt <= double(O.chiRunCutOff)) continue;
      std::multimap<uint32_t, uint32_t> front, back;  // read position of the mismatch -> parent
      for (size_t par = 0; par < sub; ++par) {
        const Clus& parent = cs[par];
        if (parent.cnt / cand.cnt < O.chiParentFreqs) continue;
        sdgpu_chimera_probe pr;
        chimeraProbe(P, nullptr, reinterpret_cast<const uint8_t*>(parent.seq.data()), parent.qual.data(),
                     uint32_t(parent.seq.size()), reinterpret_cast<const uint8_t*>(cand.seq.data()), cand.qual.data(),
                     uint32_t(cand.seq.size()), flags, O.chiOverlap, O.chiKeepLowQual != 0, mat, pr);
        ++nAlign;
        nCells += uint64_t(parent.seq.size()) * cand.seq.size();
        if (pr.keptMismatches == 0) continue;
        const uint32_t lb = uint32_t(cand.seq.size());
        if ((pr.flags & SDGPU_CHI_FRONT_PASS) && pr.firstPos >= O.chiOverlapSizeCutoff) front.emplace(pr.firstPos, uint32_t(par));
        if ((pr.flags & SDGPU_CHI_BACK_PASS) && lb - 1 - pr.lastPos >= O.chiOverlapSizeCutoff) back.emplace(pr.lastPos, uint32_t(par));
      }
      for (const auto& f : front) {
        for (const auto& b : back) {
          if (cand.chimeric) break;
          if (f.second != b.second && f.first > b.first && f.first - b.first >= O.chiPosSpacing) {
            cand.chimeric = true;
            cand.chiParent[0] = f.second, cand.chiPos[0] = f.first;
            cand.chiParent[1] = b.second, cand.chiPos[1] = b.first;
          }
        }
      }
    }
  }

  // njhseq collapser::runFullClustering with kmerBinOpts_.useKmerBinning_ (clusterDownSetUp.cpp:352-354; restated from
  // njhseq's published greedyKmerSimCluster / kmerCluster::compareRead, no call site in /root/reference shows it):
  // in list order, a cluster joins the first bin whose first cluster shares MORE THAN kmerCutOff of its k-mers
  // (kmerInfo::compareKmers(...).second > cutOff), else it opens a bin; every bin of two or more runs through
  // binIteratorMap on its own; the survivors, bin after bin, are the list the run's own iterations start from.
  // kmerInfo of one sequence as a sorted list of its k-mers (with repeats): two such lists share
  // sum over k-mers of min(count in a, count in b) entries -- compareKmers' first value -- by one merge.
  static std::vector<std::string> kmerList(const std::string& seq, uint32_t k) {
    std::vector<std::string> v;
    for (uint32_t i = 0; i + k <= seq.size(); ++i) v.emplace_back(seq, i, k);
    std::sort(v.begin(), v.end());
    return v;
  }
  static double sharedKmerFraction(const std::vector<std::string>& a, size_t lenA, const std::vector<std::string>& b, size_t lenB, uint32_t k) {
    uint32_t shared = 0;
    for (size_t i = 0, j = 0; i < a.size() && j < b.size();) {
      const int c = a[i].compare(b[j]);
      if (c == 0) {
        ++shared;
        ++i;
        ++j;
      } else if (c < 0) {
        ++i;
      } else {
        ++j;
      }
    }
    const size_t minLen = std::min(lenA, lenB);
    return minLen >= k ? shared / static_cast<double>(minLen - k + 1) : 0.0;  // compareKmers' second value
  }

  void runFullClustering(std::vector<Clus>& cs, const sdo_iter_par* its, uint32_t nIts) {
    if (O.useKmerBinning) {
      std::vector<std::vector<Clus>> bins;
      std::vector<std::vector<std::string>> seedKmers;  // kmerInfo of every bin's first cluster
      for (Clus& c : cs) {
        if (c.remove) continue;
        const std::vector<std::string> mine = kmerList(c.seq, O.kCompareLen);
        bool placed = false;
        for (size_t bx = 0; bx < bins.size(); ++bx) {
          const double frac = sharedKmerFraction(seedKmers[bx], bins[bx][0].seq.size(), mine, c.seq.size(), O.kCompareLen);
          if (SDGPU_KMER_BIN_PASSES(frac, O.kmerCutOff)) {
            bins[bx].push_back(std::move(c));
            placed = true;
            break;
          }
        }
        if (!placed) {
          bins.emplace_back();
          bins.back().push_back(std::move(c));
          seedKmers.push_back(mine);
        }
      }
      cs.clear();
      const sdo_iter_par* binIts = O.binPars ? O.binPars : its;
      const uint32_t nBinIts = O.binPars ? uint32_t(O.nBinPars) : nIts;
      for (auto& bin : bins) {
        if (bin.size() > 1) runClustering(bin, binIts, nBinIts);
        for (Clus& c : bin) cs.push_back(std::move(c));
      }
    }
    runClustering(cs, its, nIts);
  }

  void runClustering(std::vector<Clus>& cs, const sdo_iter_par* its, uint32_t nIts) {
    for (uint32_t i = 0; i < nIts; ++i) {
      cs.erase(std::remove_if(cs.begin(), cs.end(), [](const Clus& c) { return c.remove; }), cs.end());
      std::sort(cs.begin(), cs.end(), clusLess);
      if (cs.size() < 2) break;
      collapseWithParameters(cs, its[i]);
      // --converge (clusterDownSetUp.cpp:394, "Keep clustering at each iteration until there is no more collapsing"):
      // the same line again, on the re-sorted survivors, while it still merges something
      while (O.converge) {
        const size_t before = cs.size();
        cs.erase(std::remove_if(cs.begin(), cs.end(), [](const Clus& c) { return c.remove; }), cs.end());
        if (cs.size() == before || cs.size() < 2) break;
        std::sort(cs.begin(), cs.end(), clusLess);
        collapseWithParameters(cs, its[i]);
      }
    }
    cs.erase(std::remove_if(cs.begin(), cs.end(), [](const Clus& c) { return c.remove; }), cs.end());
    std::sort(cs.begin(), cs.end(), clusLess);
  }
};

}  // namespace sdo

using namespace sdo;

struct sdo_kmaps {
  KMaps* km;
};

struct sdo_qluster_result {
  std::vector<Clus> clusters;
  std::vector<uint32_t> membership;
  uint32_t nUnique = 0;
  uint64_t nAlign = 0, nCells = 0;
};

static void copyGaps(const AlnResult& r, sdgpu_gap* gaps, uint32_t gapCap) {
  if (!gaps) return;
  for (uint32_t i = 0; i < r.gaps.size() && i < gapCap; ++i)
    gaps[i] = sdgpu_gap{r.gaps[i].pos, r.gaps[i].size, r.gaps[i].gapInA ? 1u : 0u};
}

extern "C" {

int sdo_align_global(const sdgpu_params* p, const uint8_t* a, uint32_t lenA, const uint8_t* b, uint32_t lenB,
                     int32_t* scoreOut, sdgpu_gap* gaps, uint32_t gapCap) {
  if (!p || !a || !b || lenA == 0 || lenB == 0) return -1;
  std::vector<Cell> mat;
  AlnResult r;
  runNeedleSave(*p, a, lenA, b, lenB, mat, r);
  if (scoreOut) *scoreOut = r.score;
  copyGaps(r, gaps, gapCap);
  return int(r.gaps.size());
}

int sdo_align_profile(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* a, const uint8_t* qa,
                      uint32_t lenA, const uint8_t* b, const uint8_t* qb, uint32_t lenB, uint32_t flags,
                      sdgpu_comparison* out, sdgpu_gap* gaps, uint32_t gapCap, char* alnA, char* alnB) {
  if (!p || !a || !b || !out || lenA == 0 || lenB == 0) return SDGPU_E_ARG;
  std::vector<Cell> mat;
  AlnResult r;
  Gapped ga, gb;
  alignProfile(*p, kmaps ? kmaps->km : nullptr, a, qa, lenA, b, qb, lenB, flags, mat, r, ga, gb, *out);
  copyGaps(r, gaps, gapCap);
  if (gaps && r.gaps.size() > gapCap) out->status |= SDGPU_ST_GAPLIST_TRUNCATED;
  if (alnA) std::strcpy(alnA, ga.seq.c_str());
  if (alnB) std::strcpy(alnB, gb.seq.c_str());
  return SDGPU_OK;
}

// The per-event maps of aligner::comp_ after profileAlignment -- distances_.mismatches_, lowKmerMismatches_,
// alignmentGaps_ (dumped by clusterDown.cpp:759-800) -- listed in alignment-column order.  Walks the gapped strings
// the way profileAlignment does; returns the number of events (may exceed cap).
int sdo_align_events(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* a, const uint8_t* qa, uint32_t lenA,
                     const uint8_t* b, const uint8_t* qb, uint32_t lenB, uint32_t flags, sdgpu_comparison* cmpOut,
                     sdgpu_event* out, uint32_t cap) {
  if (!p || !a || !b || lenA == 0 || lenB == 0) return -1;
  const KMaps* km = kmaps ? kmaps->km : nullptr;
  std::vector<Cell> mat;
  AlnResult r;
  Gapped ga, gb;
  sdgpu_comparison c;
  alignProfile(*p, km, a, qa, lenA, b, qb, lenB, flags, mat, r, ga, gb, c);
  if (cmpOut) *cmpOut = c;
  const bool checkKmer = flags & SDGPU_CHECK_KMER, usingQuality = flags & SDGPU_USING_QUALITY;
  const std::string &sa = ga.seq, &sb = gb.seq;
  const uint32_t len = uint32_t(sa.size());
  uint32_t n = 0, pa = 0, pb = 0;
  auto emit = [&](const sdgpu_event& e) {
    if (n < cap && out) out[n] = e;
    ++n;
  };
  auto kcount = [&](const uint8_t* seq, uint32_t l, uint32_t pos) -> uint32_t {
    uint32_t start;
    if (!km || !kmerStartFor(pos, km->kLen, l, start)) return 0;
    return km->count(std::string(reinterpret_cast<const char*>(seq + start), km->kLen), start);
  };
  for (uint32_t i = 0; i < len; ++i) {
    if (sa[i] == '-' || sb[i] == '-') {
      const bool inA = sa[i] == '-';
      const std::string& gs = inA ? sa : sb;
      uint32_t size = 1;
      while (i + size < len && gs[i + size] == '-') ++size;
      const bool endGap = (i + size >= len) || i == 0;
      sdgpu_event e;
      std::memset(&e, 0, sizeof(e));
      e.kind = (inA ? SDGPU_EV_GAP_IN_REF : SDGPU_EV_GAP_IN_READ) | ((endGap && !p->countEndGaps) ? SDGPU_EV_END_GAP : 0u);
      e.alnPos = i;
      e.refPos = pa;
      e.seqPos = pb;
      e.size = size;
      if (inA) {
        e.seqBase = uint8_t(sb[i]);
        e.seqQual = qb ? qb[pb] : 0;
        pb += size;
      } else {
        e.refBase = uint8_t(sa[i]);
        e.refQual = qa ? qa[pa] : 0;
        pa += size;
      }
      emit(e);
      i += size - 1;
      continue;
    }
    if (p->scoreMat[size_t(uint8_t(sa[i]) & 127) * SDGPU_MAT_DIM + (uint8_t(sb[i]) & 127)] < 0) {
      bool hq = true;
      std::vector<uint8_t> zA, zB;
      const uint8_t* QA = qa;
      const uint8_t* QB = qb;
      if (!QA) { zA.assign(lenA, 0); QA = zA.data(); }
      if (!QB) { zB.assign(lenB, 0); QB = zB.data(); }
      if (usingQuality) hq = checkQual(QA, lenA, pa, *p) && checkQual(QB, lenB, pb, *p);
      sdgpu_event e;
      std::memset(&e, 0, sizeof(e));
      if (!hq) {
        e.kind = SDGPU_EV_MISMATCH_LQ;
      } else if (checkKmer && ((SDGPU_POLICY_LOWKMER_EITHER_SIDE && isLowFreqAt(km, a, lenA, pa)) || isLowFreqAt(km, b, lenB, pb))) {
        e.kind = SDGPU_EV_MISMATCH_LOWKMER;
      } else {
        e.kind = SDGPU_EV_MISMATCH_HQ;
      }
      e.alnPos = i;
      e.refPos = pa;
      e.seqPos = pb;
      e.size = 1;
      if (checkKmer) {
        e.kmerFreqRef = kcount(a, lenA, pa);
        e.kmerFreqSeq = kcount(b, lenB, pb);
      }
      e.refBase = uint8_t(sa[i]);
      e.seqBase = uint8_t(sb[i]);
      e.refQual = QA[pa];
      e.seqQual = QB[pb];
      emit(e);
    }
    ++pa;
    ++pb;
  }
  return int(n);
}

int sdo_align_profile_batch(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* bases,
                            const uint8_t* quals, const uint64_t* offsets, const uint32_t* refIdx,
                            const uint32_t* readIdx, uint32_t nPairs, uint32_t flags, sdgpu_comparison* out,
                            sdgpu_gap* gaps, uint32_t gapCap) {
  if (!p || !bases || !offsets || !refIdx || !readIdx || !out) return SDGPU_E_ARG;
  std::vector<Cell> mat;
  AlnResult r;
  Gapped ga, gb;
  for (uint32_t i = 0; i < nPairs; ++i) {
    const uint64_t oa = offsets[refIdx[i]], ob = offsets[readIdx[i]];
    const uint32_t la = uint32_t(offsets[refIdx[i] + 1] - oa), lb = uint32_t(offsets[readIdx[i] + 1] - ob);
    if (la == 0 || lb == 0) return SDGPU_E_ARG;
    alignProfile(*p, kmaps ? kmaps->km : nullptr, bases + oa, quals ? quals + oa : nullptr, la, bases + ob,
                 quals ? quals + ob : nullptr, lb, flags, mat, r, ga, gb, out[i]);
    if (gaps) {
      copyGaps(r, gaps + size_t(i) * gapCap, gapCap);
      if (r.gaps.size() > gapCap) out[i].status |= SDGPU_ST_GAPLIST_TRUNCATED;
    }
  }
  return SDGPU_OK;
}

sdo_kmaps* sdo_index_kmers(const uint8_t* bases, const uint64_t* offsets, const double* cnts,
                           const uint32_t* seqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                           int kmersByPosition, int expandKmerPos, uint32_t expandKmerSize) {
  sdo_kmaps* h = new sdo_kmaps();
  h->km = indexKmers(bases, offsets, cnts, seqIdx, nIdx, kLen, runCutOff, kmersByPosition != 0,
                     expandKmerPos != 0, expandKmerSize);
  return h;
}

void sdo_free_kmaps(sdo_kmaps* k) {
  if (!k) return;
  delete k->km;
  delete k;
}

uint32_t sdo_kmaps_count(const sdo_kmaps* k, const uint8_t* kmer, uint32_t pos) {
  return k->km->count(std::string(reinterpret_cast<const char*>(kmer), k->km->kLen), pos);
}

int sdo_kmer_compare(const uint8_t* a, uint32_t lenA, const uint8_t* b, uint32_t lenB, uint32_t kLen,
                     uint32_t* sharedOut, double* fracOut) {
  uint32_t s;
  double f;
  compareKmers(a, lenA, b, lenB, kLen, s, f);
  if (sharedOut) *sharedOut = s;
  if (fracOut) *fracOut = f;
  return SDGPU_OK;
}

// kmerInfo(seq, kLen, true) + refInfo.compareKmersRevComp(info) (extractorPairedEnd.cpp:819-824): the k-mers of `a`
// against the k-mers of the reverse complement of `b` (njhseq seqUtil::reverseComplement, IUPAC complements).
static char complementBase(char c) {
  switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W'; case 'K': return 'M';
    case 'M': return 'K'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; case 'n': return 'n';
    default: return c;
  }
}
int sdo_kmer_compare_revcomp(const uint8_t* a, uint32_t lenA, const uint8_t* b, uint32_t lenB, uint32_t kLen,
                             uint32_t* sharedOut, double* fracOut) {
  std::vector<uint8_t> rc(lenB);
  for (uint32_t i = 0; i < lenB; ++i) rc[i] = uint8_t(complementBase(char(b[lenB - 1 - i])));
  return sdo_kmer_compare(a, lenA, rc.data(), lenB, kLen, sharedOut, fracOut);
}

// ---------------------------------------------------------------- paired-end stitching
// aligner::alignRegGlobalNoInternalGaps (njhseq; UNVERIFIED restatement): the best alignment of the two mates
// that has no internal gap, i.e. the best shift.  Score of a shift = substitution scores of the overlapping
// columns minus the end-gap penalties of the overhangs; shifts are tried from r2 furthest left to r2 furthest
// right and a later shift must score strictly more to replace the best one.
static void alignNoInternalGaps(const sdgpu_params& P, const uint8_t* A, uint32_t la, const uint8_t* B, uint32_t lb, AlnResult& out) {
  const sdgpu_gap_pars& g = P.gaps;
  long best = LONG_MIN;
  int bestShift = 0;
  for (int d = -int(lb) + 1; d <= int(la) - 1; ++d) {
    long s = 0;
    for (int i = std::max(0, d); i < std::min<int>(la, int(lb) + d); ++i) s += P.scoreMat[size_t(A[i]) * SDGPU_MAT_DIM + B[i - d]];
    if (d > 0) s -= g.gapLeftQueryOpen + (d - 1) * g.gapLeftQueryExtend;
    if (d < 0) s -= g.gapLeftRefOpen + (-d - 1) * g.gapLeftRefExtend;
    const int tail = int(la) - (int(lb) + d);
    if (tail > 0) s -= g.gapRightQueryOpen + (tail - 1) * g.gapRightQueryExtend;
    if (tail < 0) s -= g.gapRightRefOpen + (-tail - 1) * g.gapRightRefExtend;
    if (s > best) best = s, bestShift = d;
  }
  out.score = int32_t(best);
  out.gaps.clear();  // gapInfo(pos, size, gapInA) in descending position order, like the traceback's
  const int d = bestShift, tail = int(la) - (int(lb) + d);
  if (tail > 0) out.gaps.push_back(GapRun{lb, uint32_t(tail), false});
  if (tail < 0) out.gaps.push_back(GapRun{la, uint32_t(-tail), true});
  if (d > 0) out.gaps.push_back(GapRun{0u, uint32_t(d), false});
  if (d < 0) out.gaps.push_back(GapRun{0u, uint32_t(-d), true});
}

struct sdo_stitch_result {
  std::vector<uint32_t> status;
  std::vector<sdgpu_overlap> overlaps;
  std::vector<std::string> seqs;
  std::vector<std::vector<uint8_t>> quals;
  uint64_t retried = 0;
};

// PairedReadProcessor::processPairedEnd (PairedReadProcessor.cpp:287-664), one pair after the other, on strings
// like the reference: alignObjectA_ / alignObjectB_ are built with '-' and the overhang cases read their ends.
sdo_stitch_result* sdo_stitch_pairs(const sdgpu_params* Pp, const sdst_params* Sp, const uint8_t* r1Bases, const uint8_t* r1Quals,
                                    const uint64_t* r1Offsets, const uint8_t* r2Bases, const uint8_t* r2Quals,
                                    const uint64_t* r2Offsets, uint32_t n) {
  const sdgpu_params& P = *Pp;
  const sdst_params& S = *Sp;
  auto* res = new sdo_stitch_result();
  res->status.assign(n, SDST_NOOVERLAP);
  res->overlaps.assign(n, sdgpu_overlap{});
  res->seqs.resize(n);
  res->quals.resize(n);
  const double percentId = 1 - S.errorAllowed;
  struct Seq {
    std::string seq;
    std::vector<uint8_t> qual;
  };
  auto align = [&](const Seq& a, const Seq& b, Gapped& ga, Gapped& gb, sdgpu_comparison& c) {
    AlnResult aln;
    const uint8_t* A = reinterpret_cast<const uint8_t*>(a.seq.data());
    const uint8_t* B = reinterpret_cast<const uint8_t*>(b.seq.data());
    alignNoInternalGaps(P, A, uint32_t(a.seq.size()), B, uint32_t(b.seq.size()), aln);
    rearrangeGlobal(A, a.qual.data(), uint32_t(a.seq.size()), B, b.qual.data(), uint32_t(b.seq.size()), aln, ga, gb);
    std::memset(&c, 0, sizeof(c));
    c.alnScore = aln.score;
    profileAlignment(P, nullptr, A, a.qual.data(), uint32_t(a.seq.size()), B, b.qual.data(), uint32_t(b.seq.size()), ga, gb,
                     SDGPU_USING_QUALITY, c);
  };
  auto accept = [&](const sdgpu_comparison& c) {
    return c.eventBasedIdentityHq >= percentId && c.basesInAln >= S.minOverlap && c.hqMismatches + c.lqMismatches <= S.hardMismatchCutOff &&
           c.hqMismatches <= S.hqMismatchCutOff && c.lqMismatches <= S.lqMismatchCutOff;
  };
  for (uint32_t i = 0; i < n; ++i) {
    Seq r1{std::string(reinterpret_cast<const char*>(r1Bases + r1Offsets[i]), r1Offsets[i + 1] - r1Offsets[i]),
           std::vector<uint8_t>(r1Quals + r1Offsets[i], r1Quals + r1Offsets[i + 1])};
    Seq r2{std::string(reinterpret_cast<const char*>(r2Bases + r2Offsets[i]), r2Offsets[i + 1] - r2Offsets[i]),
           std::vector<uint8_t>(r2Quals + r2Offsets[i], r2Quals + r2Offsets[i + 1])};
    if (S.r1Trim) {  // trimOffEndBases
      const size_t keep = r1.seq.size() > S.r1Trim ? r1.seq.size() - S.r1Trim : 0;
      r1.seq.resize(keep);
      r1.qual.resize(keep);
    }
    if (S.r2Trim) {  // the mate is reverse-complemented: trimOffForwardBases
      const size_t cut = std::min<size_t>(r2.seq.size(), S.r2Trim);
      r2.seq.erase(0, cut);
      r2.qual.erase(r2.qual.begin(), r2.qual.begin() + cut);
    }
    Gapped ga, gb;
    sdgpu_comparison c;
    align(r1, r2, ga, gb, c);
    if (S.trimLowQualWindows && !accept(c)) {
      uint32_t firstMateWindow = UINT32_MAX;
      if (r1.seq.size() > S.windowSize)
        for (uint32_t pos = 0; pos < r1.seq.size() - S.windowSize; pos += S.windowStep) {
          double sum = 0;
          for (uint32_t q = pos; q < pos + S.windowSize; ++q) sum += r1.qual[q];
          if (sum / S.windowSize < S.avgQualCutOff) {
            firstMateWindow = pos;
            break;
          }
        }
      uint32_t secondMateWindow = 0;
      const uint32_t end = uint32_t(r2.seq.size());
      if (end > S.windowSize + 1)
        for (uint32_t pos = 0; pos < end - S.windowSize - 1; pos += S.windowStep) {
          const uint32_t truePos = end - S.windowSize - pos - 1;
          double sum = 0;
          for (uint32_t q = truePos; q < truePos + S.windowSize; ++q) sum += r2.qual[q];
          if (sum / S.windowSize < S.avgQualCutOff) {
            secondMateWindow = truePos;
            break;
          }
        }
      const bool cut1 = firstMateWindow != UINT32_MAX && firstMateWindow > 1;
      const bool cut2 = secondMateWindow != 0 && secondMateWindow + S.windowSize + 2 < r2.seq.size();
      if (cut1 || cut2) {
        Seq c1 = r1, c2 = r2;
        if (cut1) {
          c1.seq = r1.seq.substr(0, firstMateWindow);
          c1.qual.assign(r1.qual.begin(), r1.qual.begin() + firstMateWindow);
        }
        if (cut2) {
          c2.seq = r2.seq.substr(secondMateWindow + S.windowSize);
          c2.qual.assign(r2.qual.begin() + secondMateWindow + S.windowSize, r2.qual.end());
        }
        if (double(c1.seq.size()) / r1.seq.size() >= S.percentAfterTrimCutOff && double(c2.seq.size()) / r2.seq.size() >= S.percentAfterTrimCutOff) {
          align(c1, c2, ga, gb, c);
          r1 = c1;
          r2 = c2;
          ++res->retried;
        }
      }
    }
    sdgpu_overlap& o = res->overlaps[i];
    o.alnScore = c.alnScore;
    o.shift = int32_t(gb.seq.find_first_not_of('-')) - int32_t(ga.seq.find_first_not_of('-'));
    o.basesInAln = c.basesInAln;
    o.hqMismatches = c.hqMismatches;
    o.lqMismatches = c.lqMismatches;
    o.lowKmerMismatches = c.lowKmerMismatches;
    o.highQualityMatches = c.highQualityMatches;
    o.lowQualityMatches = c.lowQualityMatches;
    o.eventBasedIdentity = c.eventBasedIdentity;
    o.eventBasedIdentityHq = c.eventBasedIdentityHq;
    if (!accept(c)) continue;
    const std::string &sa = ga.seq, &sb = gb.seq;
    enum { NOOVERHANG, R1OVERHANG, R2OVERHANG };
    const int frontCase = (sa.front() != '-' && sb.front() != '-') ? NOOVERHANG : (sa.front() != '-' ? R1OVERHANG : R2OVERHANG);
    const int backCase = (sa.back() != '-' && sb.back() != '-') ? NOOVERHANG : (sa.back() != '-' ? R1OVERHANG : R2OVERHANG);
    std::string cseq;
    std::vector<uint8_t> cq;
    auto addToConsensus = [&](size_t pos) {  // PairedReadProcessor.cpp:43-78
      const char b1 = sa[pos], b2 = sb[pos];
      const uint8_t q1 = ga.qual[pos], q2 = gb.qual[pos];
      if (b1 == '-') {
        if (q2 >= P.primaryQual) cseq.push_back(b2), cq.push_back(q2);
      } else if (b2 == '-') {
        if (q1 >= P.primaryQual) cseq.push_back(b1), cq.push_back(q1);
      } else if (P.scoreMat[size_t(uint8_t(b1)) * SDGPU_MAT_DIM + uint8_t(b2)] > 0) {
        cseq.push_back(b1);
        if (islower(b1) || islower(b2)) cseq.back() = char(tolower(cseq.back()));
        cq.push_back(q1 >= q2 ? q1 : q2);
      } else {
        cseq.push_back(q1 >= q2 ? b1 : b2);
        cq.push_back(q1 >= q2 ? q1 : q2);
        if (islower(b1) || islower(b2)) cseq.back() = char(tolower(cseq.back()));
      }
    };
    if (frontCase == NOOVERHANG && backCase == NOOVERHANG) {
      for (size_t pos = 0; pos < sa.size(); ++pos) addToConsensus(pos);
      res->status[i] = SDST_PERFECTOVERLAP;
    } else if ((frontCase == NOOVERHANG || frontCase == R1OVERHANG) && (backCase == NOOVERHANG || backCase == R2OVERHANG)) {
      const size_t r1End = sa.find_last_not_of('-') + 1, r2Start = sb.find_first_not_of('-');
      if (r2Start != 0) {
        cseq.append(sa.substr(0, r2Start));
        cq.insert(cq.end(), ga.qual.begin(), ga.qual.begin() + r2Start);
      }
      for (size_t pos = r2Start; pos < r1End; ++pos) addToConsensus(pos);
      if (sa.size() != r1End) {
        cseq.append(sb.substr(r1End));
        cq.insert(cq.end(), gb.qual.begin() + r1End, gb.qual.end());
      }
      res->status[i] = SDST_R1ENDSINR2;
    } else {
      size_t r1Start = 0, r2End = sb.size();
      if (frontCase != NOOVERHANG) r1Start = sa.find_first_not_of('-');
      if (backCase != NOOVERHANG) r2End = sb.find_last_not_of('-') + 1;
      for (size_t pos = r1Start; pos < r2End; ++pos) addToConsensus(pos);
      if ((frontCase == NOOVERHANG || frontCase == R2OVERHANG) && (backCase == NOOVERHANG || backCase == R1OVERHANG)) {
        res->status[i] = SDST_R1BEGINSINR2;
      } else if (frontCase == R2OVERHANG && backCase == R2OVERHANG) {
        res->status[i] = SDST_R1ALLINR2;
      } else {
        res->status[i] = SDST_R2ALLINR1;
      }
    }
    res->seqs[i] = cseq;
    res->quals[i] = cq;
  }
  return res;
}
void sdo_stitch_free(sdo_stitch_result* r) { delete r; }
uint64_t sdo_stitch_retried(const sdo_stitch_result* r) { return r->retried; }
uint32_t sdo_stitch_status(const sdo_stitch_result* r, uint32_t i) { return r->status[i]; }
void sdo_stitch_overlap(const sdo_stitch_result* r, uint32_t i, sdgpu_overlap* out) { *out = r->overlaps[i]; }
uint32_t sdo_stitch_len(const sdo_stitch_result* r, uint32_t i) { return uint32_t(r->seqs[i].size()); }
void sdo_stitch_seq(const sdo_stitch_result* r, uint32_t i, uint8_t* seq, uint8_t* qual) {
  std::memcpy(seq, r->seqs[i].data(), r->seqs[i].size());
  std::memcpy(qual, r->quals[i].data(), r->quals[i].size());
}

// qluster body, following clusterDown.cpp:230-563 (dedupe -> quality rep -> sort -> indexKmers
// -> singlet split -> runFullClustering x3 -> final sort).
static sdo_qluster_result* qlusterImpl(const sdgpu_params* p, const sdo_qluster_opts* o, const sdo_iter_par* initPars,
                                       uint32_t nInit, const sdo_iter_par* pars, uint32_t nPars, const uint8_t* bases,
                                       const uint8_t* quals, const uint64_t* offsets, uint32_t n, const double* clusterCnts,
                                       bool clustersIn) {
  Engine E(*p, *o);
  auto* res = new sdo_qluster_result();
  res->membership.assign(n, UINT32_MAX);
  // exact-sequence dedupe, first-seen order (clusterDown.cpp:272-288); population input (processClusters.cpp:226)
  // arrives as clusters with their counts and is not deduplicated
  std::unordered_map<std::string, uint32_t> seen;
  for (uint32_t r = 0; r < n; ++r) {
    const uint32_t len = uint32_t(offsets[r + 1] - offsets[r]);
    if (len <= o->smallReadSize) continue;  // clusterDown.cpp:269-270
    std::string s(reinterpret_cast<const char*>(bases + offsets[r]), len);
    if (clustersIn) {
      E.uniq.emplace_back();
      E.uniq.back().seq = s;
      E.uniq.back().cnt = clusterCnts ? clusterCnts[r] : 1.0;
      E.uniq.back().inputIdx.push_back(r);
      continue;
    }
    auto it = seen.find(s);
    if (it == seen.end()) {
      it = seen.emplace(s, uint32_t(E.uniq.size())).first;
      E.uniq.emplace_back();
      E.uniq.back().seq = s;
    }
    Uniq& u = E.uniq[it->second];
    u.cnt += 1.0;
    u.inputIdx.push_back(r);
  }
  // per-base median quality over members (clusterDown.cpp:293-359, qualRep = "median")
  for (Uniq& u : E.uniq) {
    const size_t len = u.seq.size();
    u.qual.resize(len);
    std::vector<uint8_t> col(u.inputIdx.size());
    for (size_t i = 0; i < len; ++i) {
      for (size_t m = 0; m < u.inputIdx.size(); ++m) col[m] = quals[offsets[u.inputIdx[m]] + i];
      std::sort(col.begin(), col.end());
      const size_t k = col.size();
      u.qual[i] = (k % 2) ? col[k / 2] : uint8_t((uint32_t(col[k / 2 - 1]) + col[k / 2]) / 2);
    }
  }
  res->nUnique = uint32_t(E.uniq.size());
  std::vector<Clus> clusters(E.uniq.size());
  for (uint32_t u = 0; u < E.uniq.size(); ++u) {
    Clus& c = clusters[u];
    c.seq = E.uniq[u].seq;
    c.qual = E.uniq[u].qual;
    c.cnt = E.uniq[u].cnt;
    c.firstId = u;
    c.members = {u};
    c.avgErr = averageErrorRate(c.qual);
  }
  std::sort(clusters.begin(), clusters.end(), clusLess);  // clusterDown.cpp:389

  auto buildIndex = [&](const std::vector<Clus>& cs) -> KMaps* {
    // pack the current cluster sequences and index them (clusterDown.cpp:412-417 / 505-510)
    std::vector<uint8_t> b;
    std::vector<uint64_t> off{0};
    std::vector<double> cn;
    for (const Clus& c : cs) {
      b.insert(b.end(), c.seq.begin(), c.seq.end());
      off.push_back(b.size());
      cn.push_back(c.cnt);
    }
    return indexKmers(b.data(), off.data(), cn.data(), nullptr, uint32_t(cs.size()), o->kLen, o->runCutOff,
                      o->kmersByPosition != 0, o->expandKmerPos != 0, o->expandKmerSize);
  };
  KMaps* km = buildIndex(clusters);
  E.km = km;
  std::vector<Clus> singletons;
  if (!o->startWithSingles) {  // clusterDown.cpp:467-484
    std::vector<Clus> keep;
    for (Clus& c : clusters) (c.cnt <= 1.01 ? singletons : keep).push_back(std::move(c));
    clusters.swap(keep);
  }
  E.runFullClustering(clusters, initPars, nInit);  // clusterDown.cpp:488-490
  if (!o->startWithSingles && !o->leaveOutSinglets) {  // clusterDown.cpp:492-499
    for (Clus& c : singletons) clusters.push_back(std::move(c));
    E.runFullClustering(clusters, pars, nPars);
  }
  if (!o->dontRecalLowFreq) {  // clusterDown.cpp:502-519
    KMaps* km2 = buildIndex(clusters);
    E.km = km2;
    delete km;
    km = km2;
    for (Clus& c : clusters) c.prev.clear();
    E.runFullClustering(clusters, pars, nPars);
  }
  delete km;
  std::sort(clusters.begin(), clusters.end(), clusLess);  // clusterDown.cpp:560
  for (uint32_t c = 0; c < clusters.size(); ++c)
    for (uint32_t u : clusters[c].members)
      for (uint32_t r : E.uniq[u].inputIdx) res->membership[r] = c;
  for (Clus& c : clusters) c.prev.clear();
  if (o->markChimeras) E.markChimeras(clusters);  // clusterDown.cpp:566-577
  res->clusters = std::move(clusters);
  res->nAlign = E.nAlign;
  res->nCells = E.nCells;
  return res;
}

sdo_qluster_result* sdo_qluster(const sdgpu_params* p, const sdo_qluster_opts* o, const sdo_iter_par* initPars,
                                uint32_t nInit, const sdo_iter_par* pars, uint32_t nPars, const uint8_t* bases,
                                const uint8_t* quals, const uint64_t* offsets, uint32_t n) {
  return qlusterImpl(p, o, initPars, nInit, pars, nPars, bases, quals, offsets, n, nullptr, false);
}

// doPopulationClustering's collapse (processClusters.cpp:226; checkKmers off, :127): one runClustering over the
// samples' clusters with their counts
sdo_qluster_result* sdo_population_cluster(const sdgpu_params* p, const sdo_qluster_opts* o, const sdo_iter_par* popPars,
                                           uint32_t nPars, const uint8_t* bases, const uint8_t* quals,
                                           const uint64_t* offsets, const double* cnts, uint32_t n) {
  sdo_qluster_opts q = *o;
  q.checkKmers = 0;
  q.startWithSingles = 1;
  q.dontRecalLowFreq = 1;
  q.markChimeras = 0;
  return qlusterImpl(p, &q, popPars, nPars, popPars, nPars, bases, quals, offsets, n, cnts, true);
}

void sdo_qluster_free(sdo_qluster_result* r) { delete r; }
uint32_t sdo_qluster_num_clusters(const sdo_qluster_result* r) { return uint32_t(r->clusters.size()); }
uint32_t sdo_qluster_num_unique(const sdo_qluster_result* r) { return r->nUnique; }
uint64_t sdo_qluster_num_alignments(const sdo_qluster_result* r) { return r->nAlign; }
uint64_t sdo_qluster_num_cells(const sdo_qluster_result* r) { return r->nCells; }
uint32_t sdo_qluster_cluster_len(const sdo_qluster_result* r, uint32_t c) { return uint32_t(r->clusters[c].seq.size()); }
double sdo_qluster_cluster_cnt(const sdo_qluster_result* r, uint32_t c) { return r->clusters[c].cnt; }
void sdo_qluster_cluster_seq(const sdo_qluster_result* r, uint32_t c, uint8_t* seq, uint8_t* qual) {
  const Clus& cl = r->clusters[c];
  if (seq) std::memcpy(seq, cl.seq.data(), cl.seq.size());
  if (qual) std::memcpy(qual, cl.qual.data(), cl.qual.size());
}
int sdo_qluster_cluster_chimeric(const sdo_qluster_result* r, uint32_t c, uint32_t* parents, uint32_t* positions) {
  const Clus& cl = r->clusters[c];
  if (cl.chimeric && parents) parents[0] = cl.chiParent[0], parents[1] = cl.chiParent[1];
  if (cl.chimeric && positions) positions[0] = cl.chiPos[0], positions[1] = cl.chiPos[1];
  return cl.chimeric ? 1 : 0;
}
int sdo_chimera_probe_batch(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* bases, const uint8_t* quals,
                            const uint64_t* offsets, const uint32_t* parentIdx, const uint32_t* candIdx,
                            uint32_t nPairs, uint32_t flags, const sdo_iter_par* chiOverlap,
                            int keepLowQualityMismatches, sdgpu_chimera_probe* out) {
  if (!p || !bases || !offsets || !parentIdx || !candIdx || !out || !chiOverlap) return SDGPU_E_ARG;
  std::vector<Cell> mat;
  for (uint32_t i = 0; i < nPairs; ++i) {
    const uint64_t oa = offsets[parentIdx[i]], ob = offsets[candIdx[i]];
    const uint32_t la = uint32_t(offsets[parentIdx[i] + 1] - oa), lb = uint32_t(offsets[candIdx[i] + 1] - ob);
    if (la == 0 || lb == 0) return SDGPU_E_ARG;
    chimeraProbe(*p, kmaps ? kmaps->km : nullptr, bases + oa, quals ? quals + oa : nullptr, la, bases + ob,
                 quals ? quals + ob : nullptr, lb, flags, *chiOverlap, keepLowQualityMismatches != 0, mat, out[i]);
  }
  return SDGPU_OK;
}
void sdo_qluster_membership(const sdo_qluster_result* r, uint32_t* clusterOfRead) {
  std::memcpy(clusterOfRead, r->membership.data(), r->membership.size() * sizeof(uint32_t));
}

}  // extern "C"
