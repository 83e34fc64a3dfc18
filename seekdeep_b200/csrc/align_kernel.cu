// align_kernel.cu — kernels 2+3 of the qluster hot path, fused:
//   forward affine-gap global DP (njhseq alignCalc::runNeedleSave semantics) with packed
//   traceback bits -> warp-cooperative traceback -> gapped strings -> errorProfile
//   (aligner::profileAlignment semantics) -> one sdgpu_comparison and/or par-line verdict mask.
//
// Replaces, per (ref, read) pair: aligner::alignCacheGlobal + aligner::profileAlignment (+
// comparison::passErrorProfile) — reference call sites clusterDown.cpp:694-696, 739-741;
// ControlBencher.hpp:117-120.
//
// Formulation (deliberately different from the reference's cell-of-three-scores matrix):
//   the read (B, DP columns) is cut into 32 strips of C columns, each lane keeps its strip's
//   state in registers; the ref (A, DP rows) streams through the lanes as a skewed wavefront
//   (lane l works on row t-l at step t).  Each cell produces its three OUTGOING maxima
//       toDiag  = max(U, L, D)                -> D of the cell down-right  (+ score)
//       toDown  = max(U-ge, L-go, D-go)       -> U of the cell below
//       toRight = max(U-go, L-ge, D-go)       -> L of the cell to the right
//   with njhseq's needleMaximum tie order (U, then L, then D; its 'B' pointer behaves as 'U'
//   everywhere in the traceback), recorded as the sign bits of five differences per cell.
//   Neighbouring lanes exchange toRight/toDiag of their last column with two shuffles per row.
//   Bits go to a per-warp slab in HBM/L2 in (step, word, lane) order so every store is a full
//   128-byte line.  End-gap penalties: last row via per-row registers, last column via
//   per-column registers, first row/column through the initial register state and lane 0's
//   boundary feed.
//
// Three forward passes share the traceback / errorProfile code:
//   alignBandKernel          first pass of every launch of >= 64 pairs: 64 diagonals, anti-diagonal
//                            by anti-diagonal, one cell per lane per step; a pair is finished here
//                            only if its banded score proves the full-matrix optimum and traceback
//                            lie inside the band (see the comment above the kernel), otherwise it is
//                            queued for one of the two full-matrix passes below.
//   alignProfileKernel<C>    one pair per warp, int32 lanes: maxima (two of them fused add+max)
//                            and funnel shifts on the ALU pipe, differences on the FMA pipe.
//   alignProfileKernelH2<C>  TWO pairs per warp in the halves of __half2 registers.  Every DP
//                            value of a launch is an integer of magnitude <= 2047 (checked on
//                            the host), which fp16 represents exactly, so HADD2/HMNMX2 compute
//                            the same numbers as the int32 path with half the instructions per
//                            cell; the sign bits of both halves are collected with one LOP3.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>

#include "sdgpu_internal.cuh"

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr unsigned FULL = 0xffffffffu;
constexpr int H2_SYMS = 5;  // the packed kernel scores through a joint (A,C,G,T,N)^2 x (A,C,G,T,N)^2 table

struct AlignArgs {
  DevScoring sc;
  DevSeqs seqs;
  DevKmerIndex idx;
  DevPairSrc src;
  uint32_t nPairs;
  uint32_t flags;
  sdgpu_comparison* out;
  sdgpu_gap* gapsOut;
  uint32_t gapCap;
  uint8_t* scratch;
  uint64_t slabBytes;  // per warp
  uint32_t ptrBytes;   // bytes of the pointer region in a slab
  uint32_t maxAln;     // capacity of each gapped string
  uint32_t edgeOff;    // WIDE: byte offset of the block-edge buffers in a slab
  uint32_t edgeLen;    // WIDE: entries per edge buffer
  uint64_t blockWords; // WIDE: words of one column block's bit region
  uint32_t* work;      // [0] = next work item, [2..3] = cell counter (uint64)
  const sdgpu_iter_par* lines;  // par-file lines for the verdict mask (device)
  uint32_t nLines;
  uint64_t* passOut;   // per pair: bit l = passes line l (or NULL)
  DevHits hits;        // grid launches: (read, cluster, mask) records of the non-zero masks (recs may be NULL)
  const uint32_t* pairMap;   // redo pass: work item -> pair index (or NULL: identity)
  const uint32_t* nWorkDev;  // redo pass: number of work items, read on the device (or NULL: nPairs)
  uint32_t* redo;            // banded pass: pairs whose band could not be certified
  int maxSub;                // largest substitution score (certificate bound)
  int minRight, minDown;     // cheapest gap column in the ref / in the read away from the right ends
  sdgpu_chimera_probe* chiOut;  // per pair: markChimeras' front/back probe (or NULL)
  sdgpu_iter_par chiPar;        // chiOpts_.chiOverlap_
  uint32_t chiKeepLq;           // chiOpts_.keepLowQaulityMismatches_
};

enum { ST_D = 0, ST_U = 1, ST_L = 2 };

// Canonical decision bits of a cell (set == ">=", njhseq's tie order U, L, D):
//   b0: L >= D | b1: U >= max(L,D) | b2: U-ge >= max(L,D)-go | b3: L-ge >= D-go | b4: U-go >= max(L-ge,D-go)
__device__ __forceinline__ int nextFromDiag(unsigned b) { return (b & 2u) ? ST_U : ((b & 1u) ? ST_L : ST_D); }
__device__ __forceinline__ int nextFromUp(unsigned b) { return (b & 4u) ? ST_U : ((b & 1u) ? ST_L : ST_D); }
__device__ __forceinline__ int nextFromLeft(unsigned b) { return (b & 16u) ? ST_U : ((b & 8u) ? ST_L : ST_D); }

// int32 forward pass: five sign bits per cell pushed MSB-first with funnel shifts, 6 cells per word
template <int C>
struct BitsI32 {
  static constexpr int CPW = 6;
  static constexpr int CW = (C + CPW - 1) / CPW;
  const unsigned* words;
  size_t blockWords;  // words of one column block's bit region (reads wider than 32*C columns are strip-mined)
  __device__ __forceinline__ unsigned operator()(int i, int j) const {
    const int gj = (j - 1) / C, c = (j - 1) - gj * C;
    const int blk = gj >> 5, lj = gj & 31;
    const int t = i + lj;
    const int wi = c / CPW, k = c - wi * CPW;
    const int inWord = (C - wi * CPW) < CPW ? (C - wi * CPW) : CPW;
    const unsigned w = __ldcg(words + size_t(blk) * blockWords + (size_t(t - 1) * CW + wi) * 32 + lj);
    const unsigned g = ~(w >> (5 * (inWord - 1 - k))) & 31u;  // g4 = first pushed
    return __brev(g) >> 27;
  }
};

// half2 forward pass: sign bits of both halves inserted at bits 15 / 31 and shifted right,
// 3 cells (15 bits) per 16-bit half; `half` selects the pair
template <int C>
struct BitsH2 {
  static constexpr int CPW = 3;
  static constexpr int CW = (C + CPW - 1) / CPW;
  const unsigned* words;
  int half;
  __device__ __forceinline__ unsigned operator()(int i, int j) const {
    const int lj = (j - 1) / C, c = (j - 1) - lj * C;
    const int t = i + lj;
    const int wi = c / CPW, k = c - wi * CPW;
    const int inWord = (C - wi * CPW) < CPW ? (C - wi * CPW) : CPW;
    const unsigned w = __ldcg(words + (size_t(t - 1) * CW + wi) * 32 + lj);
    return ~(w >> (16 * half + 16 - 5 * inWord + 5 * k)) & 31u;  // bit 0 = first inserted
  }
};

// banded forward pass: lane = (offset - dlo) / 2, one cell per lane per anti-diagonal, six
// anti-diagonals (30 bits) per word, first cell in the highest bits
struct BitsBand {
  const unsigned* words;
  int dlo;
  int aBase;  // first anti-diagonal of the pass (1 or 2: the one on which lanes meet their even offset)
  __device__ __forceinline__ unsigned operator()(int i, int j) const {
    const int st = i + j - aBase;
    int ln = (j - i - dlo) >> 1;
    ln = ln < 0 ? 0 : (ln > 31 ? 31 : ln);  // speculative traceback lanes may look outside the band; their result is discarded
    const int g = st / 6, k = st - g * 6;
    const unsigned w = __ldcg(words + size_t(g) * 32 + ln);
    const unsigned gb = ~(w >> (5 * (5 - k))) & 31u;
    return __brev(gb) >> 27;
  }
};

// Quality windows (seqInfo::checkQual) and low-frequency k-mer look-ups (KmerMaps::isKmerLowFrequency) are not
// evaluated per mismatch here: both are properties of one base of one stored sequence, kept as per-base flag
// bytes next to the codes (SD_FLAG_QUAL_OK written at upload, SD_FLAG_LOW_KMER at every index build).
__device__ __forceinline__ unsigned warpSum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

struct PairView {
  uint32_t pair;
  int la, lb;
  const uint8_t *cA, *cB, *qA, *qB;
  const uint8_t *fA, *fB;  // per-base SD_FLAG_* bytes
};

// njhseq aligner::handleGapCountingInA/InB weight of one gap run: 1, or for a homopolymer run
// under weighHomopolymers |gapSide - baseSide| / (gapSide + baseSide) over the flanking run
__device__ double gapWeightDev(const uint8_t* gA, const uint8_t* gB, int alnLen, int startPos, int size, bool inA) {
  const uint8_t* gs = inA ? gA : gB;  // string holding the '-' run
  const uint8_t* os = inA ? gB : gA;  // string holding the gapped bases
  const int base = os[startPos];
  for (int x = 1; x < size; ++x)
    if (os[startPos + x] != base) return 1.0;
  double gapSide = 0, baseSide = size;
  for (int cur = startPos + size; cur < alnLen && gs[cur] == base; ++cur) gapSide += 1;
  for (int cur = 1; cur <= startPos && gs[startPos - cur] == base; ++cur) gapSide += 1;
  for (int cur = startPos + size; cur < alnLen && os[cur] == base; ++cur) baseSide += 1;
  for (int cur = 1; cur <= startPos && os[startPos - cur] == base; ++cur) baseSide += 1;
  return fabs(gapSide - baseSide) / (gapSide + baseSide);
}

// collapser::markChimeras' per-(parent, candidate) probe on the traced alignment: first / last
// kept mismatch and the errorProfiles of the columns before the first and after the last one,
// judged against chiOverlap.  The windows hold no kept mismatch by construction, so their
// profiles are the low-quality mismatches and the counted gap runs that fall inside them.
// (compiled only into the CHI instantiations of the int32 kernels, so the probe's doubles do not
// cost the collapse loop's kernels registers)
struct ChiArgs {
  int primaryQual, secondaryQual, window, countEndGaps, weighHomopolymers;
  uint32_t flags, chiKeepLq;
  DevKmerIndex idx;
  sdgpu_chimera_probe* chiOut;
  double one, two, large, hq, lq, lk;  // chiOverlap thresholds
};
__device__ __forceinline__ void chimeraProbeDev(const ChiArgs& a, const PairView& pv, const uint8_t* gA, const uint8_t* gB,
                                             int alnLen, int alnStart, const uint32_t* gapRecs, int nRecs,
                                             const int* sTab, int score, int lane) {
  const ChiArgs& sc = a;
  const bool checkKmer = (a.flags & SDGPU_CHECK_KMER) != 0;
  const bool usingQuality = (a.flags & SDGPU_USING_QUALITY) != 0;
  const unsigned ltMask = (1u << lane) - 1u;
  int baseA = 0, baseB = 0;
  int firstCol = -1, lastCol = -1, firstPos = 0, lastPos = 0;
  unsigned kept = 0, lqFront = 0, lqBack = 0;
  for (int cs = 0; cs < alnLen; cs += 32) {
    const int k = cs + lane;
    const bool inb = k < alnLen;
    const int ca = inb ? int(gA[k]) : SDGPU_GAP_CODE;
    const int cb = inb ? int(gB[k]) : SDGPU_GAP_CODE;
    const unsigned hasA = __ballot_sync(FULL, inb && ca != SDGPU_GAP_CODE);
    const unsigned hasB = __ballot_sync(FULL, inb && cb != SDGPU_GAP_CODE);
    int cls = 0;  // 1 = hq, 2 = lq, 3 = lowKmer
    const int pa = baseA + __popc(hasA & ltMask), pb = baseB + __popc(hasB & ltMask);
    if (inb && ca != SDGPU_GAP_CODE && cb != SDGPU_GAP_CODE && sTab[ca * SDGPU_MAX_SYMBOLS + cb] < 0) {
      const unsigned fa = pv.fA[pa], fb = pv.fB[pb];
      const bool good = !usingQuality || (fa & fb & SD_FLAG_QUAL_OK);
      if (!good) {
        cls = 2;
      } else if (checkKmer && (SDGPU_LOWKMER_FLAGS(fa, fb) & SD_FLAG_LOW_KMER)) {
        cls = 3;
      } else {
        cls = 1;
      }
    }
    const unsigned lqMask = __ballot_sync(FULL, cls == 2);
    const unsigned keptMask = a.chiKeepLq ? __ballot_sync(FULL, cls != 0) : __ballot_sync(FULL, cls == 1 || cls == 3);
    const unsigned dropLq = a.chiKeepLq ? 0u : lqMask;  // mismatches markChimeras does not look at
    if (keptMask) {
      const int f = __ffs(keptMask) - 1, l = 31 - __clz(keptMask);
      if (firstCol < 0) {
        firstCol = cs + f;
        firstPos = __shfl_sync(FULL, pb, f);
        lqFront += __popc(dropLq & ((1u << f) - 1u));
      }
      lastCol = cs + l;
      lastPos = __shfl_sync(FULL, pb, l);
      lqBack = __popc(dropLq & ~((2u << l) - 1u));
      kept += __popc(keptMask);
    } else {
      if (firstCol < 0) lqFront += __popc(dropLq);
      lqBack += __popc(dropLq);
    }
    baseA += __popc(hasA);
    baseB += __popc(hasB);
  }
  // counted gap runs, by window
  double oneF = 0, twoF = 0, largeF = 0, oneB = 0, twoB = 0, largeB = 0;
  unsigned numGaps = 0;
  for (int r = nRecs - 1; r >= 0; --r) {
    const int startPos = int(gapRecs[3 * r + 0]) - alnStart;
    const int size = int(gapRecs[3 * r + 1]);
    const bool inA = (gapRecs[3 * r + 2] & 1u) != 0;
    const bool endGap = (startPos + size >= alnLen) || startPos == 0;
    if (endGap && !sc.countEndGaps) continue;
    ++numGaps;
    const bool front = firstCol >= 0 && startPos < firstCol, back = lastCol >= 0 && startPos > lastCol;
    if (!front && !back) continue;
    const double wgt = sc.weighHomopolymers ? gapWeightDev(gA, gB, alnLen, startPos, size, inA) : 1.0;
    if (front) (size >= 3 ? largeF : (size == 2 ? twoF : oneF)) += wgt;
    if (back) (size >= 3 ? largeB : (size == 2 ? twoB : oneB)) += wgt;
  }
  if (lane == 0) {
    sdgpu_chimera_probe o;
    o.alnScore = score;
    o.keptMismatches = kept;
    o.numGaps = numGaps;
    o.firstPos = uint32_t(firstPos);
    o.lastPos = uint32_t(lastPos);
    o.firstAlnPos = uint32_t(firstCol < 0 ? 0 : firstCol);
    o.lastAlnPos = uint32_t(lastCol < 0 ? 0 : lastCol);
    o.flags = 0;
    if (kept) {
      const bool base = a.hq >= 0.0 && a.lk >= 0.0;
      if (base && a.one >= oneF && a.two >= twoF && a.large >= largeF && a.lq >= double(lqFront)) o.flags |= SDGPU_CHI_FRONT_PASS;
      if (base && a.one >= oneB && a.two >= twoB && a.large >= largeB && a.lq >= double(lqBack)) o.flags |= SDGPU_CHI_BACK_PASS;
    }
    a.chiOut[pv.pair] = o;
  }
}

// Traceback (njhseq runNeedleSave's pointer walk + gapInfo records), gapped strings
// (rearrangeGlobal), errorProfile (profileAlignment + setEventBaseIdentity), verdict mask
// (passErrorProfile) and outputs for one pair; warp-cooperative, `bits` reads the forward
// pass's decision bits.
// SCORE_FROM_PATH: the caller did not keep the DP's corner value; the score is re-derived as the
// score of the traced alignment (substitution scores of the aligned columns minus each gap run's
// open + extends under the ten position-dependent penalties), which equals it by construction.
template <bool SCORE_FROM_PATH, bool CHI, class Bits>
__device__ void finishPair(const AlignArgs& a, const Bits& bits, const PairView& pv, int score, uint8_t* alnA,
                           uint8_t* alnB, uint32_t* gapRecs, const int* sTab, int lane) {
  const DevScoring& sc = a.sc;
  const int la = pv.la, lb = pv.lb;
  const uint8_t *cA = pv.cA, *cB = pv.cB, *fA = pv.fA, *fB = pv.fB;
  const uint32_t pair = pv.pair;
  // ------------------------------------------------------------------ traceback
  int i = la, j = lb;
  int state = nextFromDiag(bits(la, lb));  // arg-max of the bottom-right cell = first state
  int wp = la + lb - 1;                    // next free column (filled backwards)
  uint32_t gapA = 0, gapB = 0, nRecs = 0;
  while (i > 0 || j > 0) {
    const int di = (state != ST_L), dj = (state != ST_U);
    const int ik = i - lane * di, jk = j - lane * dj;
    bool valid;
    int ns = ST_D;
    if (state == ST_U) {
      valid = ik >= 1;
      if (valid) ns = (jk == 0) ? ST_U : (ik == 1 ? ST_L : nextFromUp(bits(ik - 1, jk)));
    } else if (state == ST_L) {
      valid = jk >= 1;
      if (valid) ns = (ik == 0) ? ST_L : (jk == 1 ? ST_U : nextFromLeft(bits(ik, jk - 1)));
    } else {
      valid = ik >= 1 && jk >= 1;
      if (valid) ns = (ik == 1) ? ST_L : (jk == 1 ? ST_U : nextFromDiag(bits(ik - 1, jk - 1)));
    }
    const unsigned cont = __ballot_sync(FULL, valid && ns == state);
    // lanes 0..n-1 continue in this state; lane n either takes the run's last step (and its
    // arg-max names the next state) or lies beyond (0,0)
    const int n = (cont == FULL) ? 32 : (__ffs(~cont) - 1);
    const unsigned validMask = __ballot_sync(FULL, valid);
    const bool lastTaken = (n < 32) && ((validMask >> n) & 1u);
    const int nsteps = n + (lastTaken ? 1 : 0);
    const int newState = lastTaken ? __shfl_sync(FULL, ns, n & 31) : state;
    if (lane < nsteps) {
      alnA[wp - lane] = (state == ST_L) ? uint8_t(SDGPU_GAP_CODE) : cA[ik - 1];
      alnB[wp - lane] = (state == ST_U) ? uint8_t(SDGPU_GAP_CODE) : cB[jk - 1];
    }
    i -= nsteps * di;
    j -= nsteps * dj;
    wp -= nsteps;
    if (state == ST_U) {
      gapB += nsteps;
      if (newState != ST_U) {  // gapInfo(jcursor, gapBSize, false)
        if (lane == 0) {
          gapRecs[3 * nRecs + 0] = uint32_t(wp + 1);
          gapRecs[3 * nRecs + 1] = gapB;
          gapRecs[3 * nRecs + 2] = uint32_t(j) << 1;
          if (a.gapsOut && nRecs < a.gapCap) a.gapsOut[size_t(pair) * a.gapCap + nRecs] = sdgpu_gap{uint32_t(j), gapB, 0u};
        }
        ++nRecs;
        gapB = 0;
      }
    } else if (state == ST_L) {
      gapA += nsteps;
      if (newState != ST_L) {  // gapInfo(icursor, gapASize, true)
        if (lane == 0) {
          gapRecs[3 * nRecs + 0] = uint32_t(wp + 1);
          gapRecs[3 * nRecs + 1] = gapA;
          gapRecs[3 * nRecs + 2] = (uint32_t(i) << 1) | 1u;
          if (a.gapsOut && nRecs < a.gapCap) a.gapsOut[size_t(pair) * a.gapCap + nRecs] = sdgpu_gap{uint32_t(i), gapA, 1u};
        }
        ++nRecs;
        gapA = 0;
      }
    }
    state = newState;
  }
  if (gapB != 0 || gapA != 0) {  // run still open when (0,0) is reached
    const bool inA = gapA != 0;
    if (lane == 0) {
      gapRecs[3 * nRecs + 0] = uint32_t(wp + 1);
      gapRecs[3 * nRecs + 1] = inA ? gapA : gapB;
      gapRecs[3 * nRecs + 2] = inA ? 1u : 0u;  // position 0
      if (a.gapsOut && nRecs < a.gapCap)
        a.gapsOut[size_t(pair) * a.gapCap + nRecs] = sdgpu_gap{0u, inA ? gapA : gapB, inA ? 1u : 0u};
    }
    ++nRecs;
  }
  __syncwarp();
  const int alnStart = wp + 1;
  const int alnLen = la + lb - alnStart;
  const uint8_t* gA = alnA + alnStart;
  const uint8_t* gB = alnB + alnStart;

  // ------------------------------------------------------------------ errorProfile
  const bool checkKmer = (a.flags & SDGPU_CHECK_KMER) != 0;
  const bool usingQuality = (a.flags & SDGPU_USING_QUALITY) != 0;
  const bool matchQuality = (a.flags & SDGPU_MATCH_QUALITY) != 0;
  unsigned hq = 0, lq = 0, lk = 0, hqM = 0, lqM = 0;
  int matchSum = 0;
  int baseA = 0, baseB = 0;
  const unsigned ltMask = (1u << lane) - 1u;
  // Pass 1 walks the columns 32 at a time and only sorts them into matches and mismatches; the
  // mismatches' (ref, read) base positions are queued (top of the gap-record area, growing down:
  // gap runs + mismatches <= columns, so the two never meet).  Pass 2 classifies the queued
  // mismatches 32 at a time -- quality windows and low-frequency k-mer look-ups run with every lane
  // busy instead of the one or two lanes that hold a mismatch in a given 32-column chunk.
  uint32_t* misTop = gapRecs + 3 * (size_t(a.maxAln) + 2);
  unsigned nMis = 0;
  for (int cs = 0; cs < alnLen; cs += 32) {
    const int k = cs + lane;
    const bool inb = k < alnLen;
    const int ca = inb ? int(gA[k]) : SDGPU_GAP_CODE;
    const int cb = inb ? int(gB[k]) : SDGPU_GAP_CODE;
    const unsigned hasA = __ballot_sync(FULL, inb && ca != SDGPU_GAP_CODE);
    const unsigned hasB = __ballot_sync(FULL, inb && cb != SDGPU_GAP_CODE);
    const bool both = inb && ca != SDGPU_GAP_CODE && cb != SDGPU_GAP_CODE;
    const int pa = baseA + __popc(hasA & ltMask), pb = baseB + __popc(hasB & ltMask);
    const int sub = both ? sTab[ca * SDGPU_MAX_SYMBOLS + cb] : 0;
    if (SCORE_FROM_PATH) matchSum += sub;
    const unsigned misMask = __ballot_sync(FULL, both && sub < 0);
    if (both && sub < 0) {
      uint32_t* e = misTop - 2 * (nMis + __popc(misMask & ltMask) + 1);
      e[0] = uint32_t(pa);
      e[1] = uint32_t(pb);
    }
    nMis += __popc(misMask);
    if (matchQuality) {
      if (both && sub >= 0) {
        if (!(fA[pa] & fB[pb] & SD_FLAG_QUAL_OK)) {
          ++lqM;
        } else {
          ++hqM;
        }
      }
    } else if (lane == 0) {
      hqM += __popc(hasA & hasB & ~misMask);
    }
    baseA += __popc(hasA);
    baseB += __popc(hasB);
  }
  __syncwarp();
  for (unsigned e0 = 0; e0 < nMis; e0 += 32) {
    const unsigned e = e0 + lane;
    int cls = 0;  // 1 = high quality, 2 = low quality, 3 = low k-mer
    if (e < nMis) {
      const int pa = int(misTop[-2 * int(e + 1)]), pb = int(misTop[-2 * int(e + 1) + 1]);
      const unsigned fa = fA[pa], fb = fB[pb];
      const bool good = !usingQuality || (fa & fb & SD_FLAG_QUAL_OK);
      if (!good) {
        cls = 2;
      } else if (checkKmer && (SDGPU_LOWKMER_FLAGS(fa, fb) & SD_FLAG_LOW_KMER)) {
        cls = 3;
      } else {
        cls = 1;
      }
    }
    hq += __popc(__ballot_sync(FULL, cls == 1));
    lq += __popc(__ballot_sync(FULL, cls == 2));
    lk += __popc(__ballot_sync(FULL, cls == 3));
  }
  if (matchQuality) {
    hqM = warpSum(hqM);
    lqM = warpSum(lqM);
  } else {
    hqM = __shfl_sync(FULL, hqM, 0);
  }

  // gap runs in forward order (the records are in traceback order); uniform across the warp
  double one = 0, two = 0, large = 0;
  int gapCost = 0;
  unsigned numGaps = 0, endA = 0, endB = 0, regA = 0, regB = 0;
  for (int r = int(nRecs) - 1; r >= 0; --r) {
    const int startPos = int(gapRecs[3 * r + 0]) - alnStart;
    const int size = int(gapRecs[3 * r + 1]);
    const bool inA = (gapRecs[3 * r + 2] & 1u) != 0;
    if (SCORE_FROM_PATH) {
      const int pos = int(gapRecs[3 * r + 2] >> 1);
      int open, ext;
      if (inA) {  // '-' in the ref: moves along row `pos`
        open = pos == 0 ? sc.goLR : (pos == la ? sc.goRR : sc.go);
        ext = pos == 0 ? sc.geLR : (pos == la ? sc.geRR : sc.ge);
      } else {    // '-' in the read: moves along column `pos`
        open = pos == 0 ? sc.goLQ : (pos == lb ? sc.goRQ : sc.go);
        ext = pos == 0 ? sc.geLQ : (pos == lb ? sc.geRQ : sc.ge);
      }
      gapCost += open + (size - 1) * ext;
    }
    const bool endGap = (startPos + size >= alnLen) || startPos == 0;
    if (endGap) {
      (inA ? endA : endB) += size;
      if (!sc.countEndGaps) continue;
    } else {
      (inA ? regA : regB) += size;
    }
    ++numGaps;
    const double wgt = sc.weighHomopolymers ? gapWeightDev(gA, gB, alnLen, startPos, size, inA) : 1.0;
    if (size >= 3) {
      large += wgt;
    } else if (size == 2) {
      two += wgt;
    } else {
      one += wgt;
    }
  }

  if (SCORE_FROM_PATH) score = int(warpSum(unsigned(matchSum))) - gapCost;
  if constexpr (CHI) {
    ChiArgs ca;
    ca.primaryQual = sc.primaryQual, ca.secondaryQual = sc.secondaryQual, ca.window = sc.window;
    ca.countEndGaps = sc.countEndGaps, ca.weighHomopolymers = sc.weighHomopolymers;
    ca.flags = a.flags, ca.chiKeepLq = a.chiKeepLq, ca.idx = a.idx, ca.chiOut = a.chiOut;
    ca.one = a.chiPar.oneBaseIndel, ca.two = a.chiPar.twoBaseIndel, ca.large = a.chiPar.largeBaseIndel;
    ca.hq = a.chiPar.hqMismatches, ca.lq = a.chiPar.lqMismatches, ca.lk = a.chiPar.lowKmerMismatches;
    chimeraProbeDev(ca, pv, gA, gB, alnLen, alnStart, gapRecs, int(nRecs), sTab, score, lane);
  }

  // comparison::passErrorProfile against every par-file line at once: lane l judges lines l and
  // l + 32, two ballots give the 64-bit verdict mask the host collapse loop consumes
  if (a.passOut || a.hits.recs) {
    const unsigned events = hqM + lqM + hq + lq + lk + numGaps;
    const double id = events ? double(hqM + lqM) / double(events) : 0.0;
    const double idHq = events ? double(hqM + lqM + lq + lk) / double(events) : 0.0;
    unsigned halves[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned li = unsigned(lane) + 32u * h;
      bool pass = false;
      if (li < a.nLines) {
        const sdgpu_iter_par it = a.lines[li];
        if (it.onPerId) {
          pass = (it.onHqPerId ? idHq : id) >= it.perId;
        } else {
          pass = it.oneBaseIndel >= one && it.twoBaseIndel >= two && it.largeBaseIndel >= large &&
                 it.hqMismatches >= double(hq) && it.lqMismatches >= double(lq) && it.lowKmerMismatches >= double(lk);
        }
      }
      halves[h] = __ballot_sync(FULL, pass);
    }
    if (lane == 0) {
      const uint64_t mask = (uint64_t(halves[1]) << 32) | halves[0];
      if (a.passOut) a.passOut[pair] = mask;
      if (a.hits.recs && mask) sd_emit_hit(a.src, a.hits, pair, mask);
    }
  }
  if (lane == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(a.work + 2), (unsigned long long)la * (unsigned long long)lb);
    if (a.out) {
      sdgpu_comparison c;
      c.alnScore = score;
      c.hqMismatches = hq;
      c.lqMismatches = lq;
      c.lowKmerMismatches = lk;
      c.highQualityMatches = hqM;
      c.lowQualityMatches = lqM;
      c.numGaps = numGaps;
      c.basesInAln = unsigned(alnLen) - endA - endB - regA - regB;
      c.overLappingEvents = hqM + lqM + hq + lq + lk + numGaps;
      c.alnLen = unsigned(alnLen);
      c.nGapRuns = nRecs;
      c.status = (a.gapsOut && nRecs > a.gapCap) ? SDGPU_ST_GAPLIST_TRUNCATED : 0u;
      c.oneBaseIndel = one;
      c.twoBaseIndel = two;
      c.largeBaseIndel = large;
      const double coverageArea = double(alnLen) - double(endA) - double(endB);
      c.refCoverage = (coverageArea - double(regA)) / double(la);
      c.queryCoverage = (coverageArea - double(regB)) / double(lb);
      if (c.overLappingEvents == 0) {
        c.eventBasedIdentity = 0;
        c.eventBasedIdentityHq = 0;
      } else {
        c.eventBasedIdentity = double(hqM + lqM) / double(c.overLappingEvents);
        c.eventBasedIdentityHq = double(hqM + lqM + lq + lk) / double(c.overLappingEvents);
      }
      a.out[pair] = c;
    }
  }
  __syncwarp();
}

__device__ __forceinline__ PairView viewOf(const AlignArgs& a, uint32_t pair) {
  PairView pv;
  pv.pair = pair;
  uint32_t ia, ib;
  sd_pair_of(a.src, pair, ia, ib);
  const uint64_t offA = a.seqs.offsets[ia], offB = a.seqs.offsets[ib];
  pv.la = int(a.seqs.offsets[ia + 1] - offA);
  pv.lb = int(a.seqs.offsets[ib + 1] - offB);
  pv.cA = a.seqs.codes + offA;
  pv.cB = a.seqs.codes + offB;
  pv.qA = a.seqs.quals + offA;
  pv.qB = a.seqs.quals + offB;
  pv.fA = a.seqs.flags + offA;
  pv.fB = a.seqs.flags + offB;
  return pv;
}

// =====================================================================================
// int32 forward pass, one pair per warp.  WIDE = reads wider than 32*C columns: the columns are
// strip-mined in blocks of 32*C; lane 31 leaves its per-row toRight / toDiag in an edge buffer and
// lane 0 of the next block picks them up instead of the closed-form column-0 boundary.
// =====================================================================================
template <int C, bool WIDE, bool CHI>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) alignProfileKernel(const AlignArgs a) {
  using Bits = BitsI32<C>;
  constexpr int CW = Bits::CW;
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  __syncthreads();
  const char* sTabBytes = reinterpret_cast<const char*>(sTab);

  const int lane = threadIdx.x & 31;
  const int warpInCta = threadIdx.x >> 5;
  const size_t warpGlobal = size_t(blockIdx.x) * WARPS_PER_CTA + warpInCta;
  uint8_t* slab = a.scratch + warpGlobal * a.slabBytes;
  uint32_t* ptrWords = reinterpret_cast<uint32_t*>(slab);
  uint8_t* alnA = slab + a.ptrBytes;
  uint8_t* alnB = alnA + a.maxAln;
  uint32_t* gapRecs = reinterpret_cast<uint32_t*>(alnB + a.maxAln);  // (colAbs, size, inA<<0 | pos<<1) triples
  int2* edges = reinterpret_cast<int2*>(slab + a.edgeOff);           // WIDE: two ping-pong arrays of a.edgeLen entries

  const DevScoring& sc = a.sc;
  const int dInt = sc.go - sc.ge;

  const uint32_t nWork = a.nWorkDev ? *a.nWorkDev : a.nPairs;
  for (;;) {
    uint32_t pair = 0;
    if (lane == 0) pair = atomicAdd(&a.work[0], 1u);
    pair = __shfl_sync(FULL, pair, 0);
    if (pair >= nWork) break;
    if (a.pairMap) pair = a.pairMap[pair];
    const PairView pv = viewOf(a, pair);
    const int la = pv.la, lb = pv.lb;
    const uint8_t *cA = pv.cA, *cB = pv.cB;
    const int nBlocks = WIDE ? (lb + 32 * C - 1) / (32 * C) : 1;
    int score = 0;

    for (int blk = 0; blk < nBlocks; ++blk) {
      const int jBase = WIDE ? blk * 32 * C : 0;
      int U[C], T[C], dcol[C], gocol[C], colofs[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int j = jBase + lane * C + c + 1;
        const int code = (j <= lb) ? int(cB[j - 1]) : 0;
        colofs[c] = code * 4;
        const int L0 = -sc.goLR - (j - 1) * sc.geLR;  // leftInherit of row 0
        const bool lastCol = (j == lb);
        gocol[c] = lastCol ? sc.goRQ : sc.go;
        dcol[c] = lastCol ? (sc.goRQ - sc.geRQ) : dInt;
        // keep the per-column penalties in registers (ptxas otherwise rematerialises the
        // j == lb select in every cell: three extra ALU instructions per cell update)
        asm volatile("" : "+r"(gocol[c]), "+r"(dcol[c]));
        U[c] = L0 - gocol[c];  // upInherit of row 1 (njhseq's i == 1 case)
        T[c] = L0;             // feeds diagInherit of row 1, column j+1
      }
      int outT = T[C - 1], outR = 0;
      int prevRecvT = __shfl_up_sync(FULL, outT, 1);
      const int colsHere = WIDE ? min(lb - jBase, 32 * C) : lb;
      const int nLanes = (colsHere + C - 1) / C;
      const int steps = la + nLanes - 1;
      const int2* edgeIn = edges + size_t((blk + 1) & 1) * a.edgeLen;  // written by the previous block
      int2* edgeOut = edges + size_t(blk & 1) * a.edgeLen;
      const bool feedNext = WIDE && (blk + 1 < nBlocks);
      uint32_t* blkWords = ptrWords + (WIDE ? size_t(blk) * a.blockWords : 0);
      int aCode = (lane == 0) ? int(cA[0]) : 0;  // base of the row this lane works on at t = 1
      for (int t = 1; t <= steps; ++t) {
        const int recvR = __shfl_up_sync(FULL, outR, 1);
        const int recvT = __shfl_up_sync(FULL, outT, 1);
        const int i = t - lane;
        const int iNext = i + 1;  // prefetch next row's base
        const int aNext = (iNext >= 1 && iNext <= la) ? int(cA[iNext - 1]) : 0;
        if (lane < nLanes && i >= 1 && i <= la) {
          const bool lastRow = (i == la);
          const int gorow = lastRow ? sc.goRR : sc.go;
          const int drow = lastRow ? (sc.goRR - sc.geRR) : dInt;
          int Lin, Tin;
          if (lane == 0) {
            if (WIDE && blk > 0) {
              Lin = edgeIn[i].x;                                              // toRight(i, jBase)
              Tin = (i == 1) ? (-sc.goLR - (jBase - 1) * sc.geLR) : edgeIn[i - 1].y;  // toDiag(i-1, jBase)
            } else {
              const int bnd = -sc.goLQ - (i - 1) * sc.geLQ;  // upInherit of column 0
              Lin = bnd - gorow;                             // njhseq's j == 1 case
              Tin = (i == 1) ? 0 : bnd + sc.geLQ;
            }
          } else {
            Lin = recvR;
            Tin = prevRecvT;
          }
          const char* rowTab = sTabBytes + aCode * (SDGPU_MAX_SYMBOLS * 4);
          uint32_t w[CW];
#pragma unroll
          for (int x = 0; x < CW; ++x) w[x] = 0;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const int s = *reinterpret_cast<const int*>(rowTab + colofs[c]);
            const int D = Tin + s;
            Tin = T[c];
            const int Uc = U[c];
            uint32_t acc = w[c / Bits::CPW];
            // maxima on the ALU pipe (two of them fused add+max, VIADDMNMX), differences on the
            // FMA pipe (IMAD.IADD); the shifted differences reuse d1/d2 instead of the sums
            const int W = max(Lin, D);
            const int d1 = Lin - D;
            acc = __funnelshift_l(uint32_t(d1), acc, 1);
            T[c] = max(Uc, W);
            const int d2 = Uc - W;
            acc = __funnelshift_l(uint32_t(d2), acc, 1);
            U[c] = __viaddmax_s32(Uc, dcol[c], W) - gocol[c];
            acc = __funnelshift_l(uint32_t(d2 + dcol[c]), acc, 1);
            const int tt = __viaddmax_s32(Lin, drow, D);
            acc = __funnelshift_l(uint32_t(d1 + drow), acc, 1);
            Lin = max(Uc, tt) - gorow;
            acc = __funnelshift_l(uint32_t(Uc - tt), acc, 1);
            w[c / Bits::CPW] = acc;
          }
          outR = Lin;
          outT = T[C - 1];
          uint32_t* dst = blkWords + size_t(t - 1) * (CW * 32) + lane;
#pragma unroll
          for (int x = 0; x < CW; ++x) dst[x * 32] = w[x];
          if (feedNext && lane == 31) edgeOut[i] = make_int2(outR, outT);
        }
        prevRecvT = recvT;
        aCode = aNext;
      }
      if (blk + 1 == nBlocks) {  // score = toDiag of the bottom-right cell
        const int lastLane = ((lb - 1) / C) & 31, lastC = (lb - 1) % C;
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (c == lastC) score = T[c];
        score = __shfl_sync(FULL, score, lastLane);
      }
      __syncwarp();
    }
    finishPair<false, CHI>(a, Bits{ptrWords, size_t(a.blockWords)}, viewOf(a, pair), score, alnA, alnB, gapRecs, sTab, lane);
  }
}

// =====================================================================================
// half2 forward pass, two pairs per warp (pair X in the low halves, pair Y in the high halves)
// =====================================================================================
__device__ __forceinline__ uint32_t h2bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 h2from(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }
// (x, y) as exact fp16 integers, packed with integer ops so the pair lives in ONE register
__device__ __forceinline__ __half2 h2of(int x, int y) {
  const uint32_t lo = __half_as_ushort(__int2half_rn(x)), hi = __half_as_ushort(__int2half_rn(y));
  return h2from(lo | (hi << 16));
}

template <int C>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) alignProfileKernelH2(const AlignArgs a) {
  using Bits = BitsH2<C>;
  constexpr int CW = Bits::CW;
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  __shared__ __half2 sTab2[H2_SYMS * H2_SYMS * H2_SYMS * H2_SYMS];  // [aX*5+aY][bX*5+bY] = (s(aX,bX), s(aY,bY))
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  for (int x = threadIdx.x; x < H2_SYMS * H2_SYMS * H2_SYMS * H2_SYMS; x += blockDim.x) {
    const int r = x / (H2_SYMS * H2_SYMS), q = x - r * (H2_SYMS * H2_SYMS);
    const int aX = r / H2_SYMS, aY = r - aX * H2_SYMS, bX = q / H2_SYMS, bY = q - bX * H2_SYMS;
    sTab2[x] = h2of(a.sc.score[aX * SDGPU_MAX_SYMBOLS + bX], a.sc.score[aY * SDGPU_MAX_SYMBOLS + bY]);
  }
  __syncthreads();
  const char* sTab2Bytes = reinterpret_cast<const char*>(sTab2);

  const int lane = threadIdx.x & 31;
  const int warpInCta = threadIdx.x >> 5;
  const size_t warpGlobal = size_t(blockIdx.x) * WARPS_PER_CTA + warpInCta;
  uint8_t* slab = a.scratch + warpGlobal * a.slabBytes;
  uint32_t* ptrWords = reinterpret_cast<uint32_t*>(slab);
  uint8_t* alnA = slab + a.ptrBytes;
  uint8_t* alnB = alnA + a.maxAln;
  uint32_t* gapRecs = reinterpret_cast<uint32_t*>(alnB + a.maxAln);

  const DevScoring& sc = a.sc;
  const int dInt = sc.go - sc.ge;
  const uint32_t nWork = a.nWorkDev ? *a.nWorkDev : a.nPairs;
  const uint32_t nItems = (nWork + 1) / 2;

  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(&a.work[0], 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= nItems) break;
    const bool haveY = 2 * item + 1 < nWork;
    const uint32_t pairX = a.pairMap ? a.pairMap[2 * item] : 2 * item;
    const uint32_t pairY = haveY ? (a.pairMap ? a.pairMap[2 * item + 1] : 2 * item + 1) : pairX;
    const PairView px = viewOf(a, pairX);
    const PairView py = viewOf(a, pairY);
    const int laM = max(px.la, py.la), lbM = max(px.lb, py.lb);

    __half2 U[C], T[C];
    uint32_t dcolU[C], gocolU[C];  // packed fp16 pairs, held as raw registers
    int colofs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = lane * C + c + 1;
      const int codeX = (j <= px.lb) ? int(px.cB[j - 1]) : 0;
      const int codeY = (j <= py.lb) ? int(py.cB[j - 1]) : 0;
      colofs[c] = (codeX * H2_SYMS + codeY) * 4;
      const int L0 = -sc.goLR - (j - 1) * sc.geLR;
      const bool lastX = (j == px.lb), lastY = (j == py.lb);
      gocolU[c] = h2bits(h2of(lastX ? sc.goRQ : sc.go, lastY ? sc.goRQ : sc.go));
      dcolU[c] = h2bits(h2of(lastX ? (sc.goRQ - sc.geRQ) : dInt, lastY ? (sc.goRQ - sc.geRQ) : dInt));
      asm volatile("" : "+r"(gocolU[c]), "+r"(dcolU[c]));  // keep the per-column penalties in registers
      const __half2 L0h = h2of(L0, L0);
      U[c] = __hsub2(L0h, h2from(gocolU[c]));
      T[c] = L0h;
    }
    __half2 outT = T[C - 1], outR = h2of(0, 0);
    __half2 prevRecvT = __shfl_up_sync(FULL, outT, 1);
    const int nLanes = (lbM + C - 1) / C;
    const int steps = laM + nLanes - 1;
    int rowCode = (lane == 0) ? (int(px.cA[0]) * H2_SYMS + int(py.cA[0])) : 0;
    // per-row penalties for the four (last row of X?, last row of Y?) cases, converted once
    const int dRR = sc.goRR - sc.geRR;
    const __half2 hGo = h2of(sc.go, sc.go), hD = h2of(dInt, dInt);
    const __half2 hGoX = h2of(sc.goRR, sc.go), hDX = h2of(dRR, dInt);
    const __half2 hGoY = h2of(sc.go, sc.goRR), hDY = h2of(dInt, dRR);
    const __half2 hGoXY = h2of(sc.goRR, sc.goRR), hDXY = h2of(dRR, dRR);
    // lane 0's boundary feed (upInherit of column 0), kept as a running fp16 pair
    const __half2 hGeLQ = h2of(sc.geLQ, sc.geLQ);
    __half2 bndH = h2of(-sc.goLQ, -sc.goLQ);  // value for row 1
    const int laX = px.la, laY = py.la;
    const uint8_t *cAx = px.cA, *cAy = py.cA;
    for (int t = 1; t <= steps; ++t) {
      const __half2 recvR = __shfl_up_sync(FULL, outR, 1);
      const __half2 recvT = __shfl_up_sync(FULL, outT, 1);
      const int i = t - lane;
      const int iNext = i + 1;
      const int aXn = (iNext >= 1 && iNext <= laX) ? int(cAx[iNext - 1]) : 0;
      const int aYn = (iNext >= 1 && iNext <= laY) ? int(cAy[iNext - 1]) : 0;
      if (lane < nLanes && i >= 1 && i <= laM) {
        const bool lastRowX = (i == laX), lastRowY = (i == laY);
        const __half2 gorow = lastRowX ? (lastRowY ? hGoXY : hGoX) : (lastRowY ? hGoY : hGo);
        const __half2 drow = lastRowX ? (lastRowY ? hDXY : hDX) : (lastRowY ? hDY : hD);
        __half2 Lin, Tin;
        if (lane == 0) {
          Lin = __hsub2(bndH, gorow);                                // njhseq's j == 1 case
          Tin = (i == 1) ? h2of(0, 0) : __hadd2(bndH, hGeLQ);        // upInherit of (i-1, 0)
          bndH = __hsub2(bndH, hGeLQ);
        } else {
          Lin = recvR;
          Tin = prevRecvT;
        }
        const char* rowTab = sTab2Bytes + rowCode * (H2_SYMS * H2_SYMS * 4);
        uint32_t w[CW];
#pragma unroll
        for (int x = 0; x < CW; ++x) w[x] = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const __half2 s = *reinterpret_cast<const __half2*>(rowTab + colofs[c]);
          const __half2 D = __hadd2(Tin, s);
          Tin = T[c];
          const __half2 Uc = U[c];
          uint32_t acc = w[c / Bits::CPW];
          const __half2 W = __hmax2(Lin, D);
          const __half2 d1 = __hsub2(Lin, D);
          acc = (acc >> 1) | (h2bits(d1) & 0x80008000u);
          T[c] = __hmax2(Uc, W);
          const __half2 d2 = __hsub2(Uc, W);
          acc = (acc >> 1) | (h2bits(d2) & 0x80008000u);
          const __half2 dc = h2from(dcolU[c]);
          U[c] = __hsub2(__hmax2(__hadd2(Uc, dc), W), h2from(gocolU[c]));
          acc = (acc >> 1) | (h2bits(__hadd2(d2, dc)) & 0x80008000u);
          const __half2 tt = __hmax2(__hadd2(Lin, drow), D);
          acc = (acc >> 1) | (h2bits(__hadd2(d1, drow)) & 0x80008000u);
          Lin = __hsub2(__hmax2(Uc, tt), gorow);
          acc = (acc >> 1) | (h2bits(__hsub2(Uc, tt)) & 0x80008000u);
          w[c / Bits::CPW] = acc;
        }
        outR = Lin;
        outT = T[C - 1];
        uint32_t* dst = ptrWords + size_t(t - 1) * (CW * 32) + lane;
#pragma unroll
        for (int x = 0; x < CW; ++x) dst[x * 32] = w[x];
      }
      prevRecvT = recvT;
      rowCode = aXn * H2_SYMS + aYn;
    }
    __syncwarp();
    // The shorter pair's registers keep running on padding rows, so the corner values are not kept:
    // each pair's score is re-derived from its traced alignment (finishPair<true>).
    // (views are re-read here so their eight pointers do not stay live across the DP loop)
    finishPair<true, false>(a, Bits{ptrWords, 0}, viewOf(a, pairX), 0, alnA, alnB, gapRecs, sTab, lane);
    if (haveY) finishPair<true, false>(a, Bits{ptrWords, 1}, viewOf(a, pairY), 0, alnA, alnB, gapRecs, sTab, lane);
  }
}

// =====================================================================================
// Banded first pass.  Most (cluster, read) pairs of a collapse are near-identical, so the optimal
// alignment hugs the main diagonal.  This kernel evaluates only the 64 diagonals (offsets j - i)
// centred between the two corner diagonals, anti-diagonal by anti-diagonal: lane l owns offsets
// dlo + 2l and dlo + 2l + 1 and updates one cell per step; the up / left inputs come from its own
// previous cell or, every other step, from one neighbour (one shuffle per step).  Same recurrence,
// same decision bits, out-of-band inputs = -inf.
// Exactness: the banded optimum S_w is reached by a real alignment, so S_w <= S_opt.  An alignment
// that touches offset hi + 1 has at least hi + 1 right moves, hence at most lb - hi - 1 diagonal
// moves, and pays at least the cheapest non-right-end gap column for each of them: its score is
// <= maxSub * (lb - hi - 1) - minRight * (hi + 1); likewise below the band with la, lo and minDown.  If S_w exceeds both bounds, every optimal
// alignment -- and every tie the traceback's decision bits could see along it -- lies inside the
// band, where banded and full values coincide: score, path and errorProfile are the full DP's.
// Otherwise the pair goes to the redo list and the full-matrix kernel recomputes it.
// =====================================================================================
constexpr int BAND_NEG = -(1 << 28);
constexpr int BAND_MAX_DIFF = 30;  // |lb - la| beyond which the 64-diagonal band is not tried

struct BandLane {  // one lane's running state: two offsets, outputs of its previous cell, bit accumulator
  int Tp0, Tp1;             // toDiag of the previous cell on offset d0 / d0 + 1
  int lastDown, lastRight;  // toDown / toRight of the cell this lane updated one step ago
  uint32_t acc;
};

// One cell update.  ODD: the lane's odd offset d0 + 1 (inputs: left from its own previous cell, up from
// the right neighbour), else d0 (left from the left neighbour, up from its own previous cell).
// MODE: BODY = interior step (every lane's cell has 2 <= i < la, 2 <= j < lb); HEAD = steps before the
// body (border rows / columns and cells above / left of the matrix possible, last row / column not);
// TAIL = steps after it (last-row / last-column penalties and cells below / right of the matrix);
// GEN = everything (matrices too small to have a body).
enum { BAND_BODY = 0, BAND_HEAD = 1, BAND_TAIL = 2, BAND_GEN = 3 };
template <bool ODD, int MODE>
__device__ __forceinline__ void bandCell(BandLane& st, int i, int j, int lane, int la, int lb, int s,
                                         const DevScoring& sc, int dInt) {
  constexpr bool LOW = (MODE & BAND_HEAD) != 0, HIGH = (MODE & BAND_TAIL) != 0;
  int recv = ODD ? __shfl_down_sync(FULL, st.lastDown, 1) : __shfl_up_sync(FULL, st.lastRight, 1);
  if (ODD ? (lane == 31) : (lane == 0)) recv = BAND_NEG;  // beyond the band
  if ((LOW && (i < 1 || j < 1)) || (HIGH && (i > la || j > lb))) {
    st.acc <<= 5;
    return;
  }
  int Lin = ODD ? st.lastRight : recv;  // toRight of (i, j-1): offset d-1
  int Uin = ODD ? recv : st.lastDown;   // toDown of (i-1, j): offset d+1
  int Tin = ODD ? st.Tp1 : st.Tp0;      // toDiag of (i-1, j-1): same offset, two steps ago
  int gorow = sc.go, drow = dInt, gocol = sc.go, dcol = dInt;
  if (HIGH) {
    if (i == la) gorow = sc.goRR, drow = sc.goRR - sc.geRR;
    if (j == lb) gocol = sc.goRQ, dcol = sc.goRQ - sc.geRQ;
  }
  if (LOW) {
    if (i == 1) {  // row 0 border (njhseq's i == 1 case)
      const int r0 = -sc.goLR - (j - 1) * sc.geLR;
      Uin = r0 - gocol;
      Tin = (j == 1) ? 0 : r0 + sc.geLR;
    }
    if (j == 1) {  // column 0 border (njhseq's j == 1 case)
      const int c0 = -sc.goLQ - (i - 1) * sc.geLQ;
      Lin = c0 - gorow;
      if (i != 1) Tin = c0 + sc.geLQ;
    }
  }
  uint32_t acc = st.acc;
  const int D = Tin + s;
  const int W = max(Lin, D);
  const int d1 = Lin - D;
  acc = __funnelshift_l(uint32_t(d1), acc, 1);
  const int T = max(Uin, W);
  const int d2 = Uin - W;
  acc = __funnelshift_l(uint32_t(d2), acc, 1);
  st.lastDown = __viaddmax_s32(Uin, dcol, W) - gocol;
  acc = __funnelshift_l(uint32_t(d2 + dcol), acc, 1);
  const int tt = __viaddmax_s32(Lin, drow, D);
  acc = __funnelshift_l(uint32_t(d1 + drow), acc, 1);
  st.lastRight = max(Uin, tt) - gorow;
  acc = __funnelshift_l(uint32_t(Uin - tt), acc, 1);
  st.acc = acc;
  if (ODD) {
    st.Tp1 = T;
  } else {
    st.Tp0 = T;
  }
}

// Sequences up to BAND_STAGE bases are staged per warp in shared memory as ready-made byte offsets
// into the score table (ref: code * 64 = its row, read: code * 4 = its column), so a cell's score is
// one LDS at base + row + column.
constexpr int BAND_STAGE = 2048;
constexpr int BAND_TABW = 20;  // row stride (words) of the band kernel's score table: the rows of A, C, G, T fall on disjoint banks
struct BandCodes {
  const uint16_t* rowOff;  // [i-1] -> byte offset of ref base i's row (shared memory, or nullptr)
  const uint8_t* colOff;   // [j-1] -> byte offset of read base j's column
  const uint8_t *cA, *cB;  // global codes (sequences too long to stage)
  const char* tab;         // score table, bytes
  __device__ __forceinline__ int row(int i) const {
    return rowOff ? int(rowOff[i - 1]) : int(__ldg(cA + i - 1)) * (BAND_TABW * 4);
  }
  __device__ __forceinline__ int col(int j) const { return rowOff ? int(colOff[j - 1]) : int(__ldg(cB + j - 1)) * 4; }
  __device__ __forceinline__ int score(int r, int c) const { return *reinterpret_cast<const int*>(tab + r + c); }
};

// three step pairs (one word of decision bits) outside the body
template <int MODE>
__device__ __forceinline__ void bandWordEdge(BandLane& st, int& i, int& j, int lane, int la, int lb, const BandCodes& bc,
                                             const DevScoring& sc, int dInt) {
#pragma unroll
  for (int k = 0; k < 3; ++k, ++i, ++j) {
    const int r = (i >= 1 && i <= la) ? bc.row(i) : 0;
    const int s0 = bc.score(r, (j >= 1 && j <= lb) ? bc.col(j) : 0);
    const int s1 = bc.score(r, (j + 1 >= 1 && j + 1 <= lb) ? bc.col(j + 1) : 0);
    bandCell<false, MODE>(st, i, j, lane, la, lb, s0, sc, dInt);
    bandCell<true, MODE>(st, i, j + 1, lane, la, lb, s1, sc, dInt);
  }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 8) alignBandKernel(const AlignArgs a) {
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  __shared__ uint16_t sRow[WARPS_PER_CTA][BAND_STAGE];
  __shared__ uint8_t sCol[WARPS_PER_CTA][BAND_STAGE];
  __shared__ int sTabW[SDGPU_MAX_SYMBOLS * BAND_TABW];  // the DP's copy of the table (sTab serves the errorProfile stage)
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * BAND_TABW; x += blockDim.x) {
    const int r = x / BAND_TABW, c = x - r * BAND_TABW;
    sTabW[x] = c < SDGPU_MAX_SYMBOLS ? int(a.sc.score[r * SDGPU_MAX_SYMBOLS + c]) : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warpInCta = threadIdx.x >> 5;
  const size_t warpGlobal = size_t(blockIdx.x) * WARPS_PER_CTA + warpInCta;
  uint8_t* slab = a.scratch + warpGlobal * a.slabBytes;
  uint32_t* ptrWords = reinterpret_cast<uint32_t*>(slab);
  uint8_t* alnA = slab + a.ptrBytes;
  uint8_t* alnB = alnA + a.maxAln;
  uint32_t* gapRecs = reinterpret_cast<uint32_t*>(alnB + a.maxAln);
  const DevScoring& sc = a.sc;
  const int dInt = sc.go - sc.ge;
  const uint32_t nWork = a.nWorkDev ? *a.nWorkDev : a.nPairs;  // behind the screen pass: its redo list
  if (blockIdx.x == 0 && threadIdx.x == 0) a.work[SD_W_BAND_PAIRS] = nWork;  // read back with the redo count

  for (;;) {
    uint32_t pair = 0;
    if (lane == 0) pair = atomicAdd(&a.work[0], 1u);
    pair = __shfl_sync(FULL, pair, 0);
    if (pair >= nWork) break;
    if (a.pairMap) pair = a.pairMap[pair];
    const PairView pv = viewOf(a, pair);
    const int la = pv.la, lb = pv.lb;
    const int diff = lb - la;
    bool certified = false;
    int score = 0;
    const int dlo = (diff - 63) >> 1;        // floor: the band [dlo, dlo + 63] is centred between offsets 0 and diff
    const int aBase = 2 - ((2 - dlo) & 1);   // lanes meet their even offset d0 on anti-diagonals aBase, aBase + 2, ...
    if (diff >= -BAND_MAX_DIFF && diff <= BAND_MAX_DIFF) {
      BandCodes bc{nullptr, nullptr, pv.cA, pv.cB, reinterpret_cast<const char*>(sTabW)};
      const bool staged = la <= BAND_STAGE && lb <= BAND_STAGE;
      if (staged) {
        for (int x = lane; x < la; x += 32) sRow[warpInCta][x] = uint16_t(int(pv.cA[x]) * (BAND_TABW * 4));
        for (int x = lane; x < lb; x += 32) sCol[warpInCta][x] = uint8_t(int(pv.cB[x]) * 4);
        bc.rowOff = sRow[warpInCta];
        bc.colOff = sCol[warpInCta];
      }
      __syncwarp();
      const int d0 = dlo + 2 * lane;
      // step pair p: anti-diagonals aBase + 2p (offset d0) and aBase + 2p + 1 (offset d0 + 1); both cells
      // of a lane sit in row i = I0 - lane + p, columns j = i + d0 and j + 1
      const int I0 = (aBase - dlo) >> 1;
      const int nPairSteps = (la + lb - aBase) / 2 + 1;
      const int nWords = (nPairSteps + 2) / 3;
      // interior pairs: for every lane 2 <= i <= la - 1 and 2 <= j, j + 1 <= lb - 1
      const int pBodyStart = max(max(33 - I0, 2 - I0 - dlo), 0);
      const int pBodyEnd = min(la - I0, lb - 32 - I0 - dlo);  // exclusive
      int wHead = (pBodyStart + 2) / 3, wBody = pBodyEnd > 0 ? pBodyEnd / 3 : 0;  // words [wHead, wBody) are all-interior
      if (!staged || wBody <= wHead) wHead = wBody = 0;
      BandLane st{0, 0, BAND_NEG, BAND_NEG, 0u};
      uint32_t* dst = ptrWords + lane;
      int i = I0 - lane, j = i + d0;
      if (wBody > wHead) {
        for (int w = 0; w < wHead; ++w, dst += 32) {  // head: borders possible, last row / column not
          st.acc = 0;
          bandWordEdge<BAND_HEAD>(st, i, j, lane, la, lb, bc, sc, dInt);
          *dst = st.acc;
        }
        const uint16_t* rp = bc.rowOff + (i - 1);
        const uint8_t* cp = bc.colOff + (j - 1);
        int c0 = int(cp[0]);
        for (int w = wHead; w < wBody; ++w, dst += 32, rp += 3, cp += 3) {  // body: lean cells
          st.acc = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int r = int(rp[k]);
            const int c1 = int(cp[k + 1]);
            bandCell<false, BAND_BODY>(st, 0, 0, lane, la, lb, bc.score(r, c0), sc, dInt);
            bandCell<true, BAND_BODY>(st, 0, 0, lane, la, lb, bc.score(r, c1), sc, dInt);
            c0 = c1;
          }
          *dst = st.acc;
        }
        i += 3 * (wBody - wHead);
        j += 3 * (wBody - wHead);
        for (int w = wBody; w < nWords; ++w, dst += 32) {  // tail: last row / column, cells past the matrix
          st.acc = 0;
          bandWordEdge<BAND_TAIL>(st, i, j, lane, la, lb, bc, sc, dInt);
          *dst = st.acc;
        }
      } else {
        for (int w = 0; w < nWords; ++w, dst += 32) {
          st.acc = 0;
          bandWordEdge<BAND_GEN>(st, i, j, lane, la, lb, bc, sc, dInt);
          *dst = st.acc;
        }
      }
      const int cornerOdd = (diff - dlo) & 1;
      score = __shfl_sync(FULL, cornerOdd ? st.Tp1 : st.Tp0, (diff - dlo) >> 1);  // toDiag of (la, lb)
      const int hi = dlo + 63;
      // (a path that touches offset hi + 1 does so above the last row, so its >= hi + 1 right moves are
      //  left-end or interior gap columns and cost at least minRight each; same below the band)
      const long long ubHigh = (hi + 1 <= lb) ? (long long)a.maxSub * (lb - hi - 1) - (long long)a.minRight * (hi + 1) : LLONG_MIN;
      const long long ubLow = (1 - dlo <= la) ? (long long)a.maxSub * (la + dlo - 1) - (long long)a.minDown * (1 - dlo) : LLONG_MIN;
      certified = (long long)score > (ubHigh > ubLow ? ubHigh : ubLow);
      // cells of the band that lie inside the matrix (what this pass actually evaluated)
      unsigned cells = 0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int d = d0 + e;
        const int n = min(la, lb - d) - max(1, 1 - d) + 1;
        cells += n > 0 ? unsigned(n) : 0u;
      }
      cells = warpSum(cells);
      if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(a.work + 6), (unsigned long long)cells);
    }
    __syncwarp();
    if (certified) {
      if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(a.work + 8), (unsigned long long)la * (unsigned long long)lb);
      finishPair<false, false>(a, BitsBand{ptrWords, dlo, aBase}, pv, score, alnA, alnB, gapRecs, sTab, lane);
    } else if (lane == 0) {
      a.redo[atomicAdd(&a.work[4], 1u)] = pair;
    }
  }
}

template <class Kernel>
int launchKernel(sdgpu_ctx* ctx, Kernel kernel, AlignArgs& args, uint32_t workItems, int CW, uint32_t maxLenA,
                 uint32_t maxLenB, uint32_t nBlocks = 1, uint64_t bandBitBytes = 0) {
  int ctasPerSM = 0;
  SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, kernel, WARPS_PER_CTA * 32, 0));
  if (ctasPerSM < 1) ctasPerSM = 1;
  uint32_t grid = uint32_t(ctx->numSMs) * uint32_t(ctasPerSM);
  const uint32_t needCtas = (workItems + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  if (grid > needCtas) grid = needCtas;
  const uint64_t maxSteps = uint64_t(maxLenA) + 31;
  const uint64_t blockBytes = bandBitBytes ? bandBitBytes : maxSteps * CW * 128;
  if (blockBytes * nBlocks >= (1ull << 32)) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "alignment too large (traceback bits >= 4 GiB per pair)");
  args.blockWords = blockBytes / 4;
  args.ptrBytes = uint32_t(blockBytes * nBlocks);
  args.maxAln = (maxLenA + maxLenB + 15) & ~15u;
  uint64_t slab = uint64_t(args.ptrBytes) + 2ull * args.maxAln + 12ull * (uint64_t(args.maxAln) + 2);
  slab = (slab + 15) & ~15ull;
  args.edgeOff = uint32_t(slab);
  args.edgeLen = maxLenA + 2;
  if (nBlocks > 1) slab += 2ull * args.edgeLen * sizeof(int2);
  slab = (slab + 255) & ~255ull;
  args.slabBytes = slab;
  // wide alignments: keep the resident slabs within a fixed HBM budget by running fewer warps
  const uint64_t budget = 48ull << 30;
  if (slab * grid * WARPS_PER_CTA > budget) grid = uint32_t(std::max<uint64_t>(1, budget / (slab * WARPS_PER_CTA)));
  int rc = sd_ensure_scratch(ctx, slab * grid * WARPS_PER_CTA);
  if (rc) return rc;
  args.scratch = ctx->dScratch;
  SD_CUDA(ctx, cudaMemsetAsync(ctx->dWork, 0, sizeof(uint32_t), ctx->stream));
  // per-launch device time (CUDA events on the launching stream) for the roofline bookkeeping
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timeKernels) {
    if (ctx->evUsed + 2 > ctx->evPool.size()) {
      const size_t old = ctx->evPool.size();
      ctx->evPool.resize(old + 256, nullptr);
      for (size_t x = old; x < ctx->evPool.size(); ++x) SD_CUDA(ctx, cudaEventCreate(&ctx->evPool[x]));
    }
    if (ctx->evBand.size() < ctx->evPool.size() / 2) ctx->evBand.resize(ctx->evPool.size() / 2, 0);
    ctx->evBand[ctx->evUsed / 2] = bandBitBytes ? 1 : 0;
    e0 = ctx->evPool[ctx->evUsed++];
    e1 = ctx->evPool[ctx->evUsed++];
    SD_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  }
  kernel<<<grid, WARPS_PER_CTA * 32, 0, ctx->stream>>>(args);
  SD_CUDA(ctx, cudaGetLastError());
  if (e1) SD_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ++ctx->launches;
  ++ctx->alignLaunches;
  return SDGPU_OK;
}

template <int C, bool CHI = false>
int launchI32(sdgpu_ctx* ctx, AlignArgs& args, uint32_t maxLenA, uint32_t maxLenB) {
  return launchKernel(ctx, alignProfileKernel<C, false, CHI>, args, args.nPairs, BitsI32<C>::CW, maxLenA, maxLenB);
}
template <bool CHI = false>
int launchWide(sdgpu_ctx* ctx, AlignArgs& args, uint32_t maxLenA, uint32_t maxLenB) {
  constexpr int C = 16;
  const uint32_t nBlocks = (maxLenB + 32 * C - 1) / (32 * C);
  return launchKernel(ctx, alignProfileKernel<C, true, CHI>, args, args.nPairs, BitsI32<C>::CW, maxLenA, maxLenB, nBlocks);
}
template <int C>
int launchH2(sdgpu_ctx* ctx, AlignArgs& args, uint32_t maxLenA, uint32_t maxLenB) {
  return launchKernel(ctx, alignProfileKernelH2<C>, args, (args.nPairs + 1) / 2, BitsH2<C>::CW, maxLenA, maxLenB);
}

// Every DP value of the launch (real and padding cells) must be an integer fp16 holds exactly and
// every difference of two of them too: |v| <= 2047 with room for the subtraction.
bool halfRangeOk(const sdgpu_ctx* ctx, uint32_t maxLenA, uint32_t cols) {
  const DevScoring& s = ctx->scoring;
  int maxPos = 0, maxNeg = 0;
  for (int a = 0; a < H2_SYMS; ++a)
    for (int b = 0; b < H2_SYMS; ++b) {
      const int v = s.score[a * SDGPU_MAX_SYMBOLS + b];
      maxPos = std::max(maxPos, v);
      maxNeg = std::max(maxNeg, -v);
    }
  const int maxOpen = std::max({s.go, s.goLQ, s.goRQ, s.goLR, s.goRR});
  const int maxExt = std::max({s.ge, s.geLQ, s.geRQ, s.geLR, s.geRR});
  const int64_t len = int64_t(maxLenA) + cols;
  const int64_t hi = int64_t(maxPos) * std::min<int64_t>(maxLenA, cols) + maxOpen;
  const int64_t lo = 3ll * maxOpen + int64_t(maxExt) * len + maxNeg + maxPos;  // magnitude of the most negative value
  return hi <= 1000 && lo <= 1040 && hi + lo <= 2040;
}

}  // namespace

static int ensureU32(sdgpu_ctx* ctx, uint32_t** buf, uint32_t* cap, uint32_t need) {
  if (need <= *cap) return SDGPU_OK;
  if (*buf) {
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(*buf);
  }
  *buf = nullptr;
  *cap = 0;
  const uint32_t n = std::max<uint32_t>(need, 1u << 16);
  SD_CUDA(ctx, cudaMalloc(buf, size_t(n) * sizeof(uint32_t)));
  *cap = n;
  return SDGPU_OK;
}

int sd_launch_align(sdgpu_ctx* ctx, const SdAlignReq& req) {
  const uint32_t nPairs = req.nPairs, flags = req.flags, maxLenA = req.maxLenA, maxLenB = req.maxLenB;
  const SdChimeraReq* chi = req.chi;
  if (nPairs == 0) return SDGPU_OK;
  if ((flags & SDGPU_CHECK_KMER) && !ctx->haveIndex)
    return sd_fail(ctx, SDGPU_E_STATE, "SDGPU_CHECK_KMER requires sdgpu_build_kmer_index first");
  const bool wantMask = req.dPassOut != nullptr || req.hits.recs != nullptr;
  if (wantMask && ctx->nLines == 0) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_set_pass_table has not been called");
  {  // the int32 passes keep every DP value far inside int32 (the fp16 and banded passes check their own ranges)
    const DevScoring& s = ctx->scoring;
    int64_t maxAbs = 0;
    for (int x = 0; x < ctx->nSym; ++x)
      for (int y = 0; y < ctx->nSym; ++y) maxAbs = std::max<int64_t>(maxAbs, std::abs(int(s.score[x * SDGPU_MAX_SYMBOLS + y])));
    const int64_t maxPen = std::max({s.go, s.ge, s.goLQ, s.geLQ, s.goRQ, s.geRQ, s.goLR, s.geLR, s.goRR, s.geRR});
    const int64_t worst = maxAbs * std::min(maxLenA, maxLenB) + maxPen * (int64_t(maxLenA) + maxLenB + 2);
    if (worst >= (int64_t(1) << 30)) return sd_fail(ctx, SDGPU_E_OVERFLOW, "penalties x lengths overflow the 32-bit DP values");
  }
  AlignArgs args;
  args.sc = ctx->scoring;
  args.seqs = sd_dev_seqs(ctx);
  args.idx = sd_dev_index(ctx);
  args.src = req.src;
  args.nPairs = nPairs;
  args.flags = flags;
  args.out = req.dOut;
  args.gapsOut = req.dGaps;
  args.gapCap = req.gapCap;
  args.work = ctx->dWork;
  args.lines = ctx->dLines;
  args.nLines = ctx->nLines;
  args.passOut = req.dPassOut;
  args.hits = req.hits;
  args.chiOut = chi ? chi->dOut : nullptr;
  args.chiPar = chi ? chi->overlap : sdgpu_iter_par{};
  args.chiKeepLq = chi ? chi->keepLq : 0u;
  args.pairMap = nullptr;
  args.nWorkDev = nullptr;
  args.redo = nullptr;
  args.maxSub = args.minRight = args.minDown = 0;
  const uint32_t needC = (maxLenB + 31) / 32;
  const bool gridSrc = req.src.clus != nullptr;
  // Screen pass (screen_kernel.cu): verdict-mask launches only.  A grid launch always starts with its scan
  // kernel, which at least lists the pairs that are not a cluster against itself.
  const bool maskOnly = wantMask && !req.dOut && !req.dGaps && !chi;
  SdScreenArgs scr{};
  const bool screen = maskOnly && ctx->screenMode == 0 && ctx->bandMode == 0 && ctx->forceKernel == 0 && nPairs >= 64 &&
                      sd_screen_plan(ctx, maxLenA, maxLenB, &scr);
  bool behindScreen = false;
  if (screen || gridSrc) {
    int rc = ensureU32(ctx, &ctx->dRedo, &ctx->redoCap, nPairs);
    if (rc) return rc;
    if (screen && nPairs > ctx->screenItemsCap) {
      if (ctx->dScreenItems) {
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dScreenItems);
      }
      ctx->dScreenItems = nullptr;
      ctx->screenItemsCap = 0;
      const uint64_t cap = std::max<uint64_t>(nPairs, 1u << 16);
      SD_CUDA(ctx, cudaMalloc(&ctx->dScreenItems, 3 * cap * sizeof(SdScreenItem)));
      ctx->screenItemsCap = cap;
    }
    SD_CUDA(ctx, cudaMemsetAsync(ctx->dWork + SD_W_SCAN, 0, 8 * sizeof(uint32_t), ctx->stream));
    scr.sc = ctx->scoring;
    scr.seqs = args.seqs;
    scr.idx = args.idx;
    scr.src = req.src;
    scr.nPairs = nPairs;
    scr.flags = flags;
    scr.lines = ctx->dLines;
    scr.nLines = ctx->nLines;
    scr.passOut = req.dPassOut;
    scr.hits = req.hits;
    scr.items = ctx->dScreenItems;
    scr.redo = ctx->dRedo;
    scr.work = ctx->dWork;
    scr.screenOn = screen ? 1 : 0;
    if ((rc = sd_launch_screen(ctx, scr))) return rc;
    if (screen) ctx->screenPairs += nPairs;
    args.pairMap = ctx->dRedo;  // everything below runs over the screen's redo list
    args.nWorkDev = ctx->dWork + SD_W_BIN0 + 3;
    behindScreen = true;
  }
  // Banded first pass + exact redo of the uncertified pairs (see alignBandKernel).  Not used when a
  // test pins a specific full-matrix kernel, for chimera probes, or when the last banded launches
  // sent most of their pairs to the redo list (dissimilar sequences); then it is re-tried every 16th launch.
  bool band = ctx->bandMode == 0 && ctx->forceKernel == 0 && !chi && nPairs >= 64 && maxLenA + maxLenB <= 16384;
  if (band) {
    const DevScoring& s = ctx->scoring;
    int maxSub = INT_MIN, maxAbs = 0;
    for (int x = 0; x < ctx->nSym; ++x)
      for (int y = 0; y < ctx->nSym; ++y) {
        const int v = s.score[x * SDGPU_MAX_SYMBOLS + y];
        maxSub = std::max(maxSub, v);
        maxAbs = std::max(maxAbs, std::abs(v));
      }
    const int pens[10] = {s.go, s.ge, s.goLQ, s.geLQ, s.goRQ, s.geRQ, s.goLR, s.geLR, s.goRR, s.geRR};
    for (int v : pens) maxAbs = std::max(maxAbs, v);  // penalties are validated >= 0 at set_params
    band = maxSub > 0 && maxAbs <= 4096;
    args.maxSub = maxSub;
    args.minRight = std::min(std::min(s.go, s.ge), std::min(s.goLR, s.geLR));
    args.minDown = std::min(std::min(s.go, s.ge), std::min(s.goLQ, s.geLQ));
  }
  // (the statistics of the last banded launch are only looked at once their copy has landed; behind the
  //  screen pass the band sees the pairs with gaps, which is what it is for: no skipping there)
  if (band && !behindScreen && ctx->bandStatEv && cudaEventQuery(ctx->bandStatEv) == cudaSuccess && ctx->hBandStat[1] >= 64 &&
      2ull * ctx->hBandStat[0] > ctx->hBandStat[1] && ++ctx->bandSkipped < 16)
    band = false;
  cudaGetLastError();  // cudaEventQuery's cudaErrorNotReady is not an error
  if (band) {
    uint32_t** redoBuf = behindScreen ? &ctx->dRedo2 : &ctx->dRedo;
    int rc = ensureU32(ctx, redoBuf, behindScreen ? &ctx->redo2Cap : &ctx->redoCap, nPairs);
    if (rc) return rc;
    if (!behindScreen) ctx->bandSkipped = 0;
    args.redo = *redoBuf;
    SD_CUDA(ctx, cudaMemsetAsync(ctx->dWork + SD_W_REDO, 0, sizeof(uint32_t), ctx->stream));
    const uint64_t bitBytes = (uint64_t(maxLenA + maxLenB) / 6 + 2) * 128;
    rc = launchKernel(ctx, alignBandKernel, args, nPairs, 0, maxLenA, maxLenB, 1, bitBytes);
    if (rc) return rc;
    if (!behindScreen) {
      if (!ctx->bandStatEv) SD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->bandStatEv, cudaEventDisableTiming));
      SD_CUDA(ctx, cudaMemcpyAsync(ctx->hBandStat, ctx->dWork + SD_W_REDO, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
      SD_CUDA(ctx, cudaEventRecord(ctx->bandStatEv, ctx->stream));
    }
    args.redo = nullptr;
    args.pairMap = *redoBuf;  // the full-matrix kernels below run over the redo list only
    args.nWorkDev = ctx->dWork + SD_W_REDO;
  }
  if (chi) {  // chimera probes: a few thousand pairs per sample, int32 kernels with the probe compiled in
    ctx->lastKernelWasH2 = false;
    if (needC > 16) return launchWide<true>(ctx, args, maxLenA, maxLenB);
    return needC <= 8 ? launchI32<8, true>(ctx, args, maxLenA, maxLenB) : launchI32<16, true>(ctx, args, maxLenA, maxLenB);
  }
  if (needC > 16) {  // reads wider than 512 columns: strip-mined int32 pass
    if (ctx->forceKernel == 2) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "packed fp16 alignment kernel not applicable to this launch");
    ctx->lastKernelWasH2 = false;
    return launchWide(ctx, args, maxLenA, maxLenB);
  }
  const uint32_t C = needC <= 4 ? 4 : needC <= 6 ? 6 : needC <= 8 ? 8 : needC <= 10 ? 10 : needC <= 13 ? 13 : 16;
  // packed fp16 path: bases within A,C,G,T,N (codes 0..4), exact-integer range, enough pairs to pair up
  const bool useH2 = ctx->forceKernel != 1 && ctx->usedSym <= H2_SYMS && nPairs >= 2 && halfRangeOk(ctx, maxLenA + 1, 32 * C);
  if (ctx->forceKernel == 2 && !useH2) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "packed fp16 alignment kernel not applicable to this launch");
  ctx->lastKernelWasH2 = useH2;
  if (useH2) {
    switch (C) {
      case 4: return launchH2<4>(ctx, args, maxLenA, maxLenB);
      case 6: return launchH2<6>(ctx, args, maxLenA, maxLenB);
      case 8: return launchH2<8>(ctx, args, maxLenA, maxLenB);
      case 10: return launchH2<10>(ctx, args, maxLenA, maxLenB);
      case 13: return launchH2<13>(ctx, args, maxLenA, maxLenB);
      default: return launchH2<16>(ctx, args, maxLenA, maxLenB);
    }
  }
  switch (C) {
    case 4: return launchI32<4>(ctx, args, maxLenA, maxLenB);
    case 6: return launchI32<6>(ctx, args, maxLenA, maxLenB);
    case 8: return launchI32<8>(ctx, args, maxLenA, maxLenB);
    case 10: return launchI32<10>(ctx, args, maxLenA, maxLenB);
    case 13: return launchI32<13>(ctx, args, maxLenA, maxLenB);
    default: return launchI32<16>(ctx, args, maxLenA, maxLenB);
  }
}
