// align_kernel.cu — kernels 2+3 of the qluster hot path, fused:
//   forward affine-gap global DP (njhseq alignCalc::runNeedleSave semantics) with packed
//   traceback bits -> warp-cooperative traceback -> gapped strings -> errorProfile
//   (aligner::profileAlignment semantics) -> one sdgpu_comparison per pair.
//
// Replaces, per (ref, read) pair: aligner::alignCacheGlobal + aligner::profileAlignment
// (reference call sites clusterDown.cpp:694-696, 739-741; ControlBencher.hpp:117-118).
//
// Formulation (deliberately different from the reference's cell-of-three-scores matrix):
//   one alignment per warp; the read (B, DP columns) is cut into 32 strips of C columns, each
//   lane keeps its strip's state in registers; the ref (A, DP rows) streams through the lanes
//   as a skewed wavefront (lane l works on row t-l at step t).  Each cell produces its three
//   OUTGOING maxima
//       toDiag  = max(U, L, D)                -> D of the cell down-right  (+ score)
//       toDown  = max(U-ge, L-go, D-go)       -> U of the cell below
//       toRight = max(U-go, L-ge, D-go)       -> L of the cell to the right
//   with njhseq's needleMaximum tie order (U, then L, then D; its 'B' pointer behaves as 'U'
//   everywhere in the traceback), recorded as 5 comparison bits per cell.  Neighbouring lanes
//   exchange toRight/toDiag of their last column with two shuffles per row.  Bits go to a
//   per-warp slab in HBM/L2 in (step, word, lane) order so every store is a full 128-byte line.
//   End-gap penalties: last row via per-row registers, last column via per-column registers,
//   first row/column through the initial register state and lane 0's boundary feed.
#include <cuda_runtime.h>

#include <algorithm>

#include "sdgpu_internal.cuh"

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr unsigned FULL = 0xffffffffu;

struct AlignArgs {
  DevScoring sc;
  DevSeqs seqs;
  DevKmerIndex idx;
  const uint32_t* refIdx;
  const uint32_t* readIdx;
  uint32_t nPairs;
  uint32_t flags;
  sdgpu_comparison* out;
  sdgpu_gap* gapsOut;
  uint32_t gapCap;
  uint8_t* scratch;
  uint64_t slabBytes;  // per warp
  uint32_t ptrBytes;   // bytes of the pointer region in a slab
  uint32_t maxAln;     // capacity of each gapped string
  uint32_t* work;      // [0] = next pair, [2..3] = cell counter (uint64)
  const sdgpu_iter_par* lines;  // par-file lines for the verdict mask (device)
  uint32_t nLines;
  uint64_t* passOut;   // per pair: bit l = passes line l (or NULL)
};

enum { ST_D = 0, ST_U = 1, ST_L = 2 };

// Traceback bits.  Each cell contributes the SIGN BITS of five differences, pushed into a
// 32-bit accumulator with one funnel shift each (6 cells = 30 bits per word):
//   g4 = L < D | g3 = U < max(L,D) | g2 = U-ge < max(L,D)-go | g1 = L-ge < D-go | g0 = U-go < max(L-ge,D-go)
// cellBits() returns the complemented group, so bit set == ">=" (njhseq's tie order U, L, D).
constexpr int CELLS_PER_WORD = 6;
__device__ __forceinline__ int nextFromDiag(unsigned b) { return (b & 8u) ? ST_U : ((b & 16u) ? ST_L : ST_D); }
__device__ __forceinline__ int nextFromUp(unsigned b) { return (b & 4u) ? ST_U : ((b & 16u) ? ST_L : ST_D); }
__device__ __forceinline__ int nextFromLeft(unsigned b) { return (b & 1u) ? ST_U : ((b & 2u) ? ST_L : ST_D); }

template <int C>
__device__ __forceinline__ unsigned cellBits(const uint8_t* ptr, int i, int j) {
  constexpr int CW = (C + CELLS_PER_WORD - 1) / CELLS_PER_WORD;
  const int lj = (j - 1) / C, c = (j - 1) - lj * C;
  const int t = i + lj;
  const int wi = c / CELLS_PER_WORD, k = c - wi * CELLS_PER_WORD;
  const int inWord = (C - wi * CELLS_PER_WORD) < CELLS_PER_WORD ? (C - wi * CELLS_PER_WORD) : CELLS_PER_WORD;
  const unsigned w = __ldcg(reinterpret_cast<const unsigned*>(ptr) + (size_t(t - 1) * CW + wi) * 32 + lj);
  return ~(w >> (5 * (inWord - 1 - k))) & 31u;
}

// njhseq seqInfo::checkQual (strictly-greater thresholds, trailing window stops before the last base)
__device__ bool checkQualDev(const uint8_t* q, int len, int pos, const DevScoring& sc) {
  if (int(q[pos]) <= sc.primaryQual) return false;
  const int lower = pos > sc.window ? pos - sc.window : 0;
  for (int i = lower; i < pos; ++i)
    if (int(q[i]) <= sc.secondaryQual) return false;
  int higher = len - 1;
  if (pos + sc.window + 1 < len) higher = pos + sc.window + 1;
  for (int i = pos + 1; i < higher; ++i)
    if (int(q[i]) <= sc.secondaryQual) return false;
  return true;
}

// KmerMaps::isKmerLowFrequency for the k-mer centred on base p (clamped to the sequence)
__device__ bool lowFreqDev(const DevKmerIndex& idx, const uint8_t* codes, int len, int p) {
  const int k = int(idx.kLen);
  if (k == 0 || len < k) return false;
  int start = p >= k / 2 ? p - k / 2 : 0;
  if (start + k > len) start = len - k;
  uint64_t key = 0;
  for (int x = 0; x < k; ++x) key = (key << 4) | codes[start + x];
  if (idx.byPos) key |= uint64_t(start) << 48;
  uint32_t lo = 0, hi = idx.n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (idx.keys[mid] < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  const uint32_t cnt = (lo < idx.n && idx.keys[lo] == key) ? idx.counts[lo] : 0u;
  return cnt <= idx.runCutOff;
}

__device__ __forceinline__ unsigned warpSum(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

template <int C>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) alignProfileKernel(const AlignArgs a) {
  constexpr int CW = (C + CELLS_PER_WORD - 1) / CELLS_PER_WORD;
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  __syncthreads();
  const char* sTabBytes = reinterpret_cast<const char*>(sTab);

  const int lane = threadIdx.x & 31;
  const int warpInCta = threadIdx.x >> 5;
  const size_t warpGlobal = size_t(blockIdx.x) * WARPS_PER_CTA + warpInCta;
  uint8_t* slab = a.scratch + warpGlobal * a.slabBytes;
  uint32_t* ptrWords = reinterpret_cast<uint32_t*>(slab);
  const uint8_t* ptrBytes = slab;
  uint8_t* alnA = slab + a.ptrBytes;
  uint8_t* alnB = alnA + a.maxAln;
  uint32_t* gapRecs = reinterpret_cast<uint32_t*>(alnB + a.maxAln);  // (colAbs, size, inA) triples

  const DevScoring& sc = a.sc;
  const int dInt = sc.go - sc.ge;

  for (;;) {
    uint32_t pair = 0;
    if (lane == 0) pair = atomicAdd(&a.work[0], 1u);
    pair = __shfl_sync(FULL, pair, 0);
    if (pair >= a.nPairs) break;

    const uint32_t ia = a.refIdx[pair], ib = a.readIdx[pair];
    const uint64_t offA = a.seqs.offsets[ia], offB = a.seqs.offsets[ib];
    const int la = int(a.seqs.offsets[ia + 1] - offA), lb = int(a.seqs.offsets[ib + 1] - offB);
    const uint8_t* cA = a.seqs.codes + offA;
    const uint8_t* cB = a.seqs.codes + offB;
    const uint8_t* qA = a.seqs.quals + offA;
    const uint8_t* qB = a.seqs.quals + offB;

    // ------------------------------------------------------------------ forward DP
    int U[C], T[C], dcol[C], gocol[C], colofs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int j = lane * C + c + 1;
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
    const int nLanes = (lb + C - 1) / C;
    const int steps = la + nLanes - 1;
    int aCode = (lane == 0) ? int(cA[0]) : 0;  // base of the row this lane works on at t = 1
    for (int t = 1; t <= steps; ++t) {
      const int recvR = __shfl_up_sync(FULL, outR, 1);
      const int recvT = __shfl_up_sync(FULL, outT, 1);
      const int i = t - lane;
      // prefetch next row's base
      const int iNext = i + 1;
      const int aNext = (iNext >= 1 && iNext <= la) ? int(cA[iNext - 1]) : 0;
      if (lane < nLanes && i >= 1 && i <= la) {
        const bool lastRow = (i == la);
        const int gorow = lastRow ? sc.goRR : sc.go;
        const int drow = lastRow ? (sc.goRR - sc.geRR) : dInt;
        int Lin, Tin;
        if (lane == 0) {
          const int bnd = -sc.goLQ - (i - 1) * sc.geLQ;  // upInherit of column 0
          Lin = bnd - gorow;                             // njhseq's j == 1 case
          Tin = (i == 1) ? 0 : bnd + sc.geLQ;
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
          uint32_t acc = w[c / CELLS_PER_WORD];
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
          w[c / CELLS_PER_WORD] = acc;
        }
        outR = Lin;
        outT = T[C - 1];
        uint32_t* dst = ptrWords + size_t(t - 1) * (CW * 32) + lane;
#pragma unroll
        for (int x = 0; x < CW; ++x) dst[x * 32] = w[x];
      }
      prevRecvT = recvT;
      aCode = aNext;
    }
    // score = toDiag of the bottom-right cell; its arg-max is the traceback's first state
    const int lastLane = (lb - 1) / C, lastC = (lb - 1) - lastLane * C;
    int score = 0;
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (c == lastC) score = T[c];
    score = __shfl_sync(FULL, score, lastLane);
    __syncwarp();

    // ------------------------------------------------------------------ traceback
    int i = la, j = lb;
    int state = nextFromDiag(cellBits<C>(ptrBytes, la, lb));
    int wp = la + lb - 1;  // next free column (filled backwards)
    uint32_t gapA = 0, gapB = 0, nRecs = 0;
    while (i > 0 || j > 0) {
      const int di = (state != ST_L), dj = (state != ST_U);
      const int ik = i - lane * di, jk = j - lane * dj;
      bool valid;
      int ns = ST_D;
      if (state == ST_U) {
        valid = ik >= 1;
        if (valid) ns = (jk == 0) ? ST_U : (ik == 1 ? ST_L : nextFromUp(cellBits<C>(ptrBytes, ik - 1, jk)));
      } else if (state == ST_L) {
        valid = jk >= 1;
        if (valid) ns = (ik == 0) ? ST_L : (jk == 1 ? ST_U : nextFromLeft(cellBits<C>(ptrBytes, ik, jk - 1)));
      } else {
        valid = ik >= 1 && jk >= 1;
        if (valid) ns = (ik == 1) ? ST_L : (jk == 1 ? ST_U : nextFromDiag(cellBits<C>(ptrBytes, ik - 1, jk - 1)));
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
            gapRecs[3 * nRecs + 2] = 0u;
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
            gapRecs[3 * nRecs + 2] = 1u;
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
        gapRecs[3 * nRecs + 2] = inA ? 1u : 0u;
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
    int baseA = 0, baseB = 0;
    const unsigned ltMask = (1u << lane) - 1u;
    for (int cs = 0; cs < alnLen; cs += 32) {
      const int k = cs + lane;
      const bool inb = k < alnLen;
      const int ca = inb ? int(gA[k]) : SDGPU_GAP_CODE;
      const int cb = inb ? int(gB[k]) : SDGPU_GAP_CODE;
      const unsigned hasA = __ballot_sync(FULL, inb && ca != SDGPU_GAP_CODE);
      const unsigned hasB = __ballot_sync(FULL, inb && cb != SDGPU_GAP_CODE);
      if (inb && ca != SDGPU_GAP_CODE && cb != SDGPU_GAP_CODE) {
        const int pa = baseA + __popc(hasA & ltMask), pb = baseB + __popc(hasB & ltMask);
        if (sTab[ca * SDGPU_MAX_SYMBOLS + cb] < 0) {
          bool good = true;
          if (usingQuality) good = checkQualDev(qA, la, pa, sc) && checkQualDev(qB, lb, pb, sc);
          if (!good) {
            ++lq;
          } else if (checkKmer && (lowFreqDev(a.idx, cA, la, pa) || lowFreqDev(a.idx, cB, lb, pb))) {
            ++lk;
          } else {
            ++hq;
          }
        } else {
          if (matchQuality && !(checkQualDev(qA, la, pa, sc) && checkQualDev(qB, lb, pb, sc))) {
            ++lqM;
          } else {
            ++hqM;
          }
        }
      }
      baseA += __popc(hasA);
      baseB += __popc(hasB);
    }
    hq = warpSum(hq);
    lq = warpSum(lq);
    lk = warpSum(lk);
    hqM = warpSum(hqM);
    lqM = warpSum(lqM);

    // gap runs in forward order (the records are in traceback order); uniform across the warp
    double one = 0, two = 0, large = 0;
    unsigned numGaps = 0, endA = 0, endB = 0, regA = 0, regB = 0;
    for (int r = int(nRecs) - 1; r >= 0; --r) {
      const int startPos = int(gapRecs[3 * r + 0]) - alnStart;
      const int size = int(gapRecs[3 * r + 1]);
      const bool inA = gapRecs[3 * r + 2] != 0;
      const bool endGap = (startPos + size >= alnLen) || startPos == 0;
      if (endGap) {
        (inA ? endA : endB) += size;
        if (!sc.countEndGaps) continue;
      } else {
        (inA ? regA : regB) += size;
      }
      ++numGaps;
      double wgt = 1.0;
      if (sc.weighHomopolymers) {
        const uint8_t* gs = inA ? gA : gB;  // string holding the '-' run
        const uint8_t* os = inA ? gB : gA;  // string holding the gapped bases
        const int base = os[startPos];
        bool homo = true;
        for (int x = 1; x < size; ++x) homo = homo && (os[startPos + x] == base);
        if (homo) {
          double gapSide = 0, baseSide = size;
          for (int cur = startPos + size; cur < alnLen && gs[cur] == base; ++cur) gapSide += 1;
          for (int cur = 1; cur <= startPos && gs[startPos - cur] == base; ++cur) gapSide += 1;
          for (int cur = startPos + size; cur < alnLen && os[cur] == base; ++cur) baseSide += 1;
          for (int cur = 1; cur <= startPos && os[startPos - cur] == base; ++cur) baseSide += 1;
          wgt = fabs(gapSide - baseSide) / (gapSide + baseSide);
        }
      }
      if (size >= 3) {
        large += wgt;
      } else if (size == 2) {
        two += wgt;
      } else {
        one += wgt;
      }
    }

    // comparison::passErrorProfile against every par-file line at once: lane l judges lines l and
    // l + 32, two ballots give the 64-bit verdict mask the host collapse loop consumes
    if (a.passOut) {
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
        a.passOut[pair] = (uint64_t(halves[1]) << 32) | halves[0];
        if (!a.out) atomicAdd(reinterpret_cast<unsigned long long*>(a.work + 2), (unsigned long long)la * (unsigned long long)lb);
      }
    }
    if (lane == 0 && a.out) {
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
      atomicAdd(reinterpret_cast<unsigned long long*>(a.work + 2), (unsigned long long)la * (unsigned long long)lb);
    }
    __syncwarp();
  }
}

template <int C>
int launchC(sdgpu_ctx* ctx, AlignArgs& args, uint32_t maxLenA, uint32_t maxLenB) {
  constexpr int CW = (C + CELLS_PER_WORD - 1) / CELLS_PER_WORD;
  int ctasPerSM = 0;
  SD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, alignProfileKernel<C>, WARPS_PER_CTA * 32, 0));
  if (ctasPerSM < 1) ctasPerSM = 1;
  uint32_t grid = uint32_t(ctx->numSMs) * uint32_t(ctasPerSM);
  const uint32_t needCtas = (args.nPairs + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  if (grid > needCtas) grid = needCtas;
  const uint64_t maxSteps = uint64_t(maxLenA) + 31;
  args.ptrBytes = uint32_t(maxSteps * CW * 128);
  args.maxAln = (maxLenA + maxLenB + 15) & ~15u;
  uint64_t slab = uint64_t(args.ptrBytes) + 2ull * args.maxAln + 12ull * (uint64_t(args.maxAln) + 2);
  slab = (slab + 255) & ~255ull;
  args.slabBytes = slab;
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
    e0 = ctx->evPool[ctx->evUsed++];
    e1 = ctx->evPool[ctx->evUsed++];
    SD_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  }
  alignProfileKernel<C><<<grid, WARPS_PER_CTA * 32, 0, ctx->stream>>>(args);
  SD_CUDA(ctx, cudaGetLastError());
  if (e1) SD_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  ++ctx->launches;
  ++ctx->alignLaunches;
  return SDGPU_OK;
}

}  // namespace

int sd_launch_align(sdgpu_ctx* ctx, const uint32_t* dRefIdx, const uint32_t* dReadIdx, uint32_t nPairs,
                    uint32_t flags, sdgpu_comparison* dOut, sdgpu_gap* dGaps, uint32_t gapCap, uint32_t maxLenA,
                    uint32_t maxLenB, uint64_t* dPassOut) {
  if (nPairs == 0) return SDGPU_OK;
  if ((flags & SDGPU_CHECK_KMER) && !ctx->haveIndex)
    return sd_fail(ctx, SDGPU_E_STATE, "SDGPU_CHECK_KMER requires sdgpu_build_kmer_index first");
  AlignArgs args;
  args.sc = ctx->scoring;
  args.seqs = sd_dev_seqs(ctx);
  args.idx = sd_dev_index(ctx);
  args.refIdx = dRefIdx;
  args.readIdx = dReadIdx;
  args.nPairs = nPairs;
  args.flags = flags;
  args.out = dOut;
  args.gapsOut = dGaps;
  args.gapCap = gapCap;
  args.work = ctx->dWork;
  args.lines = ctx->dLines;
  args.nLines = ctx->nLines;
  args.passOut = dPassOut;
  if (dPassOut && ctx->nLines == 0) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_set_pass_table has not been called");
  const uint32_t needC = (maxLenB + 31) / 32;
  if (needC <= 4) return launchC<4>(ctx, args, maxLenA, maxLenB);
  if (needC <= 6) return launchC<6>(ctx, args, maxLenA, maxLenB);
  if (needC <= 8) return launchC<8>(ctx, args, maxLenA, maxLenB);
  if (needC <= 10) return launchC<10>(ctx, args, maxLenA, maxLenB);
  if (needC <= 13) return launchC<13>(ctx, args, maxLenA, maxLenB);
  if (needC <= 16) return launchC<16>(ctx, args, maxLenA, maxLenB);
  return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "read longer than 512 bases: column strip-mining not implemented yet");
}
