// screen_kernel.cu — the screen pass in front of the verdict-mask launches (sdgpu_align_pass*,
// sdgpu_align_pass_grid): decide a pair without traceback whenever that is provably exact.
//
// Replaces nothing of the reference's interface by itself: it is the first stage of what
// aligner::alignCacheGlobal + aligner::profileAlignment + comparison::passErrorProfile compute per
// (cluster, read) pair in the collapse loop (clusterDown.cpp:694-697, 739-757; ControlBencher.hpp:117-120).
//
// Most pairs of a collapse are two reads of the same length that differ by substitutions only.  For
// such a pair the alignment the reference would trace is the gapless one -- PROVIDED no gapped
// alignment scores as much.  Two kernels establish that without recording a single traceback bit:
//
//   screenScanKernel   8 lanes per pair.  Scores the gapless alignment (S_diag = sum of mat_[a][b] down the
//                      main diagonal), classifies its mismatches exactly like profileAlignment (checkQual
//                      windows, low-frequency k-mers: per-base flag bytes of the store) and evaluates every par-file line -> the verdict mask
//                      the pair has IF its traced alignment is the gapless one.  From the deficit
//                      maxSub*L - S_diag it picks the narrowest band (16 / 32 / 64 diagonals) outside of
//                      which no path can reach S_diag, and files the pair into that band's list.  Pairs
//                      of unequal length (their alignment must contain a gap) and pairs too divergent for
//                      a 64-wide band go to the redo list.
//   screenDpKernel<G>  G = 2 / 4 / 8 lanes per pair, 8 diagonals per lane, 32 / G pairs per warp.  Score-only
//                      banded DP over anti-diagonals (no decision bits, no stores) under a PERTURBED
//                      scoring: every substitution score and penalty times K, every gap-open penalty
//                      minus 1.  Its optimum is S' = max over in-band alignments of K*score + b, where
//                      b = 0 for the gapless alignment and 1 <= b < K for any alignment with a gap
//                      (b counts gap opens; K exceeds the number of moves of any path).  Hence
//                      S' == K*S_diag  <=>  every gapped in-band alignment scores strictly less than S_diag.
//                      Together with the scan's band bound the gapless alignment is then the UNIQUE
//                      optimum of the full matrix, so njhseq's traceback yields it whatever its tie rules
//                      do, and the scan's mask is the pair's mask.  Otherwise the pair joins the redo list.
//
// The redo list is aligned by the bit-recording passes of align_kernel.cu exactly as before.
// Cell recurrence (H = best of the three states, E / F = best score entering the cell below / to the right):
//     H = max(E_in, F_in, H_diag + K*s);  E = max(E_in - ge', H - go');  F = max(F_in - ge', H - go')
// which equals the reference's three-state form whenever ge' <= go'; where a right-end pair is 0/0 the
// perturbed open (-1) is cheaper than the extend (0) and a gap run may be charged as several opens: b
// then counts up to one per gap move, still < K.  The launcher only uses the screen when ge <= go holds
// for all five penalty pairs.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>

#include "sdgpu_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SCAN_G = 8;               // lanes per pair in the scan
constexpr int SCR_NEG = -(1 << 30);     // "no path" (never accumulated: band edges are re-fed every step)
constexpr int TABW = 20;                // row stride (words) of the K-scaled score table: A,C,G,T rows hit disjoint banks
constexpr int SCR_MIN_LEN = 40;         // shorter pairs are left to the general kernels

// ------------------------------------------------------------------------------------------------
// scan: gapless score + errorProfile + verdict mask, band selection
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) screenScanKernel(const SdScreenArgs a) {
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  __shared__ sdgpu_iter_par sLines[SDGPU_MAX_PASS_LINES];
  __shared__ SdScreenItem sItem[4][256];  // the verdicts of a warp's grab, filed together
  __shared__ uint8_t sBinOf[4][256];
  const int wInCta = threadIdx.x >> 5;
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  for (int x = threadIdx.x; x < int(a.nLines); x += blockDim.x) sLines[x] = a.lines[x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gl = lane & (SCAN_G - 1), q = lane / SCAN_G;
  const unsigned gmask = 0xffu << (q * SCAN_G);
  constexpr uint32_t PPW = 32 / SCAN_G;
  const bool checkKmer = (a.flags & SDGPU_CHECK_KMER) != 0;
  const bool usingQuality = (a.flags & SDGPU_USING_QUALITY) != 0;
  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&a.work[SD_W_SCAN], 64u * PPW);  // 64 rounds of pairs (256) per grab
    base = __shfl_sync(FULL, base, 0);
    if (base >= a.nPairs) break;
    for (uint32_t round = 0; round < 64; ++round) {
      const uint32_t pair = base + round * PPW + q;
      // bin: 0..2 = band of 16 / 32 / 64 diagonals, 3 = redo (general kernels), 4 = nothing to do
      int bin = 4;
      uint64_t mask = 0;
      int sdiag = 0;
      int L = 0;
      if (pair < a.nPairs) {
        uint32_t ia, ib;
        sd_pair_of(a.src, pair, ia, ib);
        if (a.src.clus && ia == ib) {
          if (a.passOut && gl == 0) a.passOut[pair] = 0;  // a cluster is never compared with itself
        } else {
          const uint64_t offA = a.seqs.offsets[ia], offB = a.seqs.offsets[ib];
          const int la = int(a.seqs.offsets[ia + 1] - offA), lb = int(a.seqs.offsets[ib + 1] - offB);
          bin = 3;
          if (a.screenOn && la == lb && la >= SCR_MIN_LEN) {
            L = la;
            const uint8_t *cA = a.seqs.codes + offA, *cB = a.seqs.codes + offB;
            const uint8_t *fA = a.seqs.flags + offA, *fB = a.seqs.flags + offB;  // per-base SD_FLAG_* bytes
            unsigned hq = 0, lq = 0, lk = 0;
            // four columns at a time: the two sequences' codes as 32-bit words (two aligned loads + a funnel shift each:
            // sequences start at arbitrary bytes; the store ends in spare bytes).
            const unsigned shA = unsigned(reinterpret_cast<uintptr_t>(cA) & 3u) * 8u, shB = unsigned(reinterpret_cast<uintptr_t>(cB) & 3u) * 8u;
            const uint32_t* wA = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(cA) & ~uintptr_t(3));
            const uint32_t* wB = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(cB) & ~uintptr_t(3));
            const unsigned shFA = unsigned(reinterpret_cast<uintptr_t>(fA) & 3u) * 8u, shFB = unsigned(reinterpret_cast<uintptr_t>(fB) & 3u) * 8u;
            const uint32_t* wFA = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(fA) & ~uintptr_t(3));
            const uint32_t* wFB = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(fB) & ~uintptr_t(3));
            // (the kernel is bound by load latency, not by instructions: the words of five iterations are requested
            //  before the first one is looked at)
            constexpr int MLP = 5;
            for (int w0 = gl; 4 * w0 < L; w0 += SCAN_G * MLP) {
              uint32_t va[MLP], vb[MLP];
#pragma unroll
              for (int t = 0; t < MLP; ++t) {
                const int w = w0 + SCAN_G * t;
                const bool in = 4 * w < L;
                va[t] = in ? __funnelshift_r(__ldg(wA + w), __ldg(wA + w + 1), shA) : 0u;
                vb[t] = in ? __funnelshift_r(__ldg(wB + w), __ldg(wB + w + 1), shB) : 0u;
              }
#pragma unroll
              for (int t = 0; t < MLP; ++t) {
                const int w = w0 + SCAN_G * t;
                const int p0 = 4 * w, nv = min(4, L - p0);
                if (nv <= 0) continue;
                if (a.selfScore >= 0 && a.misScore < 0) {
                  // match / mismatch matrix: no table, no branches.  Codes are < 16, so adding 0x7f to every byte of
                  // the XOR raises bit 7 exactly in the bytes that differ; flag bytes are folded the same way.
                  const uint32_t valid = nv == 4 ? 0x80808080u : (0x80808080u & ((1u << (8 * nv)) - 1u));
                  const uint32_t mis = (((va[t] ^ vb[t]) + 0x7f7f7f7fu) & valid);
                  sdiag += nv * a.selfScore - __popc(mis) * (a.selfScore - a.misScore);
                  if (mis) {
                    const uint32_t fa = __funnelshift_r(__ldg(wFA + w), __ldg(wFA + w + 1), shFA);
                    const uint32_t fb = __funnelshift_r(__ldg(wFB + w), __ldg(wFB + w + 1), shFB);
                    const uint32_t good = usingQuality ? (((fa & fb) << 7) & 0x80808080u) : 0x80808080u;       // SD_FLAG_QUAL_OK in both
                    const uint32_t low = checkKmer ? (((SDGPU_LOWKMER_FLAGS(fa, fb) >> 1) << 7) & 0x80808080u) : 0u;             // SD_FLAG_LOW_KMER in either
                    lq += __popc(mis & ~good);
                    lk += __popc(mis & good & low);
                    hq += __popc(mis & good & ~low);
                  }
                  continue;
                }
                for (int x = 0; x < nv; ++x) {
                  const int ca = (va[t] >> (8 * x)) & 255, cb = (vb[t] >> (8 * x)) & 255, p = p0 + x;
                  const int sc = sTab[ca * SDGPU_MAX_SYMBOLS + cb];
                  sdiag += sc;
                  if (sc < 0) {  // aligner::profileAlignment: a column whose score is negative is a mismatch
                    const unsigned fa = fA[p], fb = fB[p];
                    if (usingQuality && !(fa & fb & SD_FLAG_QUAL_OK)) {
                      ++lq;
                    } else if (checkKmer && (SDGPU_LOWKMER_FLAGS(fa, fb) & SD_FLAG_LOW_KMER)) {
                      ++lk;
                    } else {
                      ++hq;
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int o = SCAN_G / 2; o > 0; o >>= 1) {
              sdiag += __shfl_xor_sync(gmask, sdiag, o);
              hq += __shfl_xor_sync(gmask, hq, o);
              lq += __shfl_xor_sync(gmask, lq, o);
              lk += __shfl_xor_sync(gmask, lk, o);
            }
            const long long deficit = (long long)a.maxSub * L - sdiag;
            bin = deficit < a.bandSlack[0] ? 0 : (deficit < a.bandSlack[1] ? 1 : (deficit < a.bandSlack[2] ? 2 : 3));
            if (bin < 3) {
              // comparison::passErrorProfile of the gapless alignment's errorProfile (no gap runs: the three
              // indel tallies are 0, overLappingEvents = L), lane gl judges lines gl, gl + 8, ...
              const unsigned matches = unsigned(L) - hq - lq - lk;
              const double id = double(matches) / double(L);
              const double idHq = double(matches + lq + lk) / double(L);
              unsigned lo32 = 0, hi32 = 0;
              for (unsigned li = unsigned(gl); li < a.nLines; li += SCAN_G) {
                const sdgpu_iter_par& it = sLines[li];
                bool pass;
                if (it.onPerId) {
                  pass = (it.onHqPerId ? idHq : id) >= it.perId;
                } else {
                  pass = it.oneBaseIndel >= 0.0 && it.twoBaseIndel >= 0.0 && it.largeBaseIndel >= 0.0 &&
                         it.hqMismatches >= double(hq) && it.lqMismatches >= double(lq) && it.lowKmerMismatches >= double(lk);
                }
                if (pass) (li < 32 ? lo32 : hi32) |= 1u << (li & 31);
              }
#pragma unroll
              for (int o = SCAN_G / 2; o > 0; o >>= 1) {
                lo32 |= __shfl_xor_sync(gmask, lo32, o);
                hi32 |= __shfl_xor_sync(gmask, hi32, o);
              }
              mask = (uint64_t(hi32) << 32) | lo32;
            }
          }
        }
      }
      // park the verdict: the warp files all 256 pairs of its grab at once (below)
      if (gl == 0) {
        const uint32_t slot = round * PPW + uint32_t(q);
        sBinOf[wInCta][slot] = uint8_t(bin);
        sItem[wInCta][slot] = SdScreenItem{pair, sdiag * a.K, mask};
      }
    }
    // file the grab: one atomic per list and warp for 256 pairs (the lists' tails are hot addresses for every SM)
    __syncwarp();
    const unsigned ltMask = (1u << lane) - 1u;
    uint32_t cnt[4] = {0, 0, 0, 0}, first[4] = {0, 0, 0, 0}, running[4] = {0, 0, 0, 0};
    for (int c = 0; c < 8; ++c) {
      const int bn = sBinOf[wInCta][c * 32 + lane];
#pragma unroll
      for (int x = 0; x < 4; ++x) cnt[x] += __popc(__ballot_sync(FULL, bn == x));
    }
    const uint32_t myCnt = lane == 0 ? cnt[0] : (lane == 1 ? cnt[1] : (lane == 2 ? cnt[2] : cnt[3]));
    uint32_t myFirst = 0;
    if (lane < 4 && myCnt) myFirst = atomicAdd(&a.work[SD_W_BIN0 + lane], myCnt);  // lane x reserves list x's slots
#pragma unroll
    for (int x = 0; x < 4; ++x) first[x] = __shfl_sync(FULL, myFirst, x);
    for (int c = 0; c < 8; ++c) {
      const int bn = sBinOf[wInCta][c * 32 + lane];
      const SdScreenItem it = sItem[wInCta][c * 32 + lane];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const unsigned votes = __ballot_sync(FULL, bn == x);
        if (bn == x) {
          const uint32_t slot = first[x] + running[x] + __popc(votes & ltMask);
          if (x < 3) {
            a.items[size_t(x) * a.nPairs + slot] = it;
          } else {
            a.redo[slot] = it.pair;
          }
        }
        running[x] += __popc(votes);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// score-only banded DP under the perturbed scoring
// ------------------------------------------------------------------------------------------------
struct ScrPen {  // K-scaled penalties, opens cheaper by 1
  int go, nge;      // interior open, minus interior extend
  int goRQ, ngeRQ;  // vertical moves in the last column
  int goRR, ngeRR;  // horizontal moves in the last row
};

// one interior cell: no border, no last row / column
// (pinning the two additions to the FMA pipe as multiply-adds by a run-time 1 was measured: the ALU pipe drops from
//  305 to 239 instructions per 64 cells, but register moves grow the loop by 8 % and the kernel is 5 % slower)
__device__ __forceinline__ void scrCell(int& H, int& E, int& F, int Ein, int Fin, int s, int go, int nge) {
  const int Hn = __vimax3_s32(Ein, Fin, H + s);
  const int t = Hn - go;
  E = __viaddmax_s32(Ein, nge, t);
  F = __viaddmax_s32(Fin, nge, t);
  H = Hn;
}
// any cell: outside the matrix nothing changes; last row / column use the right-end penalties
__device__ __forceinline__ void scrCellGen(int& H, int& E, int& F, int Ein, int Fin, int s, int i, int j, int L, const ScrPen& P) {
  const bool valid = unsigned(i - 1) < unsigned(L) && unsigned(j - 1) < unsigned(L);
  const int goV = j == L ? P.goRQ : P.go, ngeV = j == L ? P.ngeRQ : P.nge;
  const int goH = i == L ? P.goRR : P.go, ngeH = i == L ? P.ngeRR : P.nge;
  const int Hn = __vimax3_s32(Ein, Fin, H + s);
  const int En = __viaddmax_s32(Ein, ngeV, Hn - goV);
  const int Fn = __viaddmax_s32(Fin, ngeH, Hn - goH);
  if (valid) {
    H = Hn;
    E = En;
    F = Fn;
  }
}

// Lane state: offsets d0 .. d0 + 7 (d0 = 2h, h = -2G + 4 * groupLane), per offset the outputs of its latest cell.
// Step pair p covers anti-diagonals 2p (even offsets k = 2m: cell i = p - h - m, j = p + h + m) and 2p + 1 (odd
// offsets k = 2m + 1: same row, column j + 1).  The bases a lane needs slide by one per step pair: ref bases
// i = p-h-3 .. p-h and read bases j = p+h .. p+h+4 live in two 8-slot register windows (slot = (p - m) & 7 for
// the ref, (p + m) & 7 for the read: compile-time under the 8-fold unrolling), each refilled four step pairs
// before it is needed, as ready-made byte offsets into the K-scaled score table (row offset + column offset).
template <int G, bool GEN>
__device__ __forceinline__ void scrBlock(int (&H)[8], int (&E)[8], int (&F)[8], uint32_t (&Aw)[8], uint32_t (&Bw)[8], int p0, int h,
                                         int gl, int L, const uint8_t* cA, const uint8_t* cB, const char* tab, const ScrPen& P) {
  // refills of an interior block: ref bases cA[p0 + 3 - h + u], read bases cB[p0 + h + 7 + u] (0-based), u = 0 .. 7
  const uint8_t* pa = cA + (p0 + 3 - h);
  const uint8_t* pb = cB + (p0 + h + 7);
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int p = p0 + u;
    // ---- even anti-diagonal: offsets 0, 2, 4, 6; the left input of offset 0 is the neighbour lane's F[7]
    int Fprev = __shfl_up_sync(FULL, F[7], 1);
    if (gl == 0) Fprev = SCR_NEG;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 2 * m;
      const int s = *reinterpret_cast<const int*>(tab + (Aw[(u - m) & 7] + Bw[(u + m) & 7]));
      const int Fin = m == 0 ? Fprev : F[k - 1];
      if (GEN) {
        scrCellGen(H[k], E[k], F[k], E[k + 1], Fin, s, p - h - m, p + h + m, L, P);
      } else {
        scrCell(H[k], E[k], F[k], E[k + 1], Fin, s, P.go, P.nge);
      }
    }
    // the read window's slot u is free now: base j = p + h + 8, first used four step pairs from here
    if (GEN) {
      const int j = min(max(p + h + 8, 1), L);
      Bw[u & 7] = uint32_t(__ldg(cB + j - 1)) * 4u;
    } else {
      Bw[u & 7] = uint32_t(__ldg(pb + u)) * 4u;
    }
    // ---- odd anti-diagonal: offsets 1, 3, 5, 7; the up input of offset 7 is the neighbour lane's E[0]
    int Enext = __shfl_down_sync(FULL, E[0], 1);
    if (gl == G - 1) Enext = SCR_NEG;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int k = 2 * m + 1;
      const int s = *reinterpret_cast<const int*>(tab + (Aw[(u - m) & 7] + Bw[(u + m + 1) & 7]));
      const int Ein = m == 3 ? Enext : E[k + 1];
      if (GEN) {
        scrCellGen(H[k], E[k], F[k], Ein, F[k - 1], s, p - h - m, p + h + m + 1, L, P);
      } else {
        scrCell(H[k], E[k], F[k], Ein, F[k - 1], s, P.go, P.nge);
      }
    }
    // the ref window's slot (u + 4) & 7 was last read at step pair p - 1: base i = p + 4 - h
    if (GEN) {
      const int i = min(max(p + 4 - h, 1), L);
      Aw[(u + 4) & 7] = uint32_t(__ldg(cA + i - 1)) * uint32_t(TABW * 4);
    } else {
      Aw[(u + 4) & 7] = uint32_t(__ldg(pa + u)) * uint32_t(TABW * 4);
    }
  }
}

template <int G>
__global__ void __launch_bounds__(128, 5) screenDpKernel(const SdScreenArgs a, const int bin) {
  constexpr int W = 8 * G;          // diagonals of the band: offsets -W/2 .. W/2 - 1
  constexpr uint32_t PPW = 32 / G;  // pairs per warp
  __shared__ int sTabK[SDGPU_MAX_SYMBOLS * TABW];
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * TABW; x += blockDim.x) {
    const int r = x / TABW, c = x - r * TABW;
    sTabK[x] = c < SDGPU_MAX_SYMBOLS ? a.K * int(a.sc.score[r * SDGPU_MAX_SYMBOLS + c]) : 0;
  }
  __syncthreads();
  const char* tab = reinterpret_cast<const char*>(sTabK);
  const int lane = threadIdx.x & 31;
  const int gl = lane % G, q = lane / G;
  const int K = a.K;
  ScrPen P;
  P.go = K * a.sc.go - 1, P.nge = -K * a.sc.ge;
  P.goRQ = K * a.sc.goRQ - 1, P.ngeRQ = -K * a.sc.geRQ;
  P.goRR = K * a.sc.goRR - 1, P.ngeRR = -K * a.sc.geRR;
  // plain registers from here on (otherwise the products are folded into every use and the fused add-max is lost)
  asm volatile("" : "+r"(P.go), "+r"(P.nge), "+r"(P.goRQ), "+r"(P.ngeRQ), "+r"(P.goRR), "+r"(P.ngeRR));
  const int goLQ = K * a.sc.goLQ - 1, geLQ = K * a.sc.geLQ, goLR = K * a.sc.goLR - 1, geLR = K * a.sc.geLR;
  const SdScreenItem* items = a.items + size_t(bin) * a.nPairs;
  const uint32_t nItems = a.work[SD_W_BIN0 + bin];
  const int h = -2 * G + 4 * gl;  // d0 / 2
  const int d0 = 2 * h;

  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(&a.work[SD_W_DP0 + bin], PPW);
    base = __shfl_sync(FULL, base, 0);
    if (base >= nItems) break;
    const bool have = base + q < nItems;
    const SdScreenItem it = items[have ? base + q : base];  // idle groups shadow the first pair; their result is dropped
    uint32_t ia, ib;
    sd_pair_of(a.src, it.pair, ia, ib);
    const uint64_t offA = a.seqs.offsets[ia], offB = a.seqs.offsets[ib];
    const int L = int(a.seqs.offsets[ia + 1] - offA);
    const uint8_t *cA = a.seqs.codes + offA, *cB = a.seqs.codes + offB;
    int Lmin = L, Lmax = L;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Lmin = min(Lmin, __shfl_xor_sync(FULL, Lmin, o));
      Lmax = max(Lmax, __shfl_xor_sync(FULL, Lmax, o));
    }
    // every offset starts from its border cell ((0, d) on row 0, (-d, 0) on column 0, (0, 0)): H = the border
    // score, E / F = the best score entering the cell below / to the right of it
    int H[8], E[8], F[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int d = d0 + k;
      if (d > L || -d > L) {
        H[k] = E[k] = F[k] = SCR_NEG;
      } else if (d > 0) {
        H[k] = -(goLR + (d - 1) * geLR);
        F[k] = H[k] - geLR;
        E[k] = H[k] - (d == L ? P.goRQ : P.go);
      } else if (d < 0) {
        H[k] = -(goLQ + (-d - 1) * geLQ);
        E[k] = H[k] - geLQ;
        F[k] = H[k] - (-d == L ? P.goRR : P.go);
      } else {
        H[k] = 0;
        E[k] = -goLQ;
        F[k] = -goLR;
      }
    }
    // windows for step pairs 0 .. 3: ref bases i = p - h - m (slot (p - m) & 7), read bases j = p + h + m (slot (p + m) & 7)
    uint32_t Aw[8], Bw[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      // ref slot x' = (x - 3) & 7 holds i = (x - 3) - h for x - 3 in -3 .. 3 (slot 4 is filled by step pair 0)
      const int i = min(max(x - 3 - h, 1), L);
      Aw[(x - 3) & 7] = uint32_t(__ldg(cA + i - 1)) * uint32_t(TABW * 4);
      const int j = min(max(x + h, 1), L);
      Bw[x] = uint32_t(__ldg(cB + j - 1)) * 4u;
    }
    // step pairs 0 .. Lmax in blocks of 8; a block is "interior" when every cell of every lane of every group
    // has 1 <= i, j <= L - 1: W/4 + 1 <= p and p + W/4 <= Lmin - 1 for all its step pairs
    const int nBlocks = Lmax / 8 + 1;
    const int bBody0 = (W / 4 + 1 + 7) / 8;
    const int lastBodyP = Lmin - 1 - W / 4;  // largest interior step pair
    const int bBody1 = lastBodyP >= 7 ? (lastBodyP - 7) / 8 + 1 : 0;  // blocks [bBody0, bBody1) are interior
    for (int b = 0; b < nBlocks; ++b) {
      if (b >= bBody0 && b < bBody1) {
        scrBlock<G, false>(H, E, F, Aw, Bw, 8 * b, h, gl, L, cA, cB, tab, P);
      } else {
        scrBlock<G, true>(H, E, F, Aw, Bw, 8 * b, h, gl, L, cA, cB, tab, P);
      }
    }
    // H of the corner cell (L, L): offset 0 = offset k = 0 of group lane G / 2
    const int corner = __shfl_sync(FULL, H[0], q * G + G / 2);
    const bool certified = have && corner == it.sdiagK;
    if (gl == 0 && have) {
      if (certified) {
        if (a.passOut) a.passOut[it.pair] = it.mask;
        if (a.hits.recs && it.mask) sd_emit_hit(a.src, a.hits, it.pair, it.mask);
      } else {
        a.redo[atomicAdd(&a.work[SD_W_BIN0 + 3], 1u)] = it.pair;
      }
    }
    // bookkeeping: band cells inside the matrix, pairs and lenA x lenB finished here
    unsigned long long cells = 0, certCells = 0;
    unsigned nCert = 0;
    if (gl == 0 && have) {
      long long c = 0;
      for (int d = -W / 2; d < W / 2; ++d) c += max(0, L - (d < 0 ? -d : d));
      cells = (unsigned long long)c;
      if (certified) certCells = (unsigned long long)L * (unsigned long long)L, nCert = 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cells += __shfl_xor_sync(FULL, cells, o);
      certCells += __shfl_xor_sync(FULL, certCells, o);
      nCert += __shfl_xor_sync(FULL, nCert, o);
    }
    if (lane == 0) {
      atomicAdd(reinterpret_cast<unsigned long long*>(a.work + SD_W_SCR_CELLS), cells);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.work + SD_W_SCR_CERTCELLS), certCells);
      atomicAdd(reinterpret_cast<unsigned long long*>(a.work + SD_W_CELLS), certCells);  // the launch's finished lenA x lenB
      atomicAdd(&a.work[SD_W_SCR_CERT], nCert);
    }
  }
}

// cheapest cost of n gap moves of one direction made away from the right end: one of them opens a run
long long gapMovesLowerBound(int n, int openA, int extA, int openB, int extB) {
  const int mAll = std::min(std::min(openA, extA), std::min(openB, extB));
  const int mOpen = std::min(openA, openB);
  return (long long)(n - 1) * mAll + mOpen;
}

}  // namespace

// Can the screen be used for this launch?  Fills K and the three band slacks.
bool sd_screen_plan(const sdgpu_ctx* ctx, uint32_t maxLenA, uint32_t maxLenB, SdScreenArgs* out) {
  const DevScoring& s = ctx->scoring;
  if (s.ge > s.go || s.geLQ > s.goLQ || s.geRQ > s.goRQ || s.geLR > s.goLR || s.geRR > s.goRR) return false;
  int maxSub = INT_MIN, maxAbs = 0;
  for (int x = 0; x < ctx->nSym; ++x)
    for (int y = 0; y < ctx->nSym; ++y) {
      const int v = s.score[x * SDGPU_MAX_SYMBOLS + y];
      maxSub = std::max(maxSub, v);
      maxAbs = std::max(maxAbs, std::abs(v));
    }
  if (maxSub <= 0) return false;
  const int pens[10] = {s.go, s.ge, s.goLQ, s.geLQ, s.goRQ, s.geRQ, s.goLR, s.geLR, s.goRR, s.geRR};
  int maxPen = 0;
  for (int v : pens) maxPen = std::max(maxPen, v);
  // K: a power of two above the number of moves of any path; every perturbed DP value must stay below 2^29
  const uint64_t moves = uint64_t(maxLenA) + maxLenB + 2;
  int K = 64;
  while (uint64_t(K) <= moves) K <<= 1;
  const uint64_t magnitude = uint64_t(maxAbs) * std::max(maxLenA, maxLenB) + uint64_t(maxPen + 1) * moves;
  if (K > (1 << 16) || magnitude * uint64_t(K) >= (1ull << 29)) return false;
  out->K = K;
  out->maxSub = maxSub;
  // the score every stored symbol has against itself, if it is the same non-negative number for all of them (else -1)
  out->selfScore = s.score[0];
  for (int x = 0; x < ctx->nSym; ++x)
    if (s.score[x * SDGPU_MAX_SYMBOLS + x] != out->selfScore) out->selfScore = -1;
  if (out->selfScore < 0) out->selfScore = -1;
  // ... and the score of every pair of different symbols, if it is the same negative number for all of them (else 0)
  out->misScore = ctx->nSym > 1 ? s.score[1] : 0;
  for (int x = 0; x < ctx->nSym; ++x)
    for (int y = 0; y < ctx->nSym; ++y)
      if (x != y && s.score[x * SDGPU_MAX_SYMBOLS + y] != out->misScore) out->misScore = 0;
  if (out->misScore >= 0) out->misScore = 0;
  // band of w diagonals, offsets -w/2 .. w/2 - 1, equal lengths L: a path touching offset w/2 makes >= w/2 right
  // moves above the last row and has at most L - w/2 diagonal moves; a path touching offset -w/2 - 1 makes
  // >= w/2 + 1 down moves left of the last column.  With deficit = maxSub*L - S_diag, S_diag beats both bounds iff
  // deficit < min(maxSub*nR + cost(nR right moves), maxSub*nD + cost(nD down moves)).
  const int widths[3] = {16, 32, 64};
  for (int b = 0; b < 3; ++b) {
    const int nR = widths[b] / 2, nD = widths[b] / 2 + 1;
    const long long hiSide = (long long)maxSub * nR + gapMovesLowerBound(nR, s.go, s.ge, s.goLR, s.geLR);
    const long long loSide = (long long)maxSub * nD + gapMovesLowerBound(nD, s.go, s.ge, s.goLQ, s.geLQ);
    out->bandSlack[b] = std::min(hiSide, loSide);
  }
  return true;
}

int sd_launch_screen(sdgpu_ctx* ctx, SdScreenArgs& args) {
  const int grid = ctx->numSMs * 8;
  auto timed = [&](int kind, auto launch) -> int {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->timeKernels) {
      if (ctx->evUsed + 2 > ctx->evPool.size()) {
        const size_t old = ctx->evPool.size();
        ctx->evPool.resize(old + 256, nullptr);
        for (size_t x = old; x < ctx->evPool.size(); ++x) SD_CUDA(ctx, cudaEventCreate(&ctx->evPool[x]));
      }
      if (ctx->evBand.size() < ctx->evPool.size() / 2) ctx->evBand.resize(ctx->evPool.size() / 2, 0);
      ctx->evBand[ctx->evUsed / 2] = uint8_t(kind);
      e0 = ctx->evPool[ctx->evUsed++];
      e1 = ctx->evPool[ctx->evUsed++];
      SD_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    }
    launch();
    SD_CUDA(ctx, cudaGetLastError());
    if (e1) SD_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    ++ctx->launches;
    ++ctx->alignLaunches;
    return SDGPU_OK;
  };
  int rc = timed(2, [&] { screenScanKernel<<<grid, 128, 0, ctx->stream>>>(args); });
  if (rc || !args.screenOn) return rc;
  if ((rc = timed(3, [&] { screenDpKernel<2><<<grid, 128, 0, ctx->stream>>>(args, 0); }))) return rc;
  if ((rc = timed(3, [&] { screenDpKernel<4><<<grid, 128, 0, ctx->stream>>>(args, 1); }))) return rc;
  return timed(3, [&] { screenDpKernel<8><<<grid, 128, 0, ctx->stream>>>(args, 2); });
}
