// stitch_kernel.cu — aligner::alignRegGlobalNoInternalGaps + aligner::profileAlignment for read pairs
// (PairedReadProcessor.cpp:333-336): the alignment of two mates without internal gaps is a shift, so every
// diagonal of the pair is scored (substitution scores of the overlapping columns minus the end-gap penalties of the
// two overhangs) and the best one is profiled.  One warp per pair: the mates are staged in shared memory as ready
// byte offsets into the score table, lane l scores diagonals l, l + 32, ...; a warp arg-max (smallest shift among
// equals) picks the alignment; the errorProfile of the overlap is one more strided sweep using the store's per-base
// flag bytes (checkQual outcome, low-frequency k-mer).
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>

#include "../../include/sdgpu_stitch.h"
#include "sdgpu_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int ST_WARPS = 4;
constexpr int ST_MAXLEN = 1024;  // mates up to this many bases are staged in shared memory

struct StitchArgs {
  DevScoring sc;
  DevSeqs seqs;
  const uint32_t* r1Idx;
  const uint32_t* r2Idx;
  uint32_t nPairs;
  uint32_t flags;
  sdgpu_overlap* out;
  uint32_t* work;
};

// end-gap penalties of the alignment with shift d (r2's first base at r1 position d)
__device__ __forceinline__ int overhangCost(const DevScoring& sc, int la, int lb, int d) {
  int cost = 0;
  if (d > 0) cost += sc.goLQ + (d - 1) * sc.geLQ;     // r1 starts alone: '-' in the read (r2) at the left end
  if (d < 0) cost += sc.goLR + (-d - 1) * sc.geLR;    // r2 starts alone: '-' in the ref (r1) at the left end
  const int tail = la - (lb + d);
  if (tail > 0) cost += sc.goRQ + (tail - 1) * sc.geRQ;   // r1 ends alone
  if (tail < 0) cost += sc.goRR + (-tail - 1) * sc.geRR;  // r2 ends alone
  return cost;
}

__global__ void __launch_bounds__(ST_WARPS * 32) overlapKernel(const StitchArgs a) {
  __shared__ int sTab[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];
  __shared__ uint16_t sRow[ST_WARPS][ST_MAXLEN];
  __shared__ uint8_t sCol[ST_WARPS][ST_MAXLEN];
  for (int x = threadIdx.x; x < SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS; x += blockDim.x) sTab[x] = a.sc.score[x];
  __syncthreads();
  const char* tab = reinterpret_cast<const char*>(sTab);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool checkKmer = (a.flags & SDGPU_CHECK_KMER) != 0;
  const bool usingQuality = (a.flags & SDGPU_USING_QUALITY) != 0;
  const bool matchQuality = (a.flags & SDGPU_MATCH_QUALITY) != 0;
  for (;;) {
    uint32_t pair = 0;
    if (lane == 0) pair = atomicAdd(a.work, 1u);
    pair = __shfl_sync(FULL, pair, 0);
    if (pair >= a.nPairs) break;
    const uint32_t ia = a.r1Idx[pair], ib = a.r2Idx[pair];
    const uint64_t offA = a.seqs.offsets[ia], offB = a.seqs.offsets[ib];
    const int la = int(a.seqs.offsets[ia + 1] - offA), lb = int(a.seqs.offsets[ib + 1] - offB);
    const uint8_t *cA = a.seqs.codes + offA, *cB = a.seqs.codes + offB;
    __syncwarp();
    for (int x = lane; x < la; x += 32) sRow[w][x] = uint16_t(int(cA[x]) * (SDGPU_MAX_SYMBOLS * 4));
    for (int x = lane; x < lb; x += 32) sCol[w][x] = uint8_t(int(cB[x]) * 4);
    __syncwarp();
    // every diagonal: k = d + (lb - 1) in [0, la + lb - 2]
    int best = INT_MIN, bestK = 0;
    const int nDiag = la + lb - 1;
    for (int k = lane; k < nDiag; k += 32) {
      const int d = k - (lb - 1);
      const int i0 = d > 0 ? d : 0, i1 = min(la, lb + d);  // r1 positions of the overlap
      int s = 0;
      const uint16_t* rp = sRow[w] + i0;
      const uint8_t* cp = sCol[w] + (i0 - d);
      for (int x = 0; x < i1 - i0; ++x) s += *reinterpret_cast<const int*>(tab + rp[x] + cp[x]);
      s -= overhangCost(a.sc, la, lb, d);
      if (s > best) best = s, bestK = k;  // ascending k: the first of equals stays
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int os = __shfl_xor_sync(FULL, best, o), ok = __shfl_xor_sync(FULL, bestK, o);
      if (os > best || (os == best && ok < bestK)) best = os, bestK = ok;
    }
    const int d = bestK - (lb - 1);
    const int i0 = d > 0 ? d : 0, i1 = min(la, lb + d);
    // errorProfile of the overlap (profileAlignment: end gaps are not counted, there are no others)
    const uint8_t *fA = a.seqs.flags + offA, *fB = a.seqs.flags + offB;
    unsigned hq = 0, lq = 0, lk = 0, hqM = 0, lqM = 0;
    for (int i = i0 + lane; i < i1; i += 32) {
      const int j = i - d;
      const int s = *reinterpret_cast<const int*>(tab + sRow[w][i] + sCol[w][j]);
      const unsigned fa = fA[i], fb = fB[j];
      const bool good = fa & fb & SD_FLAG_QUAL_OK;
      if (s < 0) {
        if (usingQuality && !good) {
          ++lq;
        } else if (checkKmer && (SDGPU_LOWKMER_FLAGS(fa, fb) & SD_FLAG_LOW_KMER)) {
          ++lk;
        } else {
          ++hq;
        }
      } else if (matchQuality && !good) {
        ++lqM;
      } else {
        ++hqM;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      hq += __shfl_xor_sync(FULL, hq, o);
      lq += __shfl_xor_sync(FULL, lq, o);
      lk += __shfl_xor_sync(FULL, lk, o);
      hqM += __shfl_xor_sync(FULL, hqM, o);
      lqM += __shfl_xor_sync(FULL, lqM, o);
    }
    if (lane == 0) {
      sdgpu_overlap o;
      o.alnScore = best;
      o.shift = d;
      o.basesInAln = uint32_t(i1 - i0);
      o.hqMismatches = hq;
      o.lqMismatches = lq;
      o.lowKmerMismatches = lk;
      o.highQualityMatches = hqM;
      o.lowQualityMatches = lqM;
      const unsigned events = hqM + lqM + hq + lq + lk;
      o.eventBasedIdentity = events ? double(hqM + lqM) / double(events) : 0.0;
      o.eventBasedIdentityHq = events ? double(hqM + lqM + lq + lk) / double(events) : 0.0;
      a.out[pair] = o;
    }
  }
}

}  // namespace

extern "C" int sdgpu_align_no_internal_gaps(sdgpu_ctx* ctx, const uint32_t* r1Idx, const uint32_t* r2Idx, uint32_t nPairs, uint32_t flags,
                                            sdgpu_overlap* out) {
  if (!ctx || (nPairs && (!r1Idx || !r2Idx || !out))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->scoring.countEndGaps) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "alignment without internal gaps: countEndGaps is not supported");
  if ((flags & SDGPU_CHECK_KMER) && !ctx->haveIndex) return sd_fail(ctx, SDGPU_E_STATE, "SDGPU_CHECK_KMER requires sdgpu_build_kmer_index first");
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (r1Idx[x] >= nSeqs || r2Idx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    if (ctx->hOffsets[r1Idx[x] + 1] - ctx->hOffsets[r1Idx[x]] > uint64_t(ST_MAXLEN) ||
        ctx->hOffsets[r2Idx[x] + 1] - ctx->hOffsets[r2Idx[x]] > uint64_t(ST_MAXLEN))
      return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "alignment without internal gaps: reads longer than 1024 bases");
  }
  const uint64_t bytesIdx = uint64_t(nPairs) * 4, bytesOut = uint64_t(nPairs) * sizeof(sdgpu_overlap);
  int rc = sd_ensure_stage(ctx, bytesOut + 2 * bytesIdx + 256);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  StitchArgs args;
  args.sc = ctx->scoring;
  args.seqs = sd_dev_seqs(ctx);
  args.out = reinterpret_cast<sdgpu_overlap*>(base);
  uint32_t* dR1 = reinterpret_cast<uint32_t*>(base + bytesOut);
  args.r1Idx = dR1;
  args.r2Idx = dR1 + nPairs;
  args.nPairs = nPairs;
  args.flags = flags;
  args.work = ctx->dWork + SD_W_NEXT;
  SD_CUDA(ctx, cudaMemcpyAsync(dR1, r1Idx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dR1 + nPairs, r2Idx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemsetAsync(ctx->dWork + SD_W_NEXT, 0, sizeof(uint32_t), ctx->stream));
  const uint32_t grid = std::min<uint32_t>((nPairs + ST_WARPS - 1) / ST_WARPS, uint32_t(ctx->numSMs) * 8);
  overlapKernel<<<grid, ST_WARPS * 32, 0, ctx->stream>>>(args);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  SD_CUDA(ctx, cudaMemcpyAsync(out, args.out, bytesOut, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += bytesOut;
  return SDGPU_OK;
}
