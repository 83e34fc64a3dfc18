// allpairs.cu — all-vs-all global alignment of a set of sequences (BASELINE config 3; the
// pair-block form of §8e: any contiguous block of the upper-triangle pair list can be computed
// on any GPU, the sequences are replicated, nothing is reduced).
//
// Reference analogue: the all-by-all comparison qluster can dump for its final clusters
// (clusterDown.cpp:727-805: alignCacheGlobal + profileAlignment for every pair, writing
// alnScore_, identities, 1/2/large indels, lq/hq/lowKmer mismatches) and the population-level
// comparisons of processClusters (processClusters.cpp:226).  Pair lists never exist on the host:
// a kernel unranks the pair index p -> (i, j), the alignment kernel runs on the device lists, and
// a second kernel packs each comparison into a 28-byte summary before it crosses PCIe.
#include <cuda_runtime.h>

#include <algorithm>

#include "sdgpu_internal.cuh"

namespace {

// p in [0, n(n-1)/2)  ->  (i, j), i < j, row-major over the upper triangle
__global__ void triPairKernel(const uint32_t* seqIdx, uint32_t n, uint64_t first, uint32_t count, uint32_t* ref,
                              uint32_t* read) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= count) return;
  const uint64_t p = first + x;
  // row i starts at S(i) = i*(2n - i - 1)/2; invert with a float guess, then correct
  const double nn = double(n);
  double guess = floor(((2.0 * nn - 1.0) - sqrt((2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * double(p))) * 0.5);
  uint64_t i = guess < 0 ? 0 : uint64_t(guess);
  if (i > uint64_t(n) - 2) i = uint64_t(n) - 2;
  auto start = [&](uint64_t r) { return r * (2ull * n - r - 1ull) / 2ull; };
  while (i > 0 && start(i) > p) --i;
  while (i + 2 < uint64_t(n) && start(i + 1) <= p) ++i;
  const uint64_t j = i + 1 + (p - start(i));
  ref[x] = seqIdx ? seqIdx[i] : uint32_t(i);
  read[x] = seqIdx ? seqIdx[j] : uint32_t(j);
}

__global__ void summaryKernel(const sdgpu_comparison* in, uint32_t count, sdgpu_pair_summary* out) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= count) return;
  const sdgpu_comparison c = in[x];
  sdgpu_pair_summary s;
  s.alnScore = c.alnScore;
  s.hqMismatches = uint16_t(min(c.hqMismatches, 65535u));
  s.lqMismatches = uint16_t(min(c.lqMismatches, 65535u));
  s.lowKmerMismatches = uint16_t(min(c.lowKmerMismatches, 65535u));
  s.numGaps = uint16_t(min(c.numGaps, 65535u));
  s.oneBaseIndel = float(c.oneBaseIndel);
  s.twoBaseIndel = float(c.twoBaseIndel);
  s.largeBaseIndel = float(c.largeBaseIndel);
  s.eventBasedIdentity = float(c.eventBasedIdentity);
  out[x] = s;
}

}  // namespace

extern "C" int sdgpu_all_vs_all(sdgpu_ctx* ctx, const uint32_t* seqIdx, uint32_t n, uint32_t flags, uint64_t firstPair,
                                uint64_t nPairs, sdgpu_pair_summary* out) {
  if (!ctx || (nPairs && !out)) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  if (!seqIdx) n = nSeqs;
  if (n < 2) return nPairs ? sd_fail(ctx, SDGPU_E_ARG, "need at least two sequences") : SDGPU_OK;
  const uint64_t total = uint64_t(n) * (n - 1) / 2;
  if (firstPair > total || nPairs > total - firstPair) return sd_fail(ctx, SDGPU_E_ARG, "pair block outside the triangle");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  uint32_t maxLen = 0;
  for (uint32_t x = 0; x < n; ++x) {
    const uint32_t s = seqIdx ? seqIdx[x] : x;
    if (s >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "sequence index out of range");
    maxLen = std::max<uint32_t>(maxLen, uint32_t(ctx->hOffsets[s + 1] - ctx->hOffsets[s]));
  }
  const uint32_t CH = 1u << 21;
  const uint64_t bytes = uint64_t(CH) * (sizeof(sdgpu_comparison) + sizeof(sdgpu_pair_summary) + 8) + uint64_t(n) * 4 + 256;
  int rc = sd_ensure_stage(ctx, bytes);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  sdgpu_comparison* dCmp = reinterpret_cast<sdgpu_comparison*>(base);
  sdgpu_pair_summary* dSum = reinterpret_cast<sdgpu_pair_summary*>(base + uint64_t(CH) * sizeof(sdgpu_comparison));
  uint32_t* dRef = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(dSum) + uint64_t(CH) * sizeof(sdgpu_pair_summary));
  uint32_t* dRead = dRef + CH;
  uint32_t* dIdx = nullptr;
  if (seqIdx) {
    dIdx = dRead + CH;
    SD_CUDA(ctx, cudaMemcpyAsync(dIdx, seqIdx, uint64_t(n) * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2dBytes += uint64_t(n) * 4;
  }
  for (uint64_t off = 0; off < nPairs; off += CH) {
    const uint32_t m = uint32_t(std::min<uint64_t>(CH, nPairs - off));
    triPairKernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(dIdx, n, firstPair + off, m, dRef, dRead);
    SD_CUDA(ctx, cudaGetLastError());
    ++ctx->launches;
    rc = sd_launch_align(ctx, dRef, dRead, m, flags, dCmp, nullptr, 0, maxLen, maxLen);
    if (rc) return rc;
    summaryKernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(dCmp, m, dSum);
    SD_CUDA(ctx, cudaGetLastError());
    ++ctx->launches;
    SD_CUDA(ctx, cudaMemcpyAsync(out + off, dSum, uint64_t(m) * sizeof(sdgpu_pair_summary), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->d2hBytes += uint64_t(m) * sizeof(sdgpu_pair_summary);
  }
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}
