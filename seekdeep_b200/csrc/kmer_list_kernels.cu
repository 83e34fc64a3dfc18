// kmer_list_kernels.cu — sorted k-mer lists (kmerInfo for any kLen <= 16) and the k-mer binning masks of
// qluster's --useKmerBinning (clusterDownSetUp.cpp:352-354: kmerBinOpts_.useKmerBinning_ / kmerCutOff_ / kCompareLen_).
//
// The dense profiles of kmer_kernels.cu stop at 16 384 bins (k <= 7 over A,C,G,T); the binning k is 10 by default.
// For those lengths kmerInfo is kept the way a sequence of a few hundred bases fills it: the sequence's own k-mers,
// sorted, with repeats.  Layout: one uint64 key per base slot of the sequence store (key i of sequence s at
// lists[offsets[s] + i], i < len - k + 1), key = 4-bit codes of the k bases, first base most significant.
//   compareKmers(a, b).first = sum over k-mers of min(count in a, count in b)
//                            = #{ i : the j-th copy of key a_i (j counted inside a's run) exists in b }
// so one binary search per key of the read against the seed's list, no hashing, no atomics.
#include "sdgpu_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// One warp sorts the k-mers of one sequence at a time in its slice of shared memory (bitonic network over the next power
// of two, a lane takes every 32nd compare-exchange of a stage): the warps of a CTA work on different sequences and never
// wait for each other -- a stage ends with a warp barrier, not a CTA barrier.
__global__ void __launch_bounds__(256) kmerListKernel(DevSeqs seqs, uint32_t kLen, uint32_t slotKeys, uint64_t* lists) {
  extern __shared__ uint64_t sKeysAll[];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  uint64_t* sKeys = sKeysAll + size_t(warp) * slotKeys;
  for (uint32_t s = blockIdx.x * warps + warp; s < seqs.n; s += gridDim.x * warps) {
    const uint64_t off = seqs.offsets[s];
    const uint32_t len = uint32_t(seqs.offsets[s + 1] - off);
    if (len < kLen) continue;  // (uniform over the warp)
    const uint32_t nk = len - kLen + 1;
    uint32_t n2 = 2;
    while (n2 < nk) n2 <<= 1;
    const uint8_t* codes = seqs.codes + off;
    for (uint32_t i = lane; i < n2; i += 32) {
      uint64_t key = ~0ull;  // padding: never below a real key, never written back
      if (i < nk) {
        key = 0;
        for (uint32_t x = 0; x < kLen; ++x) key = (key << 4) | codes[i + x];
      }
      sKeys[i] = key;
    }
    __syncwarp();
    for (uint32_t size = 2; size <= n2; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = lane; t < (n2 >> 1); t += 32) {
          const uint32_t lo = 2 * t - (t & (stride - 1));
          const uint32_t hi = lo + stride;
          const bool asc = (lo & size) == 0;
          const uint64_t a = sKeys[lo], b = sKeys[hi];
          if ((a > b) == asc) {
            sKeys[lo] = b;
            sKeys[hi] = a;
          }
        }
        __syncwarp();
      }
    }
    for (uint32_t i = lane; i < nk; i += 32) lists[off + i] = sKeys[i];
    __syncwarp();
  }
}

// reads x seeds: a tile of seed lists sits in shared memory; one warp per read, each lane looks its share of the read's
// keys up in every seed list.  Up to KPL keys per lane stay in registers together with their rank inside the read's run
// of equal keys; longer reads re-read their list (L1).  Output: bit c of maskOut[r] <=> compareKmers(seed c, read r).second
// > cutOff, the fraction being the exact double shared / (min(lenA, lenB) - k + 1).
constexpr int KPL = 10;
__global__ void __launch_bounds__(512) kmerBinKernel(DevSeqs seqs, const uint64_t* lists, uint32_t kLen, const uint32_t* seeds,
                                                     uint32_t nSeeds, uint32_t tileS, uint32_t seedStride, const uint32_t* reads,
                                                     uint32_t nReads, double cutOff, uint64_t* maskOut, uint32_t* sharedOut) {
  extern __shared__ uint64_t sSeed[];  // tileS x seedStride keys
  __shared__ uint32_t sNk[64], sLen[64];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (uint32_t s0 = 0; s0 < nSeeds; s0 += tileS) {
    const uint32_t nc = min(tileS, nSeeds - s0);
    __syncthreads();
    for (uint32_t c = 0; c < nc; ++c) {
      const uint32_t is = seeds[s0 + c];
      const uint64_t off = seqs.offsets[is];
      const uint32_t len = uint32_t(seqs.offsets[is + 1] - off);
      const uint32_t nk = len >= kLen ? len - kLen + 1 : 0;
      for (uint32_t i = threadIdx.x; i < nk; i += blockDim.x) sSeed[size_t(c) * seedStride + i] = __ldg(lists + off + i);
      if (threadIdx.x == 0) {
        sNk[c] = nk;
        sLen[c] = len;
      }
    }
    __syncthreads();
    for (uint32_t r = blockIdx.x * warps + warp; r < nReads; r += gridDim.x * warps) {
      const uint32_t ir = reads[r];
      const uint64_t offR = seqs.offsets[ir];
      const uint32_t lenR = uint32_t(seqs.offsets[ir + 1] - offR);
      const uint32_t nkR = lenR >= kLen ? lenR - kLen + 1 : 0;
      const uint64_t* R = lists + offR;
      const bool inRegs = nkR <= 32u * KPL;
      uint64_t xs[KPL];
      uint32_t js[KPL];
      if (inRegs) {
#pragma unroll
        for (int q = 0; q < KPL; ++q) {
          const uint32_t i = lane + 32u * q;
          xs[q] = 0;
          js[q] = 0;
          if (i < nkR) {
            const uint64_t x = __ldg(R + i);
            uint32_t j = 0;
            while (j < i && __ldg(R + i - 1 - j) == x) ++j;
            xs[q] = x;
            js[q] = j;
          }
        }
      }
      uint64_t mask = 0;
      for (uint32_t c = 0; c < nc; ++c) {
        const uint32_t nS = sNk[c];
        const uint64_t* S = sSeed + size_t(c) * seedStride;
        uint32_t cnt = 0;
        if (inRegs) {
#pragma unroll
          for (int q = 0; q < KPL; ++q) {
            if (lane + 32u * q < nkR) {
              const uint64_t x = xs[q];
              uint32_t lo = 0, hi = nS;
              while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (S[mid] < x) {
                  lo = mid + 1;
                } else {
                  hi = mid;
                }
              }
              const uint32_t at = lo + js[q];
              cnt += (at < nS && S[at] == x) ? 1u : 0u;
            }
          }
        } else {
          for (uint32_t i = lane; i < nkR; i += 32) {
            const uint64_t x = __ldg(R + i);
            uint32_t j = 0;
            while (j < i && __ldg(R + i - 1 - j) == x) ++j;
            uint32_t lo = 0, hi = nS;
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if (S[mid] < x) {
                lo = mid + 1;
              } else {
                hi = mid;
              }
            }
            const uint32_t at = lo + j;
            cnt += (at < nS && S[at] == x) ? 1u : 0u;
          }
        }
        cnt = __reduce_add_sync(FULL, cnt);
        if (lane == 0) {
          const uint32_t minLen = lenR < sLen[c] ? lenR : sLen[c];
          const double frac = (minLen >= kLen) ? double(cnt) / double(minLen - kLen + 1) : 0.0;
          if (SDGPU_KMER_BIN_PASSES(frac, cutOff)) mask |= 1ull << (s0 + c);
          if (sharedOut) sharedOut[size_t(r) * nSeeds + s0 + c] = cnt;
        }
      }
      if (lane == 0) maskOut[r] = (s0 == 0) ? mask : (maskOut[r] | mask);
    }
  }
}

}  // namespace

int sd_build_kmer_lists(sdgpu_ctx* ctx, uint32_t kLen) {
  if (kLen == 0 || kLen > 16) return sd_fail(ctx, SDGPU_E_ARG, "k-mer lists need 1 <= kLen <= 16");
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  ctx->listK = kLen;
  ctx->listSeqs = nSeqs;
  if (nSeqs == 0) return SDGPU_OK;
  if (ctx->maxLen > 16384) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "k-mer lists need sequences of at most 16384 bases");
  const uint64_t totalBases = ctx->hOffsets.back();
  if (totalBases > ctx->listCapKeys) {  // grow-only, like the profiles
    if (ctx->dKmerLists) cudaFree(ctx->dKmerLists);
    ctx->dKmerLists = nullptr;
    ctx->listCapKeys = 0;
    const uint64_t cap = totalBases + totalBases / 8 + 1024;  // the store grows by a few consensus versions per iteration
    SD_CUDA(ctx, cudaMalloc(&ctx->dKmerLists, cap * 8));
    ctx->listCapKeys = cap;
  }
  uint32_t n2 = 2;
  while (n2 < ctx->maxLen) n2 <<= 1;
  uint32_t warps = uint32_t((160u << 10) / (size_t(n2) * 8));  // one slot of n2 keys per warp
  warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
  const size_t smem = size_t(warps) * n2 * 8;
  SD_CUDA(ctx, cudaFuncSetAttribute(kmerListKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  uint32_t grid = uint32_t(ctx->numSMs) * 8;
  if (grid > (nSeqs + warps - 1) / warps) grid = (nSeqs + warps - 1) / warps;
  kmerListKernel<<<grid, warps * 32, smem, ctx->stream>>>(sd_dev_seqs(ctx), kLen, n2, ctx->dKmerLists);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

int sd_kmer_bin_masks(sdgpu_ctx* ctx, const uint32_t* seeds, uint32_t nSeeds, const uint32_t* reads, uint32_t nReads, double cutOff,
                      uint64_t* maskOut, uint32_t* sharedOut) {
  if (!ctx->dKmerLists || ctx->listK == 0) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_build_kmer_lists has not been called");
  if (nSeeds > 64) return sd_fail(ctx, SDGPU_E_ARG, "at most 64 seeds per call (one bit each)");
  if (nReads == 0) return SDGPU_OK;
  if (nSeeds == 0) {
    for (uint32_t r = 0; r < nReads; ++r) maskOut[r] = 0;
    return SDGPU_OK;
  }
  uint32_t stride = 1;
  for (uint32_t c = 0; c < nSeeds; ++c) {
    if (seeds[c] >= ctx->listSeqs) return sd_fail(ctx, SDGPU_E_ARG, "seed index outside the listed store");
    const uint32_t len = uint32_t(ctx->hOffsets[seeds[c] + 1] - ctx->hOffsets[seeds[c]]);
    if (len >= ctx->listK && len - ctx->listK + 1 > stride) stride = len - ctx->listK + 1;
  }
  for (uint32_t r = 0; r < nReads; ++r)
    if (reads[r] >= ctx->listSeqs) return sd_fail(ctx, SDGPU_E_ARG, "read index outside the listed store");
  uint32_t tileS = uint32_t((200u << 10) / (size_t(stride) * 8));
  if (tileS == 0) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "seed k-mer list larger than shared memory");
  if (tileS > nSeeds) tileS = nSeeds;
  const size_t smem = size_t(tileS) * stride * 8;
  const size_t sharedBytes = sharedOut ? size_t(nReads) * nSeeds * 4 : 0;
  const size_t bytes = size_t(nReads) * 8 + sharedBytes + (size_t(nReads) + nSeeds) * 4 + 64;
  int rc = sd_ensure_stage(ctx, bytes);
  if (rc) return rc;
  uint64_t* dMask = reinterpret_cast<uint64_t*>(ctx->dStage);
  uint32_t* dShared = reinterpret_cast<uint32_t*>(dMask + nReads);
  uint32_t* dReads = dShared + (sharedOut ? size_t(nReads) * nSeeds : 0);
  uint32_t* dSeeds = dReads + nReads;
  SD_CUDA(ctx, cudaMemcpyAsync(dReads, reads, size_t(nReads) * 4, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dSeeds, seeds, size_t(nSeeds) * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2dBytes += (size_t(nReads) + nSeeds) * 4;
  SD_CUDA(ctx, cudaFuncSetAttribute(kmerBinKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const uint32_t ctasPerSM = smem > (100u << 10) ? 1 : 2;
  const uint32_t threads = ctasPerSM == 1 ? 512 : 256;
  uint32_t grid = uint32_t(ctx->numSMs) * ctasPerSM;
  const uint32_t need = (nReads + threads / 32 - 1) / (threads / 32);
  if (grid > need) grid = need;
  kmerBinKernel<<<grid, threads, smem, ctx->stream>>>(sd_dev_seqs(ctx), ctx->dKmerLists, ctx->listK, dSeeds, nSeeds, tileS, stride, dReads,
                                                      nReads, cutOff, dMask, sharedOut ? dShared : nullptr);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  SD_CUDA(ctx, cudaMemcpyAsync(maskOut, dMask, size_t(nReads) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (sharedOut) SD_CUDA(ctx, cudaMemcpyAsync(sharedOut, dShared, sharedBytes, cudaMemcpyDeviceToHost, ctx->stream));
  ctx->d2hBytes += size_t(nReads) * 8 + sharedBytes;
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}
