// kmer_kernels.cu — kernel 1 of the qluster hot path.
//   (1a) KmerMaps: positional k-mer frequency table over the current clusters.  Replaces
//        indexKmers(clusters, kLen, runCutOff, kmersByPosition, expandKmerPos, expandKmerSize)
//        (reference call sites clusterDown.cpp:412-417, 505-510).  The reference hashes
//        std::string k-mers per position; here every (position, k-mer) occurrence becomes a
//        64-bit key (pos << 48 | 4-bit-coded k-mer), keys are radix-sorted and run-length
//        reduced, and lookups (from the errorProfile kernel) are binary searches.
//   (1b) kmerInfo profiles + compareKmers.  Replaces kmerInfo(seq, k, false) and
//        a.compareKmers(b) (extractorPairedEnd.cpp:819-824; PrimersAndMids.cpp:49-54).
//        Dense sigma^k uint16 count vectors, streamed 128 bits at a time; shared k-mers are
//        SIMD min + add over packed 16-bit lanes.  HBM-bound by construction.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdlib>

#include "sdgpu_internal.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

// keys a k-mer at `cursor` emits before it when positions are expanded by +-e
__host__ __device__ inline uint64_t expandedBefore(uint64_t cursor, uint64_t e) {
  if (cursor <= e) return cursor * (e + 1) + cursor * (cursor - 1) / 2;
  return e * (e + 1) + e * (e - 1) / 2 + (cursor - e) * (2 * e + 1);
}

__global__ void kmerKeysKernel(DevSeqs seqs, const uint32_t* seqIdx, const uint64_t* keyStart, uint32_t nIdx,
                               uint32_t kLen, uint32_t byPos, uint32_t expand, uint64_t* keys, uint32_t* weights) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= nIdx) return;
  const uint32_t s = seqIdx ? seqIdx[warp] : warp;
  const uint64_t off = seqs.offsets[s];
  const uint32_t len = uint32_t(seqs.offsets[s + 1] - off);
  if (len < kLen) return;
  const uint8_t* c = seqs.codes + off;
  const uint32_t w = uint32_t(llrint(seqs.cnts[s]));
  const uint64_t base = keyStart[warp];
  const uint32_t nk = len - kLen + 1;
  for (uint32_t cursor = lane; cursor < nk; cursor += 32) {
    uint64_t kmer = 0;
    for (uint32_t x = 0; x < kLen; ++x) kmer = (kmer << 4) | c[cursor + x];
    if (!byPos) {
      keys[base + cursor] = kmer;
      weights[base + cursor] = w;
    } else {
      const uint32_t lo = cursor > expand ? cursor - expand : 0, hi = cursor + expand;
      uint64_t o = base + expandedBefore(cursor, expand);
      for (uint32_t p = lo; p <= hi; ++p, ++o) {
        keys[o] = (uint64_t(p) << 48) | kmer;
        weights[o] = w;
      }
    }
  }
}

// first key of every position of a by-position table: posStart[p] = lower_bound(keys, p << 48), p = 0 .. nPos
__global__ void kmerPosStartKernel(const uint64_t* keys, uint32_t n, uint32_t nPos, uint32_t* posStart) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > nPos) return;
  const uint64_t key = uint64_t(p) << 48;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  posStart[p] = lo;
}

// SD_FLAG_LOW_KMER of every base of sequences [firstSeq, firstSeq + nSeqs): one warp per sequence.  This is
// the look-up aligner::profileAlignment makes per would-be high-quality mismatch (kMaps_.isKmerLowFrequency on
// the k-mer centred on the base), made once per base and index build instead of once per mismatch and pair.
__global__ void __launch_bounds__(256) lowKmerFlagKernel(DevSeqs seqs, DevKmerIndex idx, uint32_t firstSeq, uint32_t nSeqs,
                                                        uint8_t* flags) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= nSeqs) return;
  const uint32_t s = firstSeq + warp;
  const uint64_t off = seqs.offsets[s];
  const int len = int(seqs.offsets[s + 1] - off);
  const uint8_t* c = seqs.codes + off;
  for (int p = int(lane); p < len; p += 32) {
    const bool low = sd_low_freq(idx, c, len, p);
    flags[off + p] = uint8_t((flags[off + p] & SD_FLAG_QUAL_OK) | (low ? SD_FLAG_LOW_KMER : 0u));
  }
}

// SD_FLAG_QUAL_OK of every base: njhseq seqInfo::checkQual (strictly-greater thresholds, trailing window stops
// before the last base); clears SD_FLAG_LOW_KMER (set again by lowKmerFlagKernel when an index exists)
__global__ void __launch_bounds__(256) qualFlagKernel(DevSeqs seqs, uint32_t firstSeq, uint32_t nSeqs, int primaryQual,
                                                     int secondaryQual, int window, uint8_t* flags) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= nSeqs) return;
  const uint32_t s = firstSeq + warp;
  const uint64_t off = seqs.offsets[s];
  const int len = int(seqs.offsets[s + 1] - off);
  const uint8_t* q = seqs.quals + off;
  for (int pos = int(lane); pos < len; pos += 32) {
    bool ok = SDGPU_QUAL_ABOVE(int(q[pos]), primaryQual);
    const int lower = pos > window ? pos - window : 0;
    for (int i = lower; ok && i < pos; ++i) ok = SDGPU_QUAL_ABOVE(int(q[i]), secondaryQual);
    const int higher = SDGPU_QUAL_WINDOW_END(pos, window, len);
    for (int i = pos + 1; ok && i < higher; ++i) ok = SDGPU_QUAL_ABOVE(int(q[i]), secondaryQual);
    flags[off + pos] = ok ? uint8_t(SD_FLAG_QUAL_OK) : uint8_t(0);
  }
}

__global__ void kmerLookupKernel(DevKmerIndex idx, const uint64_t* keys, uint32_t n, uint32_t* out) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const uint64_t key = keys[x];
  uint32_t lo = 0, hi = idx.n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (idx.keys[mid] < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  out[x] = (lo < idx.n && idx.keys[lo] == key) ? idx.counts[lo] : 0u;
}

// ---------------------------------------------------------------- dense profiles
// One warp per sequence.  Each lane owns a contiguous run of k-mer start positions and rolls the
// k-mer id (one new base per position); counts accumulate in shared memory with 32-bit atomics
// on packed uint16 pairs; the finished profile leaves as full 128-bit lines.  HBM-bound:
// L code bytes in, 2 * sigma^k profile bytes out per sequence.
__global__ void __launch_bounds__(256) kmerProfileKernel(DevSeqs seqs, uint32_t kLen, uint32_t sigma, uint32_t top,
                                                         uint32_t profSize, uint16_t* profiles) {
  extern __shared__ uint4 sProfVec[];  // warpsPerCta * profSize/8 vectors
  const uint32_t warpInCta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warpsPerCta = blockDim.x >> 5;
  const uint32_t vecs = profSize / 8;
  uint4* mineVec = sProfVec + warpInCta * vecs;
  uint32_t* mine = reinterpret_cast<uint32_t*>(mineVec);
  for (uint32_t s = blockIdx.x * warpsPerCta + warpInCta; s < seqs.n; s += gridDim.x * warpsPerCta) {
    for (uint32_t x = lane; x < vecs; x += 32) mineVec[x] = make_uint4(0, 0, 0, 0);
    const uint64_t off = seqs.offsets[s];
    const uint32_t len = uint32_t(seqs.offsets[s + 1] - off);
    const uint8_t* c = seqs.codes + off;
    __syncwarp();
    if (len >= kLen) {
      const uint32_t nk = len - kLen + 1;
      const uint32_t chunk = (nk + 31) / 32;
      const uint32_t p0 = lane * chunk;
      const uint32_t p1 = min(nk, p0 + chunk);
      if (p0 < p1) {
        uint32_t id = 0;
        for (uint32_t x = 0; x + 1 < kLen; ++x) id = id * sigma + c[p0 + x];
        uint32_t lead = 0;  // base leaving the window, times sigma^(k-1)
        for (uint32_t p = p0; p < p1; ++p) {
          id = (id - lead) * sigma + c[p + kLen - 1];
          lead = uint32_t(c[p]) * top;
          atomicAdd(&mine[id >> 1], 1u << (16 * (id & 1)));
        }
      }
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(profiles + size_t(s) * profSize);
    for (uint32_t x = lane; x < vecs; x += 32) dst[x] = mineVec[x];
    __syncwarp();
  }
}

// ACGT-only fast path of the profile build (every code < 4): a lane takes 4*W consecutive start
// positions, fetches the 4*W + kLen - 1 codes they cover as W + 3 aligned words, realigns them with
// funnel shifts and squeezes each word's four codes into one byte (first base most significant),
// so every k-mer id is a 2*kLen-bit field of one 64-bit register: per sequence the LSU sees
// W + 3 word loads per lane instead of two byte loads per position.
template <int W>
__device__ __forceinline__ void profileLane4(const uint32_t* codeWords, uint64_t off, uint32_t base, uint32_t len,
                                             uint32_t nk, uint32_t kLen, uint32_t mask, uint32_t lane, uint32_t* mine) {
  const uint32_t p0 = base + lane * 4 * W;
  if (p0 >= nk) return;
  const uint64_t b = off + p0;
  const uint32_t mis = uint32_t(b & 3);
  const uint64_t firstWord = b >> 2, endWord = (off + len + 3) >> 2;  // words holding this sequence's codes
  uint32_t w[W + 3];
#pragma unroll
  for (int j = 0; j < W + 3; ++j) w[j] = (firstWord + j < endWord) ? __ldg(codeWords + firstWord + j) : 0u;
  uint64_t bits = 0;
#pragma unroll
  for (int j = 0; j < W + 2; ++j) {
    const uint32_t r = __funnelshift_r(w[j], w[j + 1], 8 * mis) & 0x03030303u;
    bits = (bits << 8) | ((r * 0x40100401u) >> 24);  // c0<<6 | c1<<4 | c2<<2 | c3
  }
  constexpr int T = 4 * (W + 2);  // codes held in `bits`; code i sits at bit 2*(T-1-i)
#pragma unroll
  for (int i = 0; i < 4 * W; ++i) {
    if (p0 + i < nk) {
      const uint32_t id = uint32_t(bits >> (2 * (T - i - int(kLen)))) & mask;
      atomicAdd(&mine[id >> 1], 1u << (16 * (id & 1)));
    }
  }
}

__global__ void __launch_bounds__(256, 6) kmerProfileKernel4(DevSeqs seqs, uint32_t kLen, uint32_t profSize,
                                                             uint16_t* profiles) {
  extern __shared__ uint4 sProfVec[];
  const uint32_t warpInCta = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t warpsPerCta = blockDim.x >> 5;
  const uint32_t vecs = profSize / 8;
  uint4* mineVec = sProfVec + warpInCta * vecs;
  uint32_t* mine = reinterpret_cast<uint32_t*>(mineVec);
  const uint32_t* codeWords = reinterpret_cast<const uint32_t*>(seqs.codes);  // cudaMalloc'd: aligned
  const uint32_t mask = (1u << (2 * kLen)) - 1u;
  for (uint32_t x = lane; x < vecs; x += 32) mineVec[x] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  for (uint32_t s = blockIdx.x * warpsPerCta + warpInCta; s < seqs.n; s += gridDim.x * warpsPerCta) {
    const uint64_t off = seqs.offsets[s];
    const uint32_t len = uint32_t(seqs.offsets[s + 1] - off);
    if (len >= kLen) {
      const uint32_t nk = len - kLen + 1;
      for (uint32_t base = 0; base < nk; base += 512) {
        const uint32_t rem = nk - base;
        if (rem <= 128) {
          profileLane4<1>(codeWords, off, base, len, nk, kLen, mask, lane, mine);
        } else if (rem <= 256) {
          profileLane4<2>(codeWords, off, base, len, nk, kLen, mask, lane, mine);
        } else if (rem <= 384) {
          profileLane4<3>(codeWords, off, base, len, nk, kLen, mask, lane, mine);
        } else {
          profileLane4<4>(codeWords, off, base, len, nk, kLen, mask, lane, mine);
        }
      }
    }
    __syncwarp();
    uint4* dst = reinterpret_cast<uint4*>(profiles + size_t(s) * profSize);
    for (uint32_t x = lane; x < vecs; x += 32) {  // write out and leave the histogram zeroed for the next sequence
      dst[x] = mineVec[x];
      mineVec[x] = make_uint4(0, 0, 0, 0);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ uint32_t minSum8(const uint4& p, const uint4& q, uint32_t acc) {
  acc = __vadd2(acc, __vminu2(p.x, q.x));
  acc = __vadd2(acc, __vminu2(p.y, q.y));
  acc = __vadd2(acc, __vminu2(p.z, q.z));
  acc = __vadd2(acc, __vminu2(p.w, q.w));
  return acc;
}

// One warp per pair.  pairs given explicitly (a[], b[]) or as a dense reads x clusters grid.
// kmerInfo(seq, k, true): the reverse-complement map of every sequence = its forward profile permuted by
// k-mer -> reverse-complement k-mer.  One warp per sequence.
__global__ void kmerProfileRevKernel(const uint16_t* fwd, const uint32_t* rcIndex, uint32_t size, uint32_t profSize, uint32_t nSeqs,
                                     uint16_t* rev) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= nSeqs) return;
  const uint16_t* f = fwd + size_t(warp) * profSize;
  uint16_t* r = rev + size_t(warp) * profSize;
  for (uint32_t x = lane; x < size; x += 32) r[rcIndex[x]] = f[x];
  for (uint32_t x = size + lane; x < profSize; x += 32) r[x] = 0;
}

__global__ void kmerCompareKernel(DevSeqs seqs, const uint16_t* profiles, const uint16_t* profilesB, uint32_t profSize, uint32_t kLen,
                                  const uint32_t* a, const uint32_t* b, uint64_t nPairs, uint32_t nCols,
                                  uint32_t* sharedOut, double* fracOut) {
  const uint64_t warp0 = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nWarps = (uint64_t(gridDim.x) * blockDim.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t vecs = profSize / 8;
  for (uint64_t p = warp0; p < nPairs; p += nWarps) {
    uint32_t ia, ib;
    if (nCols) {  // dense grid: row = read, col = cluster
      ia = a[p / nCols];
      ib = b[p % nCols];
    } else {
      ia = a[p];
      ib = b[p];
    }
    const uint4* pa = reinterpret_cast<const uint4*>(profiles + size_t(ia) * profSize);
    const uint4* pb = reinterpret_cast<const uint4*>(profilesB + size_t(ib) * profSize);
    uint32_t acc = 0;
    for (uint32_t x = lane; x < vecs; x += 32) acc = minSum8(__ldg(pa + x), __ldg(pb + x), acc);
    uint32_t sum = (acc & 0xffffu) + (acc >> 16);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(FULL, sum, o);
    if (lane == 0) {
      const uint32_t la = uint32_t(seqs.offsets[ia + 1] - seqs.offsets[ia]);
      const uint32_t lb = uint32_t(seqs.offsets[ib + 1] - seqs.offsets[ib]);
      const uint32_t minLen = la < lb ? la : lb;
      sharedOut[p] = sum;
      fracOut[p] = (minLen >= kLen) ? double(sum) / double(minLen - kLen + 1) : 0.0;
    }
  }
}

}  // namespace

int sd_refresh_qual_flags(sdgpu_ctx* ctx, uint32_t firstSeq, uint32_t nSeqs) {
  if (nSeqs == 0) return SDGPU_OK;
  const DevScoring& s = ctx->scoring;
  qualFlagKernel<<<(nSeqs + 7) / 8, 256, 0, ctx->stream>>>(sd_dev_seqs(ctx), firstSeq, nSeqs, s.primaryQual, s.secondaryQual, s.window,
                                                         ctx->dFlags);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

int sd_refresh_kmer_flags(sdgpu_ctx* ctx, uint32_t firstSeq, uint32_t nSeqs) {
  if (nSeqs == 0 || !ctx->haveIndex) return SDGPU_OK;
  lowKmerFlagKernel<<<(nSeqs + 7) / 8, 256, 0, ctx->stream>>>(sd_dev_seqs(ctx), sd_dev_index(ctx), firstSeq, nSeqs, ctx->dFlags);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

int sd_build_kmer_index(sdgpu_ctx* ctx, const uint32_t* hSeqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                        int byPos, int expandPos, uint32_t expandSize) {
  if (kLen == 0 || kLen > 12) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "k-mer index supports 1 <= kLen <= 12");
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  if (!hSeqIdx) nIdx = nSeqs;
  const uint32_t e = (byPos && expandPos) ? expandSize : 0;
  std::vector<uint64_t> start(size_t(nIdx) + 1, 0);
  for (uint32_t n = 0; n < nIdx; ++n) {
    const uint32_t s = hSeqIdx ? hSeqIdx[n] : n;
    if (s >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "sequence index out of range");
    const uint64_t len = ctx->hOffsets[s + 1] - ctx->hOffsets[s];
    if (len >= 65536) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "k-mer index supports sequences < 65536 bases");
    const uint64_t nk = len >= kLen ? len - kLen + 1 : 0;
    start[n + 1] = start[n] + (byPos ? expandedBefore(nk, e) : nk);
  }
  const uint64_t total = start[nIdx];
  // (the sort / reduce below take a 32-bit signed item count)
  if (total >= (1ull << 31)) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "k-mer index too large (>= 2^31 occurrences)");
  ctx->nIdx = 0;
  ctx->idxNPos = 0;
  ctx->idxK = kLen;
  ctx->idxRunCutOff = runCutOff;
  ctx->idxByPos = byPos ? 1 : 0;
  ctx->haveIndex = true;
  if (total == 0) return sd_refresh_kmer_flags(ctx, 0, nSeqs);

  // one grow-only work arena per ctx: cudaMalloc/cudaFree of GB-sized buffers cost (variable)
  // hundreds of milliseconds and the index is rebuilt in every qluster run
  uint64_t *dKeysIn = nullptr, *dKeysOut = nullptr, *dStart = nullptr;
  uint32_t *dWIn = nullptr, *dWOut = nullptr, *dSeqIdx = nullptr, *dNumRuns = nullptr;
  void* dTemp = nullptr;
#define SD_CUDA_C(call) SD_CUDA(ctx, call)
  size_t tempSort = 0, tempReduce = 0;
  const int endBit = byPos ? 64 : int(4 * kLen);
  SD_CUDA_C(cub::DeviceRadixSort::SortPairs(nullptr, tempSort, dKeysIn, dKeysOut, dWIn, dWOut, int(total), 0, endBit,
                                            ctx->stream));
  SD_CUDA_C(cub::DeviceReduce::ReduceByKey(nullptr, tempReduce, dKeysOut, dKeysIn, dWOut, dWIn, dNumRuns, cub::Sum(),
                                           int(total), ctx->stream));
  auto up = [](uint64_t v) { return (v + 255) & ~255ull; };
  const uint64_t bKeys = up(total * 8), bW = up(total * 4), bStart = up((uint64_t(nIdx) + 1) * 8), bIdx = up(uint64_t(nIdx) * 4);
  const uint64_t bTemp = up(std::max(tempSort, tempReduce));
  const uint64_t need = 2 * bKeys + 2 * bW + bStart + bIdx + 256 + bTemp;
  if (need > ctx->idxWorkBytes) {
    if (ctx->dIdxWork) cudaFree(ctx->dIdxWork);
    ctx->dIdxWork = nullptr;
    ctx->idxWorkBytes = 0;
    SD_CUDA_C(cudaMalloc(&ctx->dIdxWork, need));
    ctx->idxWorkBytes = need;
  }
  {
    uint8_t* p = reinterpret_cast<uint8_t*>(ctx->dIdxWork);
    dKeysIn = reinterpret_cast<uint64_t*>(p); p += bKeys;
    dKeysOut = reinterpret_cast<uint64_t*>(p); p += bKeys;
    dWIn = reinterpret_cast<uint32_t*>(p); p += bW;
    dWOut = reinterpret_cast<uint32_t*>(p); p += bW;
    dStart = reinterpret_cast<uint64_t*>(p); p += bStart;
    dSeqIdx = hSeqIdx ? reinterpret_cast<uint32_t*>(p) : nullptr; p += bIdx;
    dNumRuns = reinterpret_cast<uint32_t*>(p); p += 256;
    dTemp = p;
  }
  SD_CUDA_C(cudaMemcpyAsync(dStart, start.data(), (size_t(nIdx) + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (hSeqIdx) SD_CUDA_C(cudaMemcpyAsync(dSeqIdx, hSeqIdx, size_t(nIdx) * 4, cudaMemcpyHostToDevice, ctx->stream));
  {
    const uint32_t threads = 128, warpsPerCta = threads / 32;
    const uint32_t grid = (nIdx + warpsPerCta - 1) / warpsPerCta;
    kmerKeysKernel<<<grid, threads, 0, ctx->stream>>>(sd_dev_seqs(ctx), dSeqIdx, dStart, nIdx, kLen, byPos ? 1 : 0, e,
                                                      dKeysIn, dWIn);
    SD_CUDA_C(cudaGetLastError());
    ++ctx->launches;
  }
  SD_CUDA_C(cub::DeviceRadixSort::SortPairs(dTemp, tempSort, dKeysIn, dKeysOut, dWIn, dWOut, int(total), 0, endBit,
                                            ctx->stream));
  SD_CUDA_C(cub::DeviceReduce::ReduceByKey(dTemp, tempReduce, dKeysOut, dKeysIn, dWOut, dWIn, dNumRuns, cub::Sum(),
                                           int(total), ctx->stream));
  ctx->launches += 2;
  uint32_t numRuns = 0;
  SD_CUDA_C(cudaMemcpyAsync(&numRuns, dNumRuns, 4, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA_C(cudaStreamSynchronize(ctx->stream));
  if (numRuns > ctx->idxCapRuns) {
    if (ctx->dIdxKeys) cudaFree(ctx->dIdxKeys);
    if (ctx->dIdxCnts) cudaFree(ctx->dIdxCnts);
    ctx->dIdxKeys = nullptr;
    ctx->dIdxCnts = nullptr;
    ctx->idxCapRuns = 0;
    SD_CUDA_C(cudaMalloc(&ctx->dIdxKeys, size_t(numRuns) * 8));
    SD_CUDA_C(cudaMalloc(&ctx->dIdxCnts, size_t(numRuns) * 4));
    ctx->idxCapRuns = numRuns;
  }
  SD_CUDA_C(cudaMemcpyAsync(ctx->dIdxKeys, dKeysIn, size_t(numRuns) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  SD_CUDA_C(cudaMemcpyAsync(ctx->dIdxCnts, dWIn, size_t(numRuns) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  SD_CUDA_C(cudaStreamSynchronize(ctx->stream));
  ctx->nIdx = numRuns;
  if (byPos) {  // first key of every position: look-ups search one position's keys only
    const uint32_t nPos = ctx->maxLen + e + 1;
    if (nPos + 1 > ctx->idxPosCap) {
      if (ctx->dIdxPosStart) cudaFree(ctx->dIdxPosStart);
      ctx->dIdxPosStart = nullptr;
      ctx->idxPosCap = 0;
      SD_CUDA_C(cudaMalloc(&ctx->dIdxPosStart, (size_t(nPos) + 1) * 4));
      ctx->idxPosCap = nPos + 1;
    }
    kmerPosStartKernel<<<(nPos + 256) / 256, 256, 0, ctx->stream>>>(ctx->dIdxKeys, numRuns, nPos, ctx->dIdxPosStart);
    SD_CUDA_C(cudaGetLastError());
    ++ctx->launches;
    ctx->idxNPos = nPos;
  }
#undef SD_CUDA_C
  // the low-frequency flag of every base of the store against the new table
  return sd_refresh_kmer_flags(ctx, 0, nSeqs);
}

int sd_kmer_index_lookup(sdgpu_ctx* ctx, const uint8_t* kmers, const uint32_t* positions, uint32_t n,
                         uint32_t* countsOut) {
  if (!ctx->haveIndex) return sd_fail(ctx, SDGPU_E_STATE, "no k-mer index built");
  if (n == 0) return SDGPU_OK;
  std::vector<uint64_t> keys(n);
  for (uint32_t x = 0; x < n; ++x) {
    uint64_t kmer = 0;
    bool known = true;
    for (uint32_t y = 0; y < ctx->idxK; ++y) {
      const int code = ctx->codeOf[kmers[size_t(x) * ctx->idxK + y]];
      if (code < 0) known = false;
      kmer = (kmer << 4) | uint64_t(code & 15);
    }
    keys[x] = known ? ((ctx->idxByPos ? (uint64_t(positions ? positions[x] : 0) << 48) : 0ull) | kmer) : ~0ull;
  }
  int rc = sd_ensure_stage(ctx, size_t(n) * 12);
  if (rc) return rc;
  uint64_t* dKeys = reinterpret_cast<uint64_t*>(ctx->dStage);
  uint32_t* dOut = reinterpret_cast<uint32_t*>(dKeys + n);
  SD_CUDA(ctx, cudaMemcpyAsync(dKeys, keys.data(), size_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
  kmerLookupKernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sd_dev_index(ctx), dKeys, n, dOut);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  SD_CUDA(ctx, cudaMemcpyAsync(countsOut, dOut, size_t(n) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sd_build_kmer_profiles(sdgpu_ctx* ctx, uint32_t kLen) {
  const uint32_t sigma = uint32_t(ctx->usedSym < 4 ? 4 : ctx->usedSym);
  uint64_t size = 1;
  for (uint32_t x = 0; x < kLen; ++x) {
    size *= sigma;
    if (size > 16384) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "dense k-mer profile larger than 16384 bins (kLen too large for this alphabet)");
  }
  if (kLen == 0) return sd_fail(ctx, SDGPU_E_ARG, "kLen must be >= 1");
  const uint32_t profSize = uint32_t((size + 7) & ~7ull);  // multiple of 8 uint16 = 16 bytes
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  if (ctx->maxLen >= 65536) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "profiles need sequences < 65536 bases");
  ctx->profK = kLen;
  ctx->profSize = profSize;
  ctx->profSeqs = nSeqs;
  ctx->profRevValid = false;
  if (nSeqs == 0) return SDGPU_OK;
  const uint64_t need = uint64_t(nSeqs) * profSize * 2;
  if (need > ctx->profCapBytes) {  // keep the buffer across rebuilds: cudaMalloc/cudaFree stall the stream
    if (ctx->dProfiles) cudaFree(ctx->dProfiles);
    ctx->dProfiles = nullptr;
    ctx->profCapBytes = 0;
    SD_CUDA(ctx, cudaMalloc(&ctx->dProfiles, need));
    ctx->profCapBytes = need;
  }
  uint32_t top = 1;
  for (uint32_t x = 0; x + 1 < kLen; ++x) top *= sigma;
  uint32_t warpsPerCta = uint32_t((200u << 10) / (size_t(profSize) * 2));  // stay under the 227 KB shared-memory limit
  warpsPerCta = warpsPerCta < 1 ? 1 : (warpsPerCta > 8 ? 8 : warpsPerCta);
  const bool acgtOnly = sigma <= 4 && kLen <= 8 && profSize == (1u << (2 * kLen));  // codes 0..3: 2-bit fast path
  const size_t smem = size_t(warpsPerCta) * profSize * 2;
  SD_CUDA(ctx, cudaFuncSetAttribute(acgtOnly ? (const void*)kmerProfileKernel4 : (const void*)kmerProfileKernel,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  uint32_t grid = (nSeqs + warpsPerCta - 1) / warpsPerCta;
  const uint32_t maxGrid = uint32_t(ctx->numSMs) * 32;  // (swept 6..1000 CTAs per SM on B200: 32-64 is best, 6 costs 5 %)
  if (grid > maxGrid) grid = maxGrid;
  if (acgtOnly) {
    kmerProfileKernel4<<<grid, warpsPerCta * 32, smem, ctx->stream>>>(sd_dev_seqs(ctx), kLen, profSize, ctx->dProfiles);
  } else {
    kmerProfileKernel<<<grid, warpsPerCta * 32, smem, ctx->stream>>>(sd_dev_seqs(ctx), kLen, sigma, top, profSize, ctx->dProfiles);
  }
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

// reads x clusters grid: a tile of cluster profiles sits in shared memory, every read profile is
// fetched from HBM once per tile into registers (VPL uint4 per lane) and compared against the whole
// tile; results for 32 clusters leave as one coalesced line.  HBM traffic is the compulsory
// (R + C) * P + R * C * 12 bytes; the work is min/add SIMD on the ALU pipe + LDS.
template <int VPL>
__global__ void __launch_bounds__(512) kmerCompareGridKernel(DevSeqs seqs, const uint16_t* profiles, uint32_t profSize,
                                                             uint32_t kLen, const uint32_t* reads, uint32_t nReads,
                                                             const uint32_t* clusters, uint32_t nClusters, uint32_t tileC,
                                                             uint32_t* sharedOut, double* fracOut) {
  extern __shared__ uint4 sTile[];  // tileC * vecs
  const uint32_t vecs = profSize / 8;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (uint32_t c0 = 0; c0 < nClusters; c0 += tileC) {
    const uint32_t nc = min(tileC, nClusters - c0);
    __syncthreads();
    for (uint32_t x = threadIdx.x; x < nc * vecs; x += blockDim.x) {
      const uint32_t c = x / vecs, v = x - c * vecs;
      sTile[x] = __ldg(reinterpret_cast<const uint4*>(profiles + size_t(clusters[c0 + c]) * profSize) + v);
    }
    __syncthreads();
    for (uint32_t r = blockIdx.x * warps + warp; r < nReads; r += gridDim.x * warps) {
      const uint32_t ir = reads[r];
      const uint4* pr = reinterpret_cast<const uint4*>(profiles + size_t(ir) * profSize);
      uint4 mine[VPL];
#pragma unroll
      for (int x = 0; x < VPL; ++x) mine[x] = __ldg(pr + lane + 32 * x);
      const uint32_t lr = uint32_t(seqs.offsets[ir + 1] - seqs.offsets[ir]);
      for (uint32_t g = 0; g < nc; g += 32) {  // 32 clusters at a time: 32 independent partial sums per lane
        uint32_t part[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          uint32_t acc = 0;
          if (g + c < nc) {
            const uint4* pc = sTile + size_t(g + c) * vecs;
#pragma unroll
            for (int x = 0; x < VPL; ++x) acc = minSum8(mine[x], pc[lane + 32 * x], acc);
          }
          part[c] = (acc & 0xffffu) + (acc >> 16);
        }
        // transpose-reduce: 31 shuffles turn 32 per-lane partials into one total per lane
        // (lane L ends up with the sum for cluster g + L)
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const uint32_t send = upper ? part[i] : part[i + off];
            const uint32_t keepv = upper ? part[i + off] : part[i];
            part[i] = keepv + __shfl_xor_sync(FULL, send, off);
          }
        }
        const uint32_t cc = g + lane;
        if (cc < nc) {  // 32 results -> one coalesced write
          const uint32_t ic = clusters[c0 + cc];
          const uint32_t lc = uint32_t(seqs.offsets[ic + 1] - seqs.offsets[ic]);
          const uint32_t minLen = lr < lc ? lr : lc;
          const size_t o = size_t(r) * nClusters + c0 + cc;
          sharedOut[o] = part[0];
          fracOut[o] = (minLen >= kLen) ? double(part[0]) / double(minLen - kLen + 1) : 0.0;
        }
      }
    }
  }
}

template <int VPL>
static int gridLaunch(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads, const uint32_t* dClusters,
                      uint32_t nClusters, uint32_t* dShared, double* dFrac) {
  const size_t profBytes = size_t(ctx->profSize) * 2;
  uint32_t tileC = uint32_t((200u << 10) / profBytes);
  if (tileC > nClusters) tileC = nClusters;
  const size_t smem = size_t(tileC) * profBytes;
  SD_CUDA(ctx, cudaFuncSetAttribute(kmerCompareGridKernel<VPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const uint32_t ctasPerSM = smem > (100u << 10) ? 1 : 2;
  // one CTA per SM when the tile fills shared memory: make it 16 warps to keep the ALU pipe fed
  kmerCompareGridKernel<VPL><<<ctx->numSMs * ctasPerSM, ctasPerSM == 1 ? 512 : 256, smem, ctx->stream>>>(
      sd_dev_seqs(ctx), ctx->dProfiles, ctx->profSize, ctx->profK, dReads, nReads, dClusters, nClusters, tileC, dShared, dFrac);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

// IUPAC complement of a base character (njhseq seqUtil::reverseComplement); 0 = none
static uint8_t complementChar(uint8_t c) {
  switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C'; case 'N': return 'N';
    case 'R': return 'Y'; case 'Y': return 'R'; case 'S': return 'S'; case 'W': return 'W'; case 'K': return 'M';
    case 'M': return 'K'; case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
    case 'a': return 't'; case 't': return 'a'; case 'c': return 'g'; case 'g': return 'c'; case 'n': return 'n';
    default: return 0;
  }
}

// Reverse-complement profiles of every profiled sequence, derived from the forward ones (once per profile build).
static int ensureRevProfiles(sdgpu_ctx* ctx) {
  if (ctx->profRevValid) return SDGPU_OK;
  const uint32_t sigma = uint32_t(ctx->usedSym < 4 ? 4 : ctx->usedSym), k = ctx->profK;
  uint32_t comp[SDGPU_MAX_SYMBOLS];
  for (uint32_t c = 0; c < sigma; ++c) {
    const uint8_t cc = complementChar(ctx->charOf[c]);
    const int code = cc ? ctx->codeOf[cc] : -1;
    if (code < 0 || uint32_t(code) >= sigma)
      return sd_fail(ctx, SDGPU_E_ALPHABET, "reverse complement: a stored character has no complement inside the store's alphabet");
    comp[c] = uint32_t(code);
  }
  uint64_t size = 1;
  for (uint32_t x = 0; x < k; ++x) size *= sigma;
  std::vector<uint32_t> rc(size);
  for (uint64_t id = 0; id < size; ++id) {  // id = base-sigma digits, first base most significant
    uint64_t v = id, out = 0;
    for (uint32_t x = 0; x < k; ++x) {
      out = out * sigma + comp[v % sigma];  // last base first, complemented
      v /= sigma;
    }
    rc[id] = uint32_t(out);
  }
  if (size > ctx->rcIndexCap) {
    if (ctx->dRcIndex) cudaFree(ctx->dRcIndex);
    ctx->dRcIndex = nullptr;
    ctx->rcIndexCap = 0;
    SD_CUDA(ctx, cudaMalloc(&ctx->dRcIndex, size * 4));
    ctx->rcIndexCap = uint32_t(size);
  }
  const uint64_t need = uint64_t(ctx->profSeqs) * ctx->profSize * 2;
  if (need > ctx->profRevCapBytes) {
    if (ctx->dProfilesRev) cudaFree(ctx->dProfilesRev);
    ctx->dProfilesRev = nullptr;
    ctx->profRevCapBytes = 0;
    SD_CUDA(ctx, cudaMalloc(&ctx->dProfilesRev, need));
    ctx->profRevCapBytes = need;
  }
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dRcIndex, rc.data(), size * 4, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // rc is a local
  kmerProfileRevKernel<<<(ctx->profSeqs + 3) / 4, 128, 0, ctx->stream>>>(ctx->dProfiles, ctx->dRcIndex, uint32_t(size), ctx->profSize,
                                                                       ctx->profSeqs, ctx->dProfilesRev);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  ctx->profRevValid = true;
  return SDGPU_OK;
}

static int compareLaunch(sdgpu_ctx* ctx, const uint32_t* dA, const uint32_t* dB, uint64_t nPairs, uint32_t nCols,
                         uint32_t* dShared, double* dFrac, int revCompB = 0) {
  if (revCompB) {
    int rc = ensureRevProfiles(ctx);
    if (rc) return rc;
  }
  if (nCols && !revCompB) {  // dense grid: use the tiled kernel when a profile splits evenly over the lanes
    const uint32_t vecs = ctx->profSize / 8, nReads = uint32_t(nPairs / nCols);
    if (vecs % 32 == 0 && size_t(ctx->profSize) * 2 <= (200u << 10)) {
      switch (vecs / 32) {
        case 1: return gridLaunch<1>(ctx, dA, nReads, dB, nCols, dShared, dFrac);
        case 2: return gridLaunch<2>(ctx, dA, nReads, dB, nCols, dShared, dFrac);
        case 4: return gridLaunch<4>(ctx, dA, nReads, dB, nCols, dShared, dFrac);
        case 8: return gridLaunch<8>(ctx, dA, nReads, dB, nCols, dShared, dFrac);
        case 16: return gridLaunch<16>(ctx, dA, nReads, dB, nCols, dShared, dFrac);
        default: break;
      }
    }
  }
  uint64_t ctas = (nPairs + 3) / 4;
  const uint64_t maxGrid = uint64_t(ctx->numSMs) * 16;
  if (ctas > maxGrid) ctas = maxGrid;
  kmerCompareKernel<<<uint32_t(ctas), 128, 0, ctx->stream>>>(sd_dev_seqs(ctx), ctx->dProfiles, revCompB ? ctx->dProfilesRev : ctx->dProfiles,
                                                            ctx->profSize, ctx->profK, dA, dB, nPairs, nCols, dShared, dFrac);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

int sd_kmer_compare_all_dev(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads, const uint32_t* dClusters,
                            uint32_t nClusters, uint32_t* dShared, double* dFrac) {
  if (!ctx->dProfiles) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_build_kmer_profiles has not been called");
  const uint64_t nPairs = uint64_t(nReads) * nClusters;
  if (nPairs == 0) return SDGPU_OK;
  return compareLaunch(ctx, dReads, dClusters, nPairs, nClusters, dShared, dFrac);
}

int sd_kmer_compare_pairs(sdgpu_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t nPairs, int revCompB,
                          uint32_t* sharedOut, double* fracOut) {
  if (!ctx->dProfiles) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_build_kmer_profiles has not been called");
  if (nPairs == 0) return SDGPU_OK;
  for (uint32_t x = 0; x < nPairs; ++x)
    if (a[x] >= ctx->profSeqs || b[x] >= ctx->profSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index outside the profiled store");
  const size_t bytes = size_t(nPairs) * (4 + 4 + 4 + 8) + 64;
  int rc = sd_ensure_stage(ctx, bytes);
  if (rc) return rc;
  double* dFrac = reinterpret_cast<double*>(ctx->dStage);
  uint32_t* dA = reinterpret_cast<uint32_t*>(dFrac + nPairs);
  uint32_t* dB = dA + nPairs;
  uint32_t* dShared = dB + nPairs;
  SD_CUDA(ctx, cudaMemcpyAsync(dA, a, size_t(nPairs) * 4, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dB, b, size_t(nPairs) * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = compareLaunch(ctx, dA, dB, nPairs, 0, dShared, dFrac, revCompB);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(sharedOut, dShared, size_t(nPairs) * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(fracOut, dFrac, size_t(nPairs) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sd_kmer_compare_all(sdgpu_ctx* ctx, const uint32_t* reads, uint32_t nReads, const uint32_t* clusters,
                        uint32_t nClusters, uint32_t* sharedOut, double* fracOut) {
  if (!ctx->dProfiles) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_build_kmer_profiles has not been called");
  const uint64_t nPairs = uint64_t(nReads) * nClusters;
  if (nPairs == 0) return SDGPU_OK;
  for (uint32_t x = 0; x < nReads; ++x)
    if (reads[x] >= ctx->profSeqs) return sd_fail(ctx, SDGPU_E_ARG, "read index outside the profiled store");
  for (uint32_t x = 0; x < nClusters; ++x)
    if (clusters[x] >= ctx->profSeqs) return sd_fail(ctx, SDGPU_E_ARG, "cluster index outside the profiled store");
  const size_t bytes = nPairs * 12 + (size_t(nReads) + nClusters) * 4 + 64;
  int rc = sd_ensure_stage(ctx, bytes);
  if (rc) return rc;
  double* dFrac = reinterpret_cast<double*>(ctx->dStage);
  uint32_t* dShared = reinterpret_cast<uint32_t*>(dFrac + nPairs);
  uint32_t* dA = dShared + nPairs;
  uint32_t* dB = dA + nReads;
  SD_CUDA(ctx, cudaMemcpyAsync(dA, reads, size_t(nReads) * 4, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dB, clusters, size_t(nClusters) * 4, cudaMemcpyHostToDevice, ctx->stream));
  rc = compareLaunch(ctx, dA, dB, nPairs, nClusters, dShared, dFrac);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(sharedOut, dShared, nPairs * 4, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(fracOut, dFrac, nPairs * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}
