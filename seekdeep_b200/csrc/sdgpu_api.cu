// sdgpu_api.cu — the C ABI declared in include/sdgpu.h: context, sequence store, dispatch.
#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

#include "sdgpu_internal.cuh"

// host_translate.cpp
extern "C" void sd_translate_bases(const uint8_t* src, uint8_t* dst, uint64_t len, const int16_t* codeOf, int* maxCode, bool* unknown);

static thread_local std::string g_createErr;

int sd_fail(sdgpu_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

static int growDevice(sdgpu_ctx* ctx, void** p, uint64_t* cap, uint64_t need, uint64_t keepBytes) {
  if (need <= *cap) return SDGPU_OK;
  uint64_t n = std::max<uint64_t>(need, *cap * 2);
  n = (n + 255) & ~255ull;
  void* q = nullptr;
  SD_CUDA(ctx, cudaMalloc(&q, n));
  if (*p && keepBytes) SD_CUDA(ctx, cudaMemcpyAsync(q, *p, keepBytes, cudaMemcpyDeviceToDevice, ctx->stream));
  if (*p) {
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(*p);
  }
  *p = q;
  *cap = n;
  return SDGPU_OK;
}

int sd_ensure_scratch(sdgpu_ctx* ctx, uint64_t bytes) {
  return growDevice(ctx, reinterpret_cast<void**>(&ctx->dScratch), &ctx->scratchBytes, bytes, 0);
}
int sd_ensure_stage(sdgpu_ctx* ctx, uint64_t bytes) {
  return growDevice(ctx, &ctx->dStage, &ctx->stageBytes, bytes, 0);
}
int sd_ensure_pinned(sdgpu_ctx* ctx, uint64_t bytes) {
  if (bytes <= ctx->pinnedBytes) return SDGPU_OK;
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  ctx->hPinned = nullptr;
  ctx->pinnedBytes = 0;
  const uint64_t n = std::max<uint64_t>(bytes, 1u << 20);
  SD_CUDA(ctx, cudaMallocHost(&ctx->hPinned, n));
  ctx->pinnedBytes = n;
  return SDGPU_OK;
}

DevSeqs sd_dev_seqs(const sdgpu_ctx* ctx) {
  DevSeqs s;
  s.codes = ctx->dCodes;
  s.quals = ctx->dQuals;
  s.flags = ctx->dFlags;
  s.offsets = ctx->dOffsets;
  s.cnts = ctx->dCnts;
  s.n = uint32_t(ctx->hOffsets.size() - 1);
  return s;
}

DevKmerIndex sd_dev_index(const sdgpu_ctx* ctx) {
  DevKmerIndex k;
  k.keys = ctx->dIdxKeys;
  k.counts = ctx->dIdxCnts;
  k.n = ctx->nIdx;
  k.kLen = ctx->idxK;
  k.runCutOff = ctx->idxRunCutOff;
  k.byPos = ctx->idxByPos;
  k.posStart = (ctx->idxByPos && ctx->idxNPos) ? ctx->dIdxPosStart : nullptr;
  k.nPos = ctx->idxNPos;
  return k;
}

// (Re)derive the device scoring block from params + the current alphabet.
static int refreshScoring(sdgpu_ctx* ctx) {
  const sdgpu_params& p = ctx->params;
  DevScoring& s = ctx->scoring;
  s.go = p.gaps.gapOpen;
  s.ge = p.gaps.gapExtend;
  s.goLQ = p.gaps.gapLeftQueryOpen;
  s.geLQ = p.gaps.gapLeftQueryExtend;
  s.goRQ = p.gaps.gapRightQueryOpen;
  s.geRQ = p.gaps.gapRightQueryExtend;
  s.goLR = p.gaps.gapLeftRefOpen;
  s.geLR = p.gaps.gapLeftRefExtend;
  s.goRR = p.gaps.gapRightRefOpen;
  s.geRR = p.gaps.gapRightRefExtend;
  s.primaryQual = int(p.primaryQual);
  s.secondaryQual = int(p.secondaryQual);
  s.window = int(p.qualThresWindow);
  s.countEndGaps = p.countEndGaps ? 1 : 0;
  s.weighHomopolymers = p.weighHomopolymers ? 1 : 0;
  for (int a = 0; a < SDGPU_MAX_SYMBOLS; ++a)
    for (int b = 0; b < SDGPU_MAX_SYMBOLS; ++b) {
      int v = 0;
      if (a < ctx->nSym && b < ctx->nSym) v = p.scoreMat[size_t(ctx->charOf[a]) * SDGPU_MAT_DIM + ctx->charOf[b]];
      if (v < -32768 || v > 32767) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "score matrix entry outside int16");
      s.score[a * SDGPU_MAX_SYMBOLS + b] = int16_t(v);
    }
  return SDGPU_OK;
}

__global__ void scatterCntsKernel(const uint32_t* idx, const double* vals, uint32_t n, double* cnts) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < n) cnts[idx[x]] = vals[x];
}

static int validateParams(sdgpu_ctx* ctx, const sdgpu_params* p) {
  const int32_t* g = reinterpret_cast<const int32_t*>(&p->gaps);
  for (int x = 0; x < 10; ++x)
    if (g[x] < 0 || g[x] > 1000000) return sd_fail(ctx, SDGPU_E_ARG, "gap penalties must be in [0, 1e6]");
  if (p->qualThresWindow > 64) return sd_fail(ctx, SDGPU_E_ARG, "qualThresWindow > 64");
  return SDGPU_OK;
}

extern "C" {

int sdgpu_abi_version(void) { return SDGPU_ABI_VERSION; }

int sdgpu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int sdgpu_create(int device, const sdgpu_params* params, sdgpu_ctx** out) {
  if (!params || !out) {
    g_createErr = "sdgpu_create: null argument";
    return SDGPU_E_ARG;
  }
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n) {
    g_createErr = std::string("sdgpu_create: no usable CUDA device (") +
                  (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range") + "); there is no CPU fallback";
    cudaGetLastError();
    return SDGPU_E_CUDA;
  }
  sdgpu_ctx* ctx = new (std::nothrow) sdgpu_ctx();
  if (!ctx) return SDGPU_E_ARG;
  ctx->device = device;
  auto bail = [&](int rc) {
    g_createErr = ctx->err;
    sdgpu_destroy(ctx);
    return rc;
  };
  if (validateParams(ctx, params)) return bail(SDGPU_E_ARG);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&ctx->numSMs, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->dWork, SD_W_WORDS * 4)) != cudaSuccess || (e = cudaMemset(ctx->dWork, 0, SD_W_WORDS * 4)) != cudaSuccess ||
      (e = cudaHostAlloc(reinterpret_cast<void**>(&ctx->hBandStat), 2 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) {
    ctx->err = std::string("sdgpu_create: ") + cudaGetErrorString(e);
    return bail(SDGPU_E_CUDA);
  }
  ctx->hBandStat[0] = ctx->hBandStat[1] = 0;
  ctx->params = *params;
  for (int x = 0; x < 256; ++x) ctx->codeOf[x] = -1;
  // fixed head of the alphabet so that codes are stable across runs
  const char* head = "ACGTN";
  for (int x = 0; head[x]; ++x) {
    ctx->codeOf[uint8_t(head[x])] = int16_t(ctx->nSym);
    ctx->charOf[ctx->nSym++] = uint8_t(head[x]);
  }
  int rc = refreshScoring(ctx);
  if (rc) return bail(rc);
  *out = ctx;
  return SDGPU_OK;
}

void sdgpu_destroy(sdgpu_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  cudaFree(ctx->dCodes);
  cudaFree(ctx->dQuals);
  cudaFree(ctx->dFlags);
  cudaFree(ctx->dIdxPosStart);
  cudaFree(ctx->dOffsets);
  cudaFree(ctx->dCnts);
  cudaFree(ctx->dIdxKeys);
  cudaFree(ctx->dIdxCnts);
  cudaFree(ctx->dProfiles);
  cudaFree(ctx->dProfilesRev);
  cudaFree(ctx->dKmerLists);
  cudaFree(ctx->dCharOf);
  cudaFree(ctx->dRcIndex);
  cudaFree(ctx->dScratch);
  cudaFree(ctx->dWork);
  cudaFree(ctx->dRedo);
  cudaFree(ctx->dRedo2);
  cudaFree(ctx->dScreenItems);
  cudaFree(ctx->dGridIdx);
  for (auto& G : ctx->gridSlots) {
    if (G.hHits) cudaFreeHost(G.hHits);
    if (G.hCount) cudaFreeHost(G.hCount);
    if (G.done) cudaEventDestroy(G.done);
  }
  if (ctx->bandStatEv) cudaEventDestroy(ctx->bandStatEv);
  if (ctx->hBandStat) cudaFreeHost(ctx->hBandStat);
  cudaFree(ctx->dStage);
  cudaFree(ctx->dLines);
  cudaFree(ctx->dIdxWork);
  for (auto& S : ctx->passSlots) {
    if (S.hPinned) cudaFreeHost(S.hPinned);
    cudaFree(S.dBuf);
    if (S.done) cudaEventDestroy(S.done);
  }
  for (cudaEvent_t e : ctx->evPool)
    if (e) cudaEventDestroy(e);
  if (ctx->hPinned) cudaFreeHost(ctx->hPinned);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  cudaGetLastError();
  delete ctx;
}

const char* sdgpu_last_error(sdgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

int sdgpu_set_params(sdgpu_ctx* ctx, const sdgpu_params* params) {
  if (!ctx || !params) return SDGPU_E_ARG;
  int rc = validateParams(ctx, params);
  if (rc) return rc;
  const bool qualChanged = params->primaryQual != ctx->params.primaryQual || params->secondaryQual != ctx->params.secondaryQual ||
                           params->qualThresWindow != ctx->params.qualThresWindow;
  ctx->params = *params;
  rc = refreshScoring(ctx);
  if (rc || !qualChanged) return rc;
  // the per-base checkQual flags follow qScorePars_
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  if ((rc = sd_refresh_qual_flags(ctx, 0, nSeqs))) return rc;
  if ((rc = sd_refresh_kmer_flags(ctx, 0, nSeqs))) return rc;
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_get_params(sdgpu_ctx* ctx, sdgpu_params* out) {
  if (!ctx || !out) return SDGPU_E_ARG;
  *out = ctx->params;
  return SDGPU_OK;
}

// Where an upload's sequences come from: packed arrays (bases / quals / offsets) or one pointer per sequence.
struct SeqSource {
  const uint8_t* bases = nullptr;
  const uint8_t* quals = nullptr;
  const uint64_t* offsets = nullptr;       // packed form: n + 1 entries
  const uint8_t* const* seqPtrs = nullptr;  // gathered form
  const uint8_t* const* qualPtrs = nullptr;
  const uint32_t* lens = nullptr;
  const uint8_t* seq(uint32_t s) const { return seqPtrs ? seqPtrs[s] : bases + offsets[s]; }
  const uint8_t* qual(uint32_t s) const { return seqPtrs ? (qualPtrs ? qualPtrs[s] : nullptr) : (quals ? quals + offsets[s] : nullptr); }
};

// `replace`: the new sequences become the whole store (sdgpu_upload_seqs).  Nothing of the ctx changes
// before the batch has been validated: a rejected call leaves the previous store and alphabet as they were.
static int appendImpl(sdgpu_ctx* ctx, const SeqSource& src, const double* cnts, uint32_t n, uint32_t* firstIndex, bool replace) {
  if (!ctx || (n && !(src.seqPtrs ? src.lens != nullptr : (src.bases && src.offsets)))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t oldN = replace ? 0u : uint32_t(ctx->hOffsets.size() - 1);
  if (firstIndex) *firstIndex = oldN;
  if (n == 0) {
    if (replace) {
      ctx->hOffsets.assign(1, 0);
      ctx->maxLen = 0;
      ctx->usedSym = 0;
      ctx->haveIndex = false;
      ctx->profSeqs = 0;
      ctx->listSeqs = 0;
    }
    return SDGPU_OK;
  }
  if (!src.seqPtrs && src.offsets[0] != 0) return sd_fail(ctx, SDGPU_E_ARG, "offsets[0] must be 0");
  const uint64_t oldBases = replace ? 0ull : ctx->hOffsets.back();
  // offsets of the new sequences inside the batch
  std::vector<uint64_t> off(size_t(n) + 1, 0);
  bool newSym = false;
  uint32_t maxLen = replace ? 0u : ctx->maxLen;
  int usedSym = replace ? 0 : ctx->usedSym;
  for (uint32_t s = 0; s < n; ++s) {
    uint64_t len;
    if (src.seqPtrs) {
      if (!src.seqPtrs[s]) return sd_fail(ctx, SDGPU_E_ARG, "null sequence pointer");
      len = src.lens[s];
    } else {
      if (src.offsets[s + 1] < src.offsets[s]) return sd_fail(ctx, SDGPU_E_ARG, "offsets must be non-decreasing");
      len = src.offsets[s + 1] - src.offsets[s];
    }
    if (len == 0) return sd_fail(ctx, SDGPU_E_ARG, "empty sequence");
    if (len > 0x7fffffffu) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "sequence too long");
    maxLen = std::max<uint32_t>(maxLen, uint32_t(len));
    off[s + 1] = off[s] + len;
  }
  const uint64_t addBases = off[n];
  // alphabet + code translation on the host, straight into pinned staging (one pass; the store is write-once, read-many)
  int rc = sd_ensure_pinned(ctx, addBases * 2 + 64);
  if (rc) return rc;
  uint8_t* hCodes = reinterpret_cast<uint8_t*>(ctx->hPinned);
  uint8_t* hQuals = hCodes + addBases;
  // translate with the alphabet as it stands, a few threads side by side; only if a character without
  // a code turns up (first upload with degenerate bases) the sequential pass below assigns codes
  bool unknown = false;
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    unsigned hw = std::thread::hardware_concurrency();
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = unsigned(CPU_COUNT(&set));
    const uint64_t T = addBases < (1u << 20) ? 1 : std::min<uint64_t>(hw ? hw : 1, 8);
    std::vector<int> maxCode(T, -1);
    std::vector<uint8_t> sawUnknown(T, 0);
    std::vector<uint32_t> cut(T + 1, n);  // thread t takes sequences [cut[t], cut[t + 1]): about the same number of bases each
    cut[0] = 0;
    for (uint64_t t = 1; t < T; ++t) cut[t] = uint32_t(std::lower_bound(off.begin(), off.end(), addBases * t / T) - off.begin());
    auto work = [&](uint64_t t) {
      int mc = -1;
      bool unk = false;
      for (uint32_t s = cut[t]; s < cut[t + 1]; ++s) {
        const uint8_t* b = src.seq(s);
        const uint8_t* q = src.qual(s);
        const uint64_t len = off[s + 1] - off[s];
        sd_translate_bases(b, hCodes + off[s], len, ctx->codeOf, &mc, &unk);
        if (q) {
          std::memcpy(hQuals + off[s], q, len);
        } else {
          std::memset(hQuals + off[s], 0, len);
        }
      }
      maxCode[t] = mc;
      sawUnknown[t] = unk;
    };
    std::vector<std::thread> th;
    for (uint64_t t = 1; t < T; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    for (uint64_t t = 0; t < T; ++t) {
      unknown |= sawUnknown[t] != 0;
      if (maxCode[t] + 1 > usedSym) usedSym = maxCode[t] + 1;
    }
  }
  if (unknown) {
    // which characters are new, and do they fit?  (checked before any of them gets a code)
    bool isNew[256] = {false};
    int nNew = 0;
    for (uint32_t s = 0; s < n; ++s) {
      const uint8_t* b = src.seq(s);
      for (uint64_t x = 0, len = off[s + 1] - off[s]; x < len; ++x) {
        const uint8_t ch = b[x];
        if (ctx->codeOf[ch] >= 0 || isNew[ch]) continue;
        if (ch >= 128) return sd_fail(ctx, SDGPU_E_ALPHABET, "base outside 7-bit ASCII");
        isNew[ch] = true;
        if (ctx->nSym + ++nNew > SDGPU_MAX_SYMBOLS) return sd_fail(ctx, SDGPU_E_ALPHABET, "more than 16 distinct base characters");
      }
    }
    for (uint32_t s = 0; s < n; ++s) {
      const uint8_t* b = src.seq(s);
      uint8_t* dst = hCodes + off[s];
      for (uint64_t x = 0, len = off[s + 1] - off[s]; x < len; ++x) {
        const uint8_t ch = b[x];
        int code = ctx->codeOf[ch];
        if (code < 0) {
          code = ctx->nSym;
          ctx->codeOf[ch] = int16_t(code);
          ctx->charOf[ctx->nSym++] = ch;
          newSym = true;
        }
        dst[x] = uint8_t(code);
        if (code + 1 > usedSym) usedSym = code + 1;
      }
    }
  }
  if (newSym && (rc = refreshScoring(ctx))) return rc;
  ctx->usedSym = usedSym;
  if (replace) {  // validated: the old store goes
    ctx->hOffsets.assign(1, 0);
    ctx->maxLen = 0;
    ctx->haveIndex = false;
    ctx->profSeqs = 0;
    ctx->listSeqs = 0;
  }
  // grow device arrays
  uint64_t capB = ctx->capBases, capB2 = ctx->capBases, capB3 = ctx->capBases;
  // (64 spare bytes behind the last base, kept at code 0: the score-only screen pass prefetches a few bases
  //  past the end of a sequence and must find valid codes there)
  if ((rc = growDevice(ctx, reinterpret_cast<void**>(&ctx->dCodes), &capB, oldBases + addBases + 64, oldBases))) return rc;
  if ((rc = growDevice(ctx, reinterpret_cast<void**>(&ctx->dQuals), &capB2, oldBases + addBases + 64, oldBases))) return rc;
  if ((rc = growDevice(ctx, reinterpret_cast<void**>(&ctx->dFlags), &capB3, oldBases + addBases + 64, oldBases))) return rc;
  ctx->capBases = std::min(std::min(capB, capB2), capB3);
  const uint64_t needSeqs = uint64_t(oldN) + n + 1;
  uint64_t capO = uint64_t(ctx->capSeqs) * 8, capC = uint64_t(ctx->capSeqs) * 8;
  if ((rc = growDevice(ctx, reinterpret_cast<void**>(&ctx->dOffsets), &capO, needSeqs * 8, (uint64_t(oldN) + 1) * 8))) return rc;
  if ((rc = growDevice(ctx, reinterpret_cast<void**>(&ctx->dCnts), &capC, needSeqs * 8, uint64_t(oldN) * 8))) return rc;
  ctx->capSeqs = uint32_t(std::min(capO, capC) / 8);
  ctx->h2dBytes += 2 * addBases + (uint64_t(n) + 1) * 8 + uint64_t(n) * 8;
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dCodes + oldBases, hCodes, addBases, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dQuals + oldBases, hQuals, addBases, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemsetAsync(ctx->dCodes + oldBases + addBases, 0, 64, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pinned staging is reused below
  ctx->hOffsets.reserve(needSeqs);
  for (uint32_t s = 1; s <= n; ++s) ctx->hOffsets.push_back(oldBases + off[s]);
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dOffsets + oldN, ctx->hOffsets.data() + oldN, (uint64_t(n) + 1) * 8,
                               cudaMemcpyHostToDevice, ctx->stream));
  double* hC = reinterpret_cast<double*>(ctx->hPinned);
  if ((rc = sd_ensure_pinned(ctx, uint64_t(n) * 8))) return rc;
  hC = reinterpret_cast<double*>(ctx->hPinned);
  for (uint32_t s = 0; s < n; ++s) hC[s] = cnts ? cnts[s] : 1.0;
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dCnts + oldN, hC, uint64_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
  ctx->maxLen = maxLen;
  // per-base flags of the new sequences: checkQual outcome, and the low-frequency k-mer bit when a table exists
  if ((rc = sd_refresh_qual_flags(ctx, oldN, n))) return rc;
  if ((rc = sd_refresh_kmer_flags(ctx, oldN, n))) return rc;
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_upload_seqs(sdgpu_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                      const double* cnts, uint32_t n) {
  if (!ctx) return SDGPU_E_ARG;
  SeqSource src;
  src.bases = bases, src.quals = quals, src.offsets = offsets;
  return appendImpl(ctx, src, cnts, n, nullptr, true);
}

int sdgpu_upload_seqs_gather(sdgpu_ctx* ctx, const uint8_t* const* seqPtrs, const uint8_t* const* qualPtrs, const uint32_t* lens,
                             const double* cnts, uint32_t n) {
  if (!ctx) return SDGPU_E_ARG;
  SeqSource src;
  src.seqPtrs = seqPtrs, src.qualPtrs = qualPtrs, src.lens = lens;
  return appendImpl(ctx, src, cnts, n, nullptr, true);
}

int sdgpu_append_seqs(sdgpu_ctx* ctx, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets,
                      const double* cnts, uint32_t n, uint32_t* firstIndex) {
  SeqSource src;
  src.bases = bases, src.quals = quals, src.offsets = offsets;
  return appendImpl(ctx, src, cnts, n, firstIndex, false);
}

int sdgpu_get_alphabet(sdgpu_ctx* ctx, uint8_t chars[16], uint32_t* n) {
  if (!ctx || !chars || !n) return SDGPU_E_ARG;
  for (int x = 0; x < ctx->nSym; ++x) chars[x] = ctx->charOf[x];
  *n = uint32_t(ctx->nSym);
  return SDGPU_OK;
}

int sdgpu_num_seqs(sdgpu_ctx* ctx, uint32_t* n) {
  if (!ctx || !n) return SDGPU_E_ARG;
  *n = uint32_t(ctx->hOffsets.size() - 1);
  return SDGPU_OK;
}

int sdgpu_update_cnts(sdgpu_ctx* ctx, const uint32_t* seqIdx, const double* cnts, uint32_t n) {
  if (!ctx || (n && (!seqIdx || !cnts))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  for (uint32_t x = 0; x < n; ++x)
    if (seqIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "sequence index out of range");
  if (n == 0) return SDGPU_OK;
  // one staged copy + a scatter kernel (after a collapse the live clusters' store indices are scattered:
  // per-run copies meant ~20 000 tiny transfers)
  const uint64_t bytes = uint64_t(n) * 12;
  int rc = sd_ensure_pinned(ctx, bytes);
  if (rc) return rc;
  if ((rc = sd_ensure_stage(ctx, bytes + 256))) return rc;
  double* hVals = reinterpret_cast<double*>(ctx->hPinned);
  uint32_t* hIdx = reinterpret_cast<uint32_t*>(hVals + n);
  std::memcpy(hVals, cnts, size_t(n) * 8);
  std::memcpy(hIdx, seqIdx, size_t(n) * 4);
  double* dVals = reinterpret_cast<double*>(ctx->dStage);
  uint32_t* dIdx = reinterpret_cast<uint32_t*>(dVals + n);
  SD_CUDA(ctx, cudaMemcpyAsync(dVals, hVals, bytes, cudaMemcpyHostToDevice, ctx->stream));
  scatterCntsKernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(dIdx, dVals, n, ctx->dCnts);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  ctx->h2dBytes += bytes;
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_build_kmer_index(sdgpu_ctx* ctx, const uint32_t* seqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                           int kmersByPosition, int expandKmerPos, uint32_t expandKmerSize) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_build_kmer_index(ctx, seqIdx, nIdx, kLen, runCutOff, kmersByPosition, expandKmerPos, expandKmerSize);
}

int sdgpu_kmer_index_lookup(sdgpu_ctx* ctx, const uint8_t* kmers, const uint32_t* positions, uint32_t n,
                            uint32_t* countsOut) {
  if (!ctx || (n && (!kmers || !countsOut))) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_kmer_index_lookup(ctx, kmers, positions, n, countsOut);
}

int sdgpu_build_kmer_profiles(sdgpu_ctx* ctx, uint32_t kLen) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_build_kmer_profiles(ctx, kLen);
}

int sdgpu_build_kmer_lists(sdgpu_ctx* ctx, uint32_t kLen) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_build_kmer_lists(ctx, kLen);
}

int sdgpu_kmer_bin_masks(sdgpu_ctx* ctx, const uint32_t* seeds, uint32_t nSeeds, const uint32_t* reads, uint32_t nReads,
                         double cutOff, uint64_t* maskOut, uint32_t* sharedOut) {
  if (!ctx || (nSeeds && !seeds) || (nReads && (!reads || !maskOut))) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_kmer_bin_masks(ctx, seeds, nSeeds, reads, nReads, cutOff, maskOut, sharedOut);
}

int sdgpu_kmer_compare(sdgpu_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t nPairs, int revCompB,
                       uint32_t* sharedOut, double* fracOut) {
  if (!ctx || (nPairs && (!a || !b || !sharedOut || !fracOut))) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_kmer_compare_pairs(ctx, a, b, nPairs, revCompB, sharedOut, fracOut);
}

int sdgpu_kmer_compare_all_dev(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads, const uint32_t* dClusters,
                               uint32_t nClusters, uint32_t* dShared, double* dFrac) {
  if (!ctx || !dReads || !dClusters || !dShared || !dFrac) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_kmer_compare_all_dev(ctx, dReads, nReads, dClusters, nClusters, dShared, dFrac);
}

int sdgpu_kmer_compare_all(sdgpu_ctx* ctx, const uint32_t* reads, uint32_t nReads, const uint32_t* clusters,
                           uint32_t nClusters, uint32_t* sharedOut, double* fracOut) {
  if (!ctx || !reads || !clusters || !sharedOut || !fracOut) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_kmer_compare_all(ctx, reads, nReads, clusters, nClusters, sharedOut, fracOut);
}

int sdgpu_align_profile(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs,
                        uint32_t flags, sdgpu_comparison* out, sdgpu_gap* gapsOut, uint32_t gapCap) {
  if (!ctx || (nPairs && (!refIdx || !readIdx || !out))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (refIdx[x] >= nSeqs || readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[refIdx[x] + 1] - ctx->hOffsets[refIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  const bool wantGaps = gapsOut != nullptr && gapCap > 0;
  const uint64_t bytesIdx = uint64_t(nPairs) * 4;
  const uint64_t bytesOut = uint64_t(nPairs) * sizeof(sdgpu_comparison);
  const uint64_t bytesGaps = wantGaps ? uint64_t(nPairs) * gapCap * sizeof(sdgpu_gap) : 0;
  int rc = sd_ensure_stage(ctx, bytesOut + bytesGaps + 2 * bytesIdx + 256);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  sdgpu_comparison* dOut = reinterpret_cast<sdgpu_comparison*>(base);
  sdgpu_gap* dGaps = wantGaps ? reinterpret_cast<sdgpu_gap*>(base + bytesOut) : nullptr;
  uint32_t* dRef = reinterpret_cast<uint32_t*>(base + bytesOut + ((bytesGaps + 15) & ~15ull));
  uint32_t* dRead = dRef + nPairs;
  SD_CUDA(ctx, cudaMemcpyAsync(dRef, refIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dRead, readIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  if (wantGaps) SD_CUDA(ctx, cudaMemsetAsync(dGaps, 0, bytesGaps, ctx->stream));
  rc = sd_launch_align(ctx, dRef, dRead, nPairs, flags, dOut, dGaps, wantGaps ? gapCap : 0, maxA, maxB);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(out, dOut, bytesOut, cudaMemcpyDeviceToHost, ctx->stream));
  if (wantGaps) SD_CUDA(ctx, cudaMemcpyAsync(gapsOut, dGaps, bytesGaps, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += bytesOut + bytesGaps;
  return SDGPU_OK;
}

int sdgpu_align_gap_lists(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs, uint32_t gapCap,
                          const uint32_t** recsOut, uint64_t* nWordsOut) {
  if (!ctx || !recsOut || !nWordsOut || (nPairs && (!refIdx || !readIdx)) || gapCap == 0) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  *recsOut = nullptr;
  *nWordsOut = 0;
  ctx->gapRecs.clear();
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (refIdx[x] >= nSeqs || readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[refIdx[x] + 1] - ctx->hOffsets[refIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  const uint64_t bytesIdx = uint64_t(nPairs) * 4;
  const uint64_t bytesOut = uint64_t(nPairs) * sizeof(sdgpu_comparison);
  const uint64_t bytesGaps = (uint64_t(nPairs) * gapCap * sizeof(sdgpu_gap) + 15) & ~15ull;
  const uint64_t bytesPacked = (uint64_t(nPairs) * (2 + 3ull * gapCap) * 4 + 15) & ~15ull;
  int rc = sd_ensure_stage(ctx, bytesOut + bytesGaps + bytesPacked + 2 * bytesIdx + 256);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  sdgpu_comparison* dOut = reinterpret_cast<sdgpu_comparison*>(base);
  sdgpu_gap* dGaps = reinterpret_cast<sdgpu_gap*>(base + bytesOut);
  uint32_t* dPacked = reinterpret_cast<uint32_t*>(base + bytesOut + bytesGaps);
  unsigned long long* dNWords = reinterpret_cast<unsigned long long*>(base + bytesOut + bytesGaps + bytesPacked);
  uint32_t* dRef = reinterpret_cast<uint32_t*>(dNWords + 2);
  uint32_t* dRead = dRef + nPairs;
  SD_CUDA(ctx, cudaMemcpyAsync(dRef, refIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dRead, readIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemsetAsync(dGaps, 0, bytesGaps, ctx->stream));
  SD_CUDA(ctx, cudaMemsetAsync(dNWords, 0, 16, ctx->stream));
  rc = sd_launch_align(ctx, dRef, dRead, nPairs, 0u, dOut, dGaps, gapCap, maxA, maxB);
  if (rc) return rc;
  if ((rc = sd_launch_pack_gaps(ctx, dOut, dGaps, gapCap, nPairs, dPacked, dNWords))) return rc;
  unsigned long long nWords = 0;
  SD_CUDA(ctx, cudaMemcpyAsync(&nWords, dNWords, 8, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->gapRecs.resize(nWords);
  if (nWords) {
    SD_CUDA(ctx, cudaMemcpyAsync(ctx->gapRecs.data(), dPacked, nWords * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += 8 + nWords * 4;
  *recsOut = ctx->gapRecs.data();
  *nWordsOut = nWords;
  return SDGPU_OK;
}

int sdgpu_align_events(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs, uint32_t flags,
                       sdgpu_comparison* out, sdgpu_event* eventsOut, uint32_t eventCap, uint32_t* nEventsOut) {
  if (!ctx || (nPairs && (!refIdx || !readIdx || !nEventsOut || (eventCap && !eventsOut)))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (refIdx[x] >= nSeqs || readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[refIdx[x] + 1] - ctx->hOffsets[refIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  std::vector<sdgpu_comparison> localOut;
  if (!out) {
    localOut.resize(nPairs);
    out = localOut.data();
  }
  const uint64_t bytesIdx = uint64_t(nPairs) * 4;
  const uint64_t bytesOut = uint64_t(nPairs) * sizeof(sdgpu_comparison);
  const uint64_t bytesEv = (uint64_t(nPairs) * eventCap * sizeof(sdgpu_event) + 15) & ~15ull;
  uint32_t gapCap = 16;
  for (int attempt = 0;; ++attempt) {  // a second round only when some pair has more gap runs than the first one had room for
    const uint64_t bytesGaps = (uint64_t(nPairs) * gapCap * sizeof(sdgpu_gap) + 15) & ~15ull;
    int rc = sd_ensure_stage(ctx, bytesOut + bytesGaps + bytesEv + 3 * bytesIdx + 256);
    if (rc) return rc;
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
    sdgpu_comparison* dOut = reinterpret_cast<sdgpu_comparison*>(base);
    sdgpu_gap* dGaps = reinterpret_cast<sdgpu_gap*>(base + bytesOut);
    sdgpu_event* dEvents = reinterpret_cast<sdgpu_event*>(base + bytesOut + bytesGaps);
    uint32_t* dN = reinterpret_cast<uint32_t*>(base + bytesOut + bytesGaps + bytesEv);
    uint32_t* dRef = dN + nPairs;
    uint32_t* dRead = dRef + nPairs;
    SD_CUDA(ctx, cudaMemcpyAsync(dRef, refIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(dRead, readIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
    SD_CUDA(ctx, cudaMemsetAsync(dGaps, 0, bytesGaps, ctx->stream));
    rc = sd_launch_align(ctx, dRef, dRead, nPairs, flags, dOut, dGaps, gapCap, maxA, maxB);
    if (rc) return rc;
    rc = sd_launch_events(ctx, dRef, dRead, nPairs, flags, dOut, dGaps, gapCap, dEvents, eventCap, dN);
    if (rc) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(out, dOut, bytesOut, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaMemcpyAsync(nEventsOut, dN, bytesIdx, cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->h2dBytes += 2 * bytesIdx;
    ctx->d2hBytes += bytesOut + bytesIdx;
    uint32_t need = 0;
    for (uint32_t x = 0; x < nPairs; ++x)
      if (nEventsOut[x] == 0xFFFFFFFFu) need = std::max(need, out[x].nGapRuns);
    if (need == 0) {
      if (eventCap) {
        SD_CUDA(ctx, cudaMemcpyAsync(eventsOut, dEvents, uint64_t(nPairs) * eventCap * sizeof(sdgpu_event), cudaMemcpyDeviceToHost, ctx->stream));
        SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->d2hBytes += uint64_t(nPairs) * eventCap * sizeof(sdgpu_event);
      }
      for (uint32_t x = 0; x < nPairs; ++x) out[x].status &= ~SDGPU_ST_GAPLIST_TRUNCATED;  // (no gap list is returned here)
      return SDGPU_OK;
    }
    if (attempt > 0) return sd_fail(ctx, SDGPU_E_OVERFLOW, "gap list still truncated after resizing");
    gapCap = need;
  }
}

int sdgpu_set_pass_table(sdgpu_ctx* ctx, const sdgpu_iter_par* lines, uint32_t nLines) {
  if (!ctx || (nLines && !lines)) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nLines > SDGPU_MAX_PASS_LINES) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "more than 64 par-file lines in one pass table");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  if (!ctx->dLines) SD_CUDA(ctx, cudaMalloc(&ctx->dLines, SDGPU_MAX_PASS_LINES * sizeof(sdgpu_iter_par)));
  if (nLines) SD_CUDA(ctx, cudaMemcpyAsync(ctx->dLines, lines, size_t(nLines) * sizeof(sdgpu_iter_par), cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->nLines = nLines;
  return SDGPU_OK;
}

int sdgpu_align_pass(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs, uint32_t flags,
                     uint64_t* passMaskOut) {
  if (!ctx || (nPairs && (!refIdx || !readIdx || !passMaskOut))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (refIdx[x] >= nSeqs || readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[refIdx[x] + 1] - ctx->hOffsets[refIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  const uint64_t bytesIdx = uint64_t(nPairs) * 4, bytesOut = uint64_t(nPairs) * 8;
  int rc = sd_ensure_stage(ctx, bytesOut + 2 * bytesIdx + 256);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  uint64_t* dPass = reinterpret_cast<uint64_t*>(base);
  uint32_t* dRef = reinterpret_cast<uint32_t*>(base + bytesOut);
  uint32_t* dRead = dRef + nPairs;
  SD_CUDA(ctx, cudaMemcpyAsync(dRef, refIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dRead, readIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  rc = sd_launch_align(ctx, dRef, dRead, nPairs, flags, nullptr, nullptr, 0, maxA, maxB, dPass);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(passMaskOut, dPass, bytesOut, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += bytesOut;
  return SDGPU_OK;
}

int sdgpu_chimera_probe_pairs(sdgpu_ctx* ctx, const uint32_t* parentIdx, const uint32_t* candIdx, uint32_t nPairs,
                              uint32_t flags, const sdgpu_iter_par* chiOverlap, int keepLowQualityMismatches,
                              sdgpu_chimera_probe* out) {
  if (!ctx || !chiOverlap || (nPairs && (!parentIdx || !candIdx || !out))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  if (nPairs == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (parentIdx[x] >= nSeqs || candIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[parentIdx[x] + 1] - ctx->hOffsets[parentIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[candIdx[x] + 1] - ctx->hOffsets[candIdx[x]]));
  }
  const uint64_t bytesIdx = uint64_t(nPairs) * 4, bytesOut = uint64_t(nPairs) * sizeof(sdgpu_chimera_probe);
  int rc = sd_ensure_stage(ctx, bytesOut + 2 * bytesIdx + 256);
  if (rc) return rc;
  uint8_t* base = reinterpret_cast<uint8_t*>(ctx->dStage);
  SdChimeraReq req;
  req.dOut = reinterpret_cast<sdgpu_chimera_probe*>(base);
  req.overlap = *chiOverlap;
  req.keepLq = keepLowQualityMismatches ? 1u : 0u;
  uint32_t* dRef = reinterpret_cast<uint32_t*>(base + bytesOut);
  uint32_t* dRead = dRef + nPairs;
  SD_CUDA(ctx, cudaMemcpyAsync(dRef, parentIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaMemcpyAsync(dRead, candIdx, bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  rc = sd_launch_align(ctx, dRef, dRead, nPairs, flags, nullptr, nullptr, 0, maxA, maxB, nullptr, &req);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(out, req.dOut, bytesOut, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += bytesOut;
  return SDGPU_OK;
}

int sdgpu_align_pass_submit(sdgpu_ctx* ctx, int slot, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs,
                            uint32_t flags) {
  if (!ctx || slot < 0 || slot > 1 || !nPairs || !refIdx || !readIdx) return sd_fail(ctx, SDGPU_E_ARG, "bad argument");
  sdgpu_ctx::PassSlot& S = ctx->passSlots[slot];
  if (S.busy) return sd_fail(ctx, SDGPU_E_STATE, "slot still in flight: call sdgpu_align_pass_wait first");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nPairs; ++x) {
    if (refIdx[x] >= nSeqs || readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "pair index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[refIdx[x] + 1] - ctx->hOffsets[refIdx[x]]));
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  if (nPairs > S.capPairs) {
    if (S.hPinned) cudaFreeHost(S.hPinned);
    if (S.dBuf) cudaFree(S.dBuf);
    S.hPinned = nullptr;
    S.dBuf = nullptr;
    S.capPairs = 0;
    const uint64_t cap = std::max<uint64_t>(nPairs, 1u << 16);
    SD_CUDA(ctx, cudaMallocHost(&S.hPinned, cap * 16));
    SD_CUDA(ctx, cudaMalloc(&S.dBuf, cap * 16));
    S.capPairs = cap;
  }
  if (!S.done) SD_CUDA(ctx, cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
  const uint64_t bytesIdx = uint64_t(nPairs) * 4;
  std::memcpy(S.hPinned, refIdx, bytesIdx);
  std::memcpy(S.hPinned + bytesIdx, readIdx, bytesIdx);
  uint32_t* dRef = reinterpret_cast<uint32_t*>(S.dBuf);
  uint32_t* dRead = dRef + nPairs;
  uint64_t* dPass = reinterpret_cast<uint64_t*>(S.dBuf + S.capPairs * 8);
  SD_CUDA(ctx, cudaMemcpyAsync(dRef, S.hPinned, 2 * bytesIdx, cudaMemcpyHostToDevice, ctx->stream));
  int rc = sd_launch_align(ctx, dRef, dRead, nPairs, flags, nullptr, nullptr, 0, maxA, maxB, dPass);
  if (rc) return rc;
  SD_CUDA(ctx, cudaMemcpyAsync(S.hPinned + S.capPairs * 8, dPass, uint64_t(nPairs) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaEventRecord(S.done, ctx->stream));
  S.n = nPairs;
  S.busy = true;
  ctx->h2dBytes += 2 * bytesIdx;
  ctx->d2hBytes += uint64_t(nPairs) * 8;
  return SDGPU_OK;
}

int sdgpu_align_pass_wait(sdgpu_ctx* ctx, int slot, const uint64_t** masks, uint32_t* nPairs) {
  if (!ctx || slot < 0 || slot > 1 || !masks) return sd_fail(ctx, SDGPU_E_ARG, "bad argument");
  sdgpu_ctx::PassSlot& S = ctx->passSlots[slot];
  if (!S.busy) return sd_fail(ctx, SDGPU_E_STATE, "nothing submitted on this slot");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaEventSynchronize(S.done));
  S.busy = false;
  *masks = reinterpret_cast<const uint64_t*>(S.hPinned + S.capPairs * 8);
  if (nPairs) *nPairs = S.n;
  return SDGPU_OK;
}

int sdgpu_set_kernel(sdgpu_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2) return sd_fail(ctx, SDGPU_E_ARG, "mode must be 0 (auto), 1 (int32) or 2 (packed fp16)");
  ctx->forceKernel = mode;
  return SDGPU_OK;
}

int sdgpu_last_kernel(sdgpu_ctx* ctx) { return ctx ? (ctx->lastKernelWasH2 ? 2 : 1) : 0; }

int sdgpu_set_band(sdgpu_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return sd_fail(ctx, SDGPU_E_ARG, "sdgpu_set_band: mode must be 0 (automatic) or 1 (off)");
  ctx->bandMode = mode;
  return SDGPU_OK;
}

int sdgpu_set_screen(sdgpu_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return sd_fail(ctx, SDGPU_E_ARG, "sdgpu_set_screen: mode must be 0 (automatic) or 1 (off)");
  ctx->screenMode = mode;
  return SDGPU_OK;
}

// Every read of readIdx[] against every cluster of clusterIdx[]: pair lists are generated on the device in
// chunks of ~2 M pairs; each chunk's launches append (read, cluster, mask) records of the non-zero masks to one
// of two mapped pinned buffers, which the host empties while the next chunk runs.
int sdgpu_align_pass_grid(sdgpu_ctx* ctx, const uint32_t* clusterIdx, uint32_t nClusters, const uint32_t* readIdx,
                          uint32_t nReads, uint32_t flags, const sdgpu_grid_hit** hits, uint64_t* nHits) {
  if (!ctx || !hits || !nHits) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  *hits = nullptr;
  *nHits = 0;
  const int rc = sdgpu_align_pass_grid_stream(ctx, clusterIdx, nClusters, readIdx, nReads, flags, nullptr, nullptr);
  if (rc) return rc;
  *hits = ctx->gridHits.data();
  *nHits = ctx->gridHits.size();
  return SDGPU_OK;
}

int sdgpu_align_pass_grid_stream(sdgpu_ctx* ctx, const uint32_t* clusterIdx, uint32_t nClusters, const uint32_t* readIdx,
                                 uint32_t nReads, uint32_t flags, sdgpu_grid_chunk_fn onChunk, void* user) {
  if (!ctx || (nClusters && !clusterIdx) || (nReads && !readIdx)) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  ctx->gridHits.clear();
  if (nClusters == 0 || nReads == 0) return SDGPU_OK;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  for (sdgpu_ctx::GridSlot& G : ctx->gridSlots)  // a slot still marked busy belongs to a call that failed half-way: its records are not ours
    if (G.busy) {
      cudaEventSynchronize(G.done);
      G.busy = false;
    }
  if (ctx->nLines == 0) return sd_fail(ctx, SDGPU_E_STATE, "sdgpu_set_pass_table has not been called");
  const uint32_t nSeqs = uint32_t(ctx->hOffsets.size() - 1);
  uint32_t maxA = 0, maxB = 0;
  for (uint32_t x = 0; x < nClusters; ++x) {
    if (clusterIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "cluster index out of range");
    maxA = std::max<uint32_t>(maxA, uint32_t(ctx->hOffsets[clusterIdx[x] + 1] - ctx->hOffsets[clusterIdx[x]]));
  }
  for (uint32_t x = 0; x < nReads; ++x) {
    if (readIdx[x] >= nSeqs) return sd_fail(ctx, SDGPU_E_ARG, "read index out of range");
    maxB = std::max<uint32_t>(maxB, uint32_t(ctx->hOffsets[readIdx[x] + 1] - ctx->hOffsets[readIdx[x]]));
  }
  const uint64_t nIdx = uint64_t(nClusters) + nReads;
  int rc = sd_ensure_pinned(ctx, nIdx * 4);
  if (rc) return rc;
  if (nIdx > ctx->gridIdxCap) {
    if (ctx->dGridIdx) {
      SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      cudaFree(ctx->dGridIdx);
    }
    ctx->dGridIdx = nullptr;
    ctx->gridIdxCap = 0;
    const uint64_t cap = std::max<uint64_t>(nIdx * 2, 1u << 16);
    SD_CUDA(ctx, cudaMalloc(&ctx->dGridIdx, cap * 4));
    ctx->gridIdxCap = cap;
  }
  uint32_t* hIdx = reinterpret_cast<uint32_t*>(ctx->hPinned);
  std::memcpy(hIdx, clusterIdx, size_t(nClusters) * 4);
  std::memcpy(hIdx + nClusters, readIdx, size_t(nReads) * 4);
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dGridIdx, hIdx, nIdx * 4, cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2dBytes += nIdx * 4;
  uint64_t chunkPairs = 1ull << 21;
  if (const char* e = getenv("SDGPU_GRID_CHUNK_PAIRS")) {  // test hook: many small chunks
    const long long v = atoll(e);
    if (v > 0) chunkPairs = uint64_t(v);
  }
  const uint32_t readsPerChunk = uint32_t(std::max<uint64_t>(1, std::min<uint64_t>(nReads, chunkPairs / nClusters)));
  if (uint64_t(readsPerChunk) * nClusters >= (1ull << 32)) return sd_fail(ctx, SDGPU_E_UNSUPPORTED, "more than 2^32 clusters in one grid call");
  auto drain = [&](sdgpu_ctx::GridSlot& G) -> int {
    SD_CUDA(ctx, cudaEventSynchronize(G.done));
    G.busy = false;
    const uint32_t n = *G.hCount;
    ctx->d2hBytes += uint64_t(n) * sizeof(sdgpu_grid_hit) + 4;
    if (onChunk) {  // the caller consumes the chunk while the GPU works on the next one
      onChunk(user, G.hHits, n, G.readsDone);
    } else {
      ctx->gridHits.insert(ctx->gridHits.end(), G.hHits, G.hHits + n);
    }
    return SDGPU_OK;
  };
  uint32_t k = 0;
  for (uint32_t r0 = 0; r0 < nReads; r0 += readsPerChunk, ++k) {
    const uint32_t nr = std::min(readsPerChunk, nReads - r0);
    const uint64_t pairs = uint64_t(nr) * nClusters;
    const int slot = int(k & 1);
    sdgpu_ctx::GridSlot& G = ctx->gridSlots[slot];
    if (G.busy && (rc = drain(G))) return rc;
    if (pairs > G.cap) {
      if (G.hHits) cudaFreeHost(G.hHits);
      G.hHits = G.dHits = nullptr;
      G.cap = 0;
      const uint64_t cap = std::max<uint64_t>(pairs, 1u << 14);
      SD_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&G.hHits), cap * sizeof(sdgpu_grid_hit), cudaHostAllocMapped));
      SD_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&G.dHits), G.hHits, 0));
      G.cap = cap;
    }
    if (!G.hCount) SD_CUDA(ctx, cudaHostAlloc(reinterpret_cast<void**>(&G.hCount), 64, cudaHostAllocDefault));
    if (!G.done) SD_CUDA(ctx, cudaEventCreateWithFlags(&G.done, cudaEventDisableTiming));
    SD_CUDA(ctx, cudaMemsetAsync(ctx->dWork + SD_W_HITS + slot, 0, sizeof(uint32_t), ctx->stream));
    SdAlignReq req;
    req.src.clus = ctx->dGridIdx;
    req.src.reads = ctx->dGridIdx + nClusters + r0;
    req.src.nClus = nClusters;
    req.src.readBase = r0;
    req.nPairs = uint32_t(pairs);
    req.flags = flags;
    req.maxLenA = maxA;
    req.maxLenB = maxB;
    req.hits.recs = G.dHits;
    req.hits.count = ctx->dWork + SD_W_HITS + slot;
    req.hits.cap = uint32_t(std::min<uint64_t>(G.cap, 0xffffffffu));
    if ((rc = sd_launch_align(ctx, req))) return rc;
    SD_CUDA(ctx, cudaMemcpyAsync(G.hCount, ctx->dWork + SD_W_HITS + slot, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SD_CUDA(ctx, cudaEventRecord(G.done, ctx->stream));
    G.busy = true;
    G.readsDone = r0 + nr;
  }
  // the older slot first
  for (uint32_t x = 0; x < 2; ++x) {
    sdgpu_ctx::GridSlot& G = ctx->gridSlots[(k + x) & 1];
    if (G.busy && (rc = drain(G))) return rc;
  }
  return SDGPU_OK;
}

int sdgpu_profile_enable(sdgpu_ctx* ctx, int on) {
  if (!ctx) return SDGPU_E_ARG;
  ctx->timeKernels = on != 0;
  return SDGPU_OK;
}

int sdgpu_profile_read(sdgpu_ctx* ctx, sdgpu_profile* out, int reset) {
  if (!ctx || !out) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  for (size_t x = 0; x + 1 < ctx->evUsed; x += 2) {
    float ms = 0;
    SD_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->evPool[x], ctx->evPool[x + 1]));
    ctx->alignMs += ms;
    if (ctx->evBand[x / 2] == 1) ctx->bandMs += ms;
    if (ctx->evBand[x / 2] >= 2) ctx->screenMs += ms;
    if (ctx->evBand[x / 2] == 2) ctx->screenScanMs += ms;
  }
  ctx->evUsed = 0;
  uint64_t cells = 0, band[2] = {0, 0}, scr[2] = {0, 0};
  uint32_t scrCert = 0;
  SD_CUDA(ctx, cudaMemcpy(&cells, ctx->dWork + SD_W_CELLS, 8, cudaMemcpyDeviceToHost));
  SD_CUDA(ctx, cudaMemcpy(band, ctx->dWork + SD_W_BAND_CELLS, 16, cudaMemcpyDeviceToHost));
  SD_CUDA(ctx, cudaMemcpy(scr, ctx->dWork + SD_W_SCR_CELLS, 16, cudaMemcpyDeviceToHost));
  SD_CUDA(ctx, cudaMemcpy(&scrCert, ctx->dWork + SD_W_SCR_CERT, 4, cudaMemcpyDeviceToHost));
  // cells actually evaluated = all la*lb of the finished pairs, minus those the screen and the banded pass
  // finished, plus the band cells the two really evaluated (certified or not)
  out->bandCells = band[0] - ctx->bandCells;
  out->bandCertifiedCells = band[1] - ctx->bandCertCells;
  out->screenCells = scr[0] - ctx->screenCells;
  out->screenCertCells = scr[1] - ctx->screenCertCells;
  out->screenCertified = uint64_t(scrCert) - ctx->screenCert;  // (a 32-bit device counter: read before it wraps)
  out->screenPairs = ctx->screenPairs;
  out->screenKernelMs = ctx->screenMs;
  out->screenScanMs = ctx->screenScanMs;
  out->evaluatedCells = (cells - ctx->cells) - out->bandCertifiedCells - out->screenCertCells + out->bandCells + out->screenCells;
  out->alignKernelMs = ctx->alignMs;
  out->bandKernelMs = ctx->bandMs;
  out->alignKernelLaunches = ctx->alignLaunches;
  out->kernelLaunches = ctx->launches;
  out->cellUpdates = cells - ctx->cells;
  out->h2dBytes = ctx->h2dBytes;
  out->d2hBytes = ctx->d2hBytes;
  if (reset) {
    ctx->alignMs = 0;
    ctx->bandMs = 0;
    ctx->alignLaunches = 0;
    ctx->launches = 0;
    ctx->cells = cells;
    ctx->bandCells = band[0];
    ctx->bandCertCells = band[1];
    ctx->screenCells = scr[0];
    ctx->screenCertCells = scr[1];
    ctx->screenCert = scrCert;
    ctx->screenPairs = 0;
    ctx->screenMs = 0;
    ctx->screenScanMs = 0;
    ctx->h2dBytes = ctx->d2hBytes = 0;
  }
  return SDGPU_OK;
}

int sdgpu_align_profile_dev(sdgpu_ctx* ctx, const uint32_t* dRefIdx, const uint32_t* dReadIdx, uint32_t nPairs,
                            uint32_t flags, sdgpu_comparison* dOut) {
  if (!ctx || (nPairs && (!dRefIdx || !dReadIdx || !dOut))) return sd_fail(ctx, SDGPU_E_ARG, "null argument");
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  return sd_launch_align(ctx, dRefIdx, dReadIdx, nPairs, flags, dOut, nullptr, 0, ctx->maxLen, ctx->maxLen);
}

int sdgpu_dev_alloc(sdgpu_ctx* ctx, uint64_t bytes, void** dptr) {
  if (!ctx || !dptr) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaMalloc(dptr, bytes));
  return SDGPU_OK;
}

int sdgpu_dev_free(sdgpu_ctx* ctx, void* dptr) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  SD_CUDA(ctx, cudaFree(dptr));
  return SDGPU_OK;
}

int sdgpu_memcpy_h2d(sdgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_memcpy_d2h(sdgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_sync(sdgpu_ctx* ctx) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return SDGPU_OK;
}

int sdgpu_timer_start(sdgpu_ctx* ctx) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return SDGPU_OK;
}

int sdgpu_timer_stop(sdgpu_ctx* ctx, float* msOut) {
  if (!ctx || !msOut) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  SD_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
  SD_CUDA(ctx, cudaEventElapsedTime(msOut, ctx->ev0, ctx->ev1));
  return SDGPU_OK;
}

int sdgpu_counters(sdgpu_ctx* ctx, uint64_t* kernelLaunches, uint64_t* cellUpdates) {
  if (!ctx) return SDGPU_E_ARG;
  SD_CUDA(ctx, cudaSetDevice(ctx->device));
  SD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  uint64_t cells = 0;
  SD_CUDA(ctx, cudaMemcpy(&cells, ctx->dWork + SD_W_CELLS, 8, cudaMemcpyDeviceToHost));
  if (kernelLaunches) *kernelLaunches = ctx->launches;
  if (cellUpdates) *cellUpdates = cells;
  return SDGPU_OK;
}

}  // extern "C"
