// sdgpu_internal.cuh — context, device-side views and shared helpers of the sdgpu library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sdgpu.h"

#define SDGPU_MAX_SYMBOLS 16  // distinct base characters per ctx (dynamic alphabet)
#define SDGPU_GAP_CODE 0xFF   // '-' in the device-side gapped strings

// Everything a kernel needs to know about scoring; passed by value (fits the 4 KB param space).
struct DevScoring {
  int go, ge;      // interior
  int goLQ, geLQ;  // left gap in query  (first column, ref bases consumed)
  int goRQ, geRQ;  // right gap in query (last column)
  int goLR, geLR;  // left gap in ref    (first row)
  int goRR, geRR;  // right gap in ref   (last row)
  int primaryQual, secondaryQual, window;
  int countEndGaps, weighHomopolymers;
  int16_t score[SDGPU_MAX_SYMBOLS * SDGPU_MAX_SYMBOLS];  // [refCode*16 + readCode]
};

struct DevSeqs {
  const uint8_t* codes;  // 1 byte per base, value < 16
  const uint8_t* quals;
  const uint64_t* offsets;  // n+1
  const double* cnts;
  uint32_t n;
};

// Sorted positional k-mer table (KmerMaps): key = pos << 48 | 4-bit-coded k-mer.
struct DevKmerIndex {
  const uint64_t* keys;
  const uint32_t* counts;
  uint32_t n;
  uint32_t kLen, runCutOff, byPos;
};

struct sdgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int numSMs = 0;
  sdgpu_params params;
  DevScoring scoring;
  // dynamic alphabet: byte value -> code
  int16_t codeOf[256];
  uint8_t charOf[SDGPU_MAX_SYMBOLS];
  int nSym = 0;
  int usedSym = 0;  // highest code present in the store + 1 (A,C,G,T,N are pre-assigned 0..4)
  // sequence store
  uint8_t* dCodes = nullptr;
  uint8_t* dQuals = nullptr;
  uint64_t* dOffsets = nullptr;
  double* dCnts = nullptr;
  uint64_t capBases = 0;
  uint32_t capSeqs = 0;
  std::vector<uint64_t> hOffsets{0};
  uint32_t maxLen = 0;
  // k-mer index
  uint64_t* dIdxKeys = nullptr;
  uint32_t* dIdxCnts = nullptr;
  uint32_t nIdx = 0, idxK = 0, idxRunCutOff = 0, idxByPos = 0;
  bool haveIndex = false;
  uint32_t idxCapRuns = 0;
  void* dIdxWork = nullptr;  // grow-only arena for the index build (keys, weights, CUB temp)
  uint64_t idxWorkBytes = 0;
  // dense k-mer profiles
  uint16_t* dProfiles = nullptr;
  uint32_t profK = 0, profSize = 0, profSeqs = 0;
  uint64_t profCapBytes = 0;
  // scratch for the alignment kernel (pointer slabs), pair staging, work counter
  uint8_t* dScratch = nullptr;
  uint64_t scratchBytes = 0;
  uint32_t* dWork = nullptr;
  void* dStage = nullptr;
  uint64_t stageBytes = 0;
  void* hPinned = nullptr;
  uint64_t pinnedBytes = 0;
  // double-buffered asynchronous sdgpu_align_pass_submit / _wait
  struct PassSlot {
    uint8_t* hPinned = nullptr;  // [refIdx | readIdx | masks]
    uint8_t* dBuf = nullptr;
    uint64_t capPairs = 0;
    cudaEvent_t done = nullptr;
    uint32_t n = 0;
    bool busy = false;
  } passSlots[2];
  // par-file lines for on-device passErrorProfile
  sdgpu_iter_par* dLines = nullptr;
  uint32_t nLines = 0;
  // profiling bookkeeping (sdgpu_profile_*): per-launch events of the alignment kernel, PCIe bytes
  bool timeKernels = false;
  std::vector<cudaEvent_t> evPool;
  size_t evUsed = 0;
  std::vector<uint8_t> evBand;  // per event pair: 1 = banded pass
  double alignMs = 0, bandMs = 0;
  uint64_t alignLaunches = 0;
  uint64_t h2dBytes = 0, d2hBytes = 0;
  // alignment kernel selection: 0 = automatic, 1 = int32 only, 2 = packed fp16 only (tests)
  int forceKernel = 0;
  bool lastKernelWasH2 = false;
  // banded first pass (align_kernel.cu): 0 = automatic, 1 = off; redo list + its async read-back
  int bandMode = 0;
  uint32_t* dRedo = nullptr;
  uint32_t redoCap = 0;
  uint32_t* hBandStat = nullptr;  // pinned: [0] = redo count of the last banded launch, [1] = its pair count
  uint32_t bandSkipped = 0;       // launches since banding was last tried while it was not paying off
  uint64_t bandCells = 0, bandCertCells = 0, bandRedo = 0, bandPairs = 0;  // read-back baselines for profile_read
  // bookkeeping
  uint64_t launches = 0, cells = 0;
  std::string err;
};

#define SD_CUDA(ctx, call)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                      \
      return SDGPU_E_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

int sd_fail(sdgpu_ctx* ctx, int code, const std::string& msg);
int sd_ensure_scratch(sdgpu_ctx* ctx, uint64_t bytes);
int sd_ensure_stage(sdgpu_ctx* ctx, uint64_t bytes);
int sd_ensure_pinned(sdgpu_ctx* ctx, uint64_t bytes);
DevSeqs sd_dev_seqs(const sdgpu_ctx* ctx);
DevKmerIndex sd_dev_index(const sdgpu_ctx* ctx);

// align_kernel.cu
struct SdChimeraReq {  // sdgpu_chimera_probe_pairs: per-pair probe output + chiOpts_
  sdgpu_chimera_probe* dOut;
  sdgpu_iter_par overlap;
  uint32_t keepLq;
};
int sd_launch_align(sdgpu_ctx* ctx, const uint32_t* dRefIdx, const uint32_t* dReadIdx, uint32_t nPairs,
                    uint32_t flags, sdgpu_comparison* dOut, sdgpu_gap* dGaps, uint32_t gapCap,
                    uint32_t maxLenA, uint32_t maxLenB, uint64_t* dPassOut = nullptr,
                    const SdChimeraReq* chi = nullptr);
// kmer_kernels.cu
int sd_build_kmer_index(sdgpu_ctx* ctx, const uint32_t* hSeqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                        int byPos, int expandPos, uint32_t expandSize);
int sd_kmer_index_lookup(sdgpu_ctx* ctx, const uint8_t* kmers, const uint32_t* positions, uint32_t n,
                         uint32_t* countsOut);
int sd_build_kmer_profiles(sdgpu_ctx* ctx, uint32_t kLen);
int sd_kmer_compare_pairs(sdgpu_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t nPairs,
                          uint32_t* sharedOut, double* fracOut);
int sd_kmer_compare_all_dev(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads, const uint32_t* dClusters,
                            uint32_t nClusters, uint32_t* dShared, double* dFrac);
int sd_kmer_compare_all(sdgpu_ctx* ctx, const uint32_t* reads, uint32_t nReads, const uint32_t* clusters,
                        uint32_t nClusters, uint32_t* sharedOut, double* fracOut);
