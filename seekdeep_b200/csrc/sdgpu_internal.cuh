// sdgpu_internal.cuh — context, device-side views and shared helpers of the sdgpu library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sdgpu.h"
#include "../../include/sdgpu_policy.h"

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

#define SD_FLAG_QUAL_OK 1u   // seqInfo::checkQual(pos, qScorePars_) holds for this base
#define SD_FLAG_LOW_KMER 2u  // the k-mer centred on this base is low-frequency in the current KmerMaps

struct DevSeqs {
  const uint8_t* codes;  // 1 byte per base, value < 16
  const uint8_t* quals;
  const uint8_t* flags;  // 1 byte per base: SD_FLAG_* (what an errorProfile needs to know about a mismatching base)
  const uint64_t* offsets;  // n+1
  const double* cnts;
  uint32_t n;
};

// Where a launch's (ref, read) pairs come from: two explicit index lists, or the (read x cluster)
// grid of the collapse loop (pair p = r * nClus + c -> ref clus[c], read reads[r]; pairs whose two
// store indices coincide are skipped: a cluster is never compared with itself).
struct DevPairSrc {
  const uint32_t* refIdx;  // list mode
  const uint32_t* readIdx;
  const uint32_t* clus;    // grid mode when non-null
  const uint32_t* reads;
  uint32_t nClus;
  uint32_t readBase;       // ordinal of reads[0] in the caller's read list (hit records carry ordinals)
};
__device__ __forceinline__ void sd_pair_of(const DevPairSrc& s, uint32_t p, uint32_t& ia, uint32_t& ib) {
  if (s.clus) {
    const uint32_t r = p / s.nClus;
    ia = s.clus[p - r * s.nClus];
    ib = s.reads[r];
  } else {
    ia = s.refIdx[p];
    ib = s.readIdx[p];
  }
}
// compacted output of a grid launch: one record per pair whose verdict mask is not zero
struct DevHits {
  sdgpu_grid_hit* recs;  // may live in mapped pinned host memory
  uint32_t* count;
  uint32_t cap;
};
__device__ __forceinline__ void sd_emit_hit(const DevPairSrc& s, const DevHits& h, uint32_t p, uint64_t mask) {
  const uint32_t slot = atomicAdd(h.count, 1u);
  if (slot < h.cap) {
    const uint32_t r = p / s.nClus;
    h.recs[slot] = sdgpu_grid_hit{s.readBase + r, p - r * s.nClus, mask};
  }
}

// Sorted positional k-mer table (KmerMaps): key = pos << 48 | 4-bit-coded k-mer.
struct DevKmerIndex {
  const uint64_t* keys;
  const uint32_t* counts;
  uint32_t n;
  uint32_t kLen, runCutOff, byPos;
  const uint32_t* posStart;  // by-position tables: first key of every position (nPos + 1 entries), else NULL
  uint32_t nPos;
};

// KmerMaps count of the k-mer centred on base p (clamped to the sequence); false: the sequence has no such k-mer
__device__ __forceinline__ bool sd_kmer_count_at(const DevKmerIndex& idx, const uint8_t* codes, int len, int p, uint32_t& cnt) {
  const int k = int(idx.kLen);
  cnt = 0;
  if (k == 0 || len < k) return false;
  int start = p >= k / 2 ? p - k / 2 : 0;
  if (start + k > len) start = len - k;
  uint64_t key = 0;
  for (int x = 0; x < k; ++x) key = (key << 4) | codes[start + x];
  uint32_t lo = 0, hi = idx.n;
  if (idx.byPos) {
    key |= uint64_t(start) << 48;
    if (idx.posStart) {  // the keys of one position are contiguous
      lo = idx.posStart[min(uint32_t(start), idx.nPos)];
      if (uint32_t(start) < idx.nPos) hi = idx.posStart[start + 1];
    }
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(idx.keys + mid) < key) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  cnt = (lo < idx.n && __ldg(idx.keys + lo) == key) ? __ldg(idx.counts + lo) : 0u;
  return true;
}

// KmerMaps::isKmerLowFrequency for the k-mer centred on base p: count <= runCutOff
__device__ __forceinline__ bool sd_low_freq(const DevKmerIndex& idx, const uint8_t* codes, int len, int p) {
  uint32_t cnt;
  if (!sd_kmer_count_at(idx, codes, len, p, cnt)) return false;
  return cnt <= idx.runCutOff;
}

// screen_kernel.cu work item
struct SdScreenItem {  // a pair whose gapless alignment still has to be proved the unique optimum
  uint32_t pair;
  int32_t sdiagK;  // K x score of the gapless alignment
  uint64_t mask;   // its verdict mask
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
  uint8_t* dFlags = nullptr;  // SD_FLAG_* per base
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
  uint32_t* dIdxPosStart = nullptr;  // by-position index: first key of every position
  uint32_t idxPosCap = 0, idxNPos = 0;
  void* dIdxWork = nullptr;  // grow-only arena for the index build (keys, weights, CUB temp)
  uint64_t idxWorkBytes = 0;
  // dense k-mer profiles
  uint16_t* dProfiles = nullptr;
  uint32_t profK = 0, profSize = 0, profSeqs = 0;
  uint64_t profCapBytes = 0;
  uint16_t* dProfilesRev = nullptr;  // kmerInfo(seq, k, true)'s reverse-complement maps, derived on first use
  uint64_t profRevCapBytes = 0;
  bool profRevValid = false;
  uint32_t* dRcIndex = nullptr;      // k-mer id -> id of its reverse complement
  uint32_t rcIndexCap = 0;
  // sorted k-mer lists (kmer_list_kernels.cu): one uint64 key per base slot of the store
  uint64_t* dKmerLists = nullptr;
  uint64_t listCapKeys = 0;
  uint32_t listK = 0, listSeqs = 0;
  uint8_t* dCharOf = nullptr;  // code -> character, for the event lists (events_kernel.cu)
  std::vector<uint32_t> gapRecs;  // packed gap lists of the last sdgpu_align_gap_lists call
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
  cudaEvent_t bandStatEv = nullptr;  // completes when hBandStat holds the last banded launch's numbers
  uint32_t* dRedo2 = nullptr;     // redo list of the banded pass when it runs behind the screen pass
  uint32_t redo2Cap = 0;
  // screen pass (screen_kernel.cu): 0 = automatic, 1 = off; band lists
  int screenMode = 0;
  SdScreenItem* dScreenItems = nullptr;
  uint64_t screenItemsCap = 0;  // pairs per list
  double screenMs = 0, screenScanMs = 0;
  uint64_t screenPairs = 0;
  uint64_t screenCells = 0, screenCertCells = 0, screenCert = 0;  // read-back baselines for profile_read
  // sdgpu_align_pass_grid: device copies of the two index lists, two zero-copy hit slots, the result vector
  uint32_t* dGridIdx = nullptr;
  uint64_t gridIdxCap = 0;
  struct GridSlot {
    sdgpu_grid_hit* hHits = nullptr;  // mapped pinned memory, written by the kernels
    sdgpu_grid_hit* dHits = nullptr;  // device alias
    uint32_t* hCount = nullptr;       // pinned
    uint64_t cap = 0;
    cudaEvent_t done = nullptr;
    bool busy = false;
    uint32_t readsDone = 0;           // reads of the caller's list finished once this slot's chunk is
  } gridSlots[2];
  std::vector<sdgpu_grid_hit> gridHits;
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
struct SdAlignReq {  // one alignment launch; every pointer is a device pointer
  DevPairSrc src{};
  uint32_t nPairs = 0;
  uint32_t flags = 0;
  sdgpu_comparison* dOut = nullptr;
  sdgpu_gap* dGaps = nullptr;
  uint32_t gapCap = 0;
  uint32_t maxLenA = 0, maxLenB = 0;
  uint64_t* dPassOut = nullptr;
  const SdChimeraReq* chi = nullptr;
  DevHits hits{};  // grid launches: compacted non-zero verdict masks
};
int sd_launch_align(sdgpu_ctx* ctx, const SdAlignReq& req);
inline int sd_launch_align(sdgpu_ctx* ctx, const uint32_t* dRefIdx, const uint32_t* dReadIdx, uint32_t nPairs,
                           uint32_t flags, sdgpu_comparison* dOut, sdgpu_gap* dGaps, uint32_t gapCap,
                           uint32_t maxLenA, uint32_t maxLenB, uint64_t* dPassOut = nullptr,
                           const SdChimeraReq* chi = nullptr) {
  SdAlignReq r;
  r.src.refIdx = dRefIdx;
  r.src.readIdx = dReadIdx;
  r.nPairs = nPairs;
  r.flags = flags;
  r.dOut = dOut;
  r.dGaps = dGaps;
  r.gapCap = gapCap;
  r.maxLenA = maxLenA;
  r.maxLenB = maxLenB;
  r.dPassOut = dPassOut;
  r.chi = chi;
  return sd_launch_align(ctx, r);
}

// screen_kernel.cu: the screen pass in front of the verdict-mask launches
struct SdScreenArgs {
  DevScoring sc;
  DevSeqs seqs;
  DevKmerIndex idx;
  DevPairSrc src;
  uint32_t nPairs;
  uint32_t flags;
  const sdgpu_iter_par* lines;
  uint32_t nLines;
  uint64_t* passOut;
  DevHits hits;
  SdScreenItem* items;  // three lists of nPairs entries: bands of 16 / 32 / 64 diagonals
  uint32_t* redo;       // pairs left to the bit-recording kernels
  uint32_t* work;       // ctx->dWork (SD_W_* words)
  int K;                // perturbation scale: a power of two above the move count of any path
  int maxSub;
  int selfScore;        // score of every stored symbol against itself when they all agree and it is >= 0, else -1
  int misScore;         // score of every pair of different stored symbols when they all agree and it is < 0, else 0
  long long bandSlack[3];  // a pair whose deficit maxSub*L - S_diag is below bandSlack[b] cannot be beaten outside band b
  int screenOn;         // 0: the scan only lists the non-self pairs of a grid launch
};
bool sd_screen_plan(const sdgpu_ctx* ctx, uint32_t maxLenA, uint32_t maxLenB, SdScreenArgs* out);
int sd_launch_screen(sdgpu_ctx* ctx, SdScreenArgs& args);

// words of ctx->dWork (uint32 indices; 64-bit counters sit on even indices)
enum {
  SD_W_NEXT = 0,           // work counter of the running alignment kernel
  SD_W_CELLS = 2,          // u64, cumulative: sum of lenA x lenB over finished pairs
  SD_W_REDO = 4,           // banded pass: pairs it could not certify
  SD_W_BAND_PAIRS = 5,     // banded pass: pairs it looked at
  SD_W_BAND_CELLS = 6,     // u64, cumulative
  SD_W_BAND_CERTCELLS = 8, // u64, cumulative
  SD_W_SCAN = 16,          // screen scan: work counter
  SD_W_BIN0 = 17,          // 17..19: entries of the three band lists, 20: entries of the screen's redo list
  SD_W_DP0 = 21,           // 21..23: work counters of the three score-only DP launches
  SD_W_SCR_CELLS = 24,     // u64, cumulative: band cells the score-only DP evaluated
  SD_W_SCR_CERTCELLS = 26, // u64, cumulative: lenA x lenB of the pairs the screen finished
  SD_W_SCR_CERT = 28,      // cumulative: pairs the screen finished
  SD_W_HITS = 30,          // 30, 31: hit counters of the two grid slots
  SD_W_WORDS = 64
};

// kmer_kernels.cu
int sd_build_kmer_index(sdgpu_ctx* ctx, const uint32_t* hSeqIdx, uint32_t nIdx, uint32_t kLen, uint32_t runCutOff,
                        int byPos, int expandPos, uint32_t expandSize);
// per-base flags of store sequences [firstSeq, firstSeq + nSeqs): checkQual outcome / low-frequency k-mer
int sd_refresh_qual_flags(sdgpu_ctx* ctx, uint32_t firstSeq, uint32_t nSeqs);
int sd_refresh_kmer_flags(sdgpu_ctx* ctx, uint32_t firstSeq, uint32_t nSeqs);
int sd_kmer_index_lookup(sdgpu_ctx* ctx, const uint8_t* kmers, const uint32_t* positions, uint32_t n,
                         uint32_t* countsOut);
int sd_build_kmer_profiles(sdgpu_ctx* ctx, uint32_t kLen);
int sd_kmer_compare_pairs(sdgpu_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t nPairs, int revCompB,
                          uint32_t* sharedOut, double* fracOut);
int sd_kmer_compare_all_dev(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads, const uint32_t* dClusters,
                            uint32_t nClusters, uint32_t* dShared, double* dFrac);
int sd_kmer_compare_all(sdgpu_ctx* ctx, const uint32_t* reads, uint32_t nReads, const uint32_t* clusters,
                        uint32_t nClusters, uint32_t* sharedOut, double* fracOut);
// events_kernel.cu: the per-event lists of a batch whose gap lists are on the device
int sd_launch_events(sdgpu_ctx* ctx, const uint32_t* dRef, const uint32_t* dRead, uint32_t nPairs, uint32_t flags,
                     const sdgpu_comparison* dCmp, const sdgpu_gap* dGaps, uint32_t gapCap, sdgpu_event* dEvents, uint32_t eventCap,
                     uint32_t* dNEvents);
int sd_launch_pack_gaps(sdgpu_ctx* ctx, const sdgpu_comparison* dCmp, const sdgpu_gap* dGaps, uint32_t gapCap, uint32_t nPairs,
                        uint32_t* dPacked, unsigned long long* dNWords);
// kmer_list_kernels.cu
int sd_build_kmer_lists(sdgpu_ctx* ctx, uint32_t kLen);
int sd_kmer_bin_masks(sdgpu_ctx* ctx, const uint32_t* seeds, uint32_t nSeeds, const uint32_t* reads, uint32_t nReads, double cutOff,
                      uint64_t* maskOut, uint32_t* sharedOut);
