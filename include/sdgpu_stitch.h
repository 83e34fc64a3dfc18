/*
 * sdgpu_stitch.h — C ABI of paired-end stitching, the step in front of qluster in
 * `SeekDeep extractorPairedEnd`.
 *
 * Replaces PairedReadProcessor::processPairedEnd (src/SeekDeep/objects/IlluminaUtils/PairedReadProcessor.cpp:287-664)
 * for a batch of read pairs:
 *   alignerObj.alignRegGlobalNoInternalGaps(r1, r2)  +  alignerObj.profileAlignment(r1, r2, false, true, false)   (:333-336)
 *   the accept rule on eventBasedIdentityHq_ / basesInAln_ / hq / lq mismatches                                  (:448-452)
 *   the low-quality-window trim and second attempt for pairs that fail it                                        (:358-412)
 *   the overhang cases and the per-column consensus (addToConsensus, :43-78)                                     (:453-628)
 * The aligner is the one extractorPairedEnd builds for it (extractorPairedEnd.cpp:338-348: gap 10,1, every end gap
 * 0,0, match 2 / mismatch -2, qualThresWindow_ 0, countEndGaps false); mates arrive reverse-complemented like
 * PairedRead's mateSeqBase_ (mateRComplemented_).
 *
 * The alignment has no internal gaps, so it is a shift: every (r1, r2) diagonal is scored on the device
 * (substitution scores of the overlapping columns minus the end-gap penalties of the two overhangs), the best one
 * is profiled there, and 48 bytes per pair return.  Trimming, the overhang cases and the consensus stay on the host.
 */
#ifndef SDGPU_STITCH_H_
#define SDGPU_STITCH_H_

#include "sdgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* aligner::alignRegGlobalNoInternalGaps + profileAlignment of one pair.  `shift` is the position of r2's first base
 * in r1's coordinates (negative: r2 starts before r1), which is the whole alignment:
 *   alignObjectA_ = '-' x max(0, -shift) + r1 + '-' x max(0, shift + len2 - len1)
 *   alignObjectB_ = '-' x max(0,  shift) + r2 + '-' x max(0, len1 - shift - len2)            */
typedef struct sdgpu_overlap {
  int32_t alnScore;
  int32_t shift;
  uint32_t basesInAln;        /* overlapping columns (distances_.basesInAln_) */
  uint32_t hqMismatches, lqMismatches, lowKmerMismatches;
  uint32_t highQualityMatches, lowQualityMatches;
  double eventBasedIdentity;
  double eventBasedIdentityHq;
} sdgpu_overlap;

/* Per pair i: alignRegGlobalNoInternalGaps(seq[r1Idx[i]], seq[r2Idx[i]]) + profileAlignment(.., checkKmer, usingQuality,
 * matchQuality) on the ctx's store.  Among equally good shifts the smallest one wins.  countEndGaps must be off. */
int sdgpu_align_no_internal_gaps(sdgpu_ctx* ctx, const uint32_t* r1Idx, const uint32_t* r2Idx, uint32_t nPairs, uint32_t flags,
                                 sdgpu_overlap* out);

/* PairedReadProcessor::ProcessParams (PairedReadProcessor.hpp:104-133) */
typedef struct sdst_params {
  uint32_t minOverlap;          /* 10 */
  uint32_t hardMismatchCutOff;  /* 20 */
  uint32_t lqMismatchCutOff;    /* 10 */
  uint32_t hqMismatchCutOff;    /* 10 */
  uint32_t r1Trim, r2Trim;      /* 0 */
  uint32_t trimLowQualWindows;  /* 1 (trimLowQaulWindows_) */
  uint32_t windowSize;          /* 5 (qualWindowPar_) */
  uint32_t windowStep;          /* 1 */
  uint32_t reserved;
  double errorAllowed;          /* 0.01 */
  double avgQualCutOff;         /* 25 */
  double percentAfterTrimCutOff;/* 0.70 */
} sdst_params;

/* ReadPairOverLapStatus (PairedReadProcessor.hpp) */
enum {
  SDST_NOOVERLAP = 0,
  SDST_R1BEGINSINR2 = 1,
  SDST_R1ENDSINR2 = 2,
  SDST_PERFECTOVERLAP = 3,
  SDST_R1ALLINR2 = 4,
  SDST_R2ALLINR1 = 5
};

typedef struct sdst_result sdst_result;

/* Stitches n pairs (reads packed like sdgpu_upload_seqs; r2 already reverse-complemented).  The ctx's store is
 * replaced.  Results: status per pair, the combined sequences of the stitched pairs (packed, in pair order; a pair
 * that did not stitch has an empty entry), and ProcessedResultsCounts. */
int sdst_process_pairs(sdgpu_ctx* ctx, const sdst_params* pars, const uint8_t* r1Bases, const uint8_t* r1Quals,
                       const uint64_t* r1Offsets, const uint8_t* r2Bases, const uint8_t* r2Quals, const uint64_t* r2Offsets,
                       uint32_t n, sdst_result** out);
void sdst_free(sdst_result* r);
const char* sdst_last_error(void);
uint64_t sdst_total_len(const sdst_result* r);
/* status[n], overlaps[n] (the alignment each decision was taken on), offsets[n + 1], bases / quals of sdst_total_len bytes;
 * any pointer may be NULL */
int sdst_fetch(const sdst_result* r, uint32_t* status, sdgpu_overlap* overlaps, uint8_t* bases, uint8_t* quals, uint64_t* offsets);
/* counts[6]: pairs per status, indexed by SDST_*; counts[6] = pairs whose second (trimmed) attempt was made */
int sdst_counts(const sdst_result* r, uint64_t counts[7]);

#ifdef __cplusplus
}
#endif
#endif /* SDGPU_STITCH_H_ */
