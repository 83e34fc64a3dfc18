/*
 * sdgpu.h — C ABI of the B200-native qluster hot path (k-mer prefilter -> affine-gap global
 * alignment with traceback -> errorProfile/`comparison`).
 *
 * This is the drop-in boundary for SeekDeep's data-parallel hot path.  The reference has no
 * FFI: the seam is the C++ class API of njhseq's `aligner` / `kmerInfo` / `indexKmers`, as
 * called from the reference at (paths relative to the SeekDeep repo root):
 *
 *   aligner ctor (maxSize, gapInfo, scoring, kMaps, qScorePars, countEndGaps, weighHomopolymers)
 *       src/SeekDeepPrograms/SeekDeepProgram/clusterDown/clusterDown.cpp:420-426
 *   aligner::alignCacheGlobal(ref, read)                  clusterDown.cpp:694,739
 *                                                          src/SeekDeep/objects/ControlBenchmarking/ControlBencher.hpp:117
 *   aligner::profileAlignment(ref, read, checkKmer, usingQuality, doingMatchQuality)
 *                                                          clusterDown.cpp:696,741; ControlBencher.hpp:118
 *   comparison fields (alnScore_, hq/lq/lowKmer mismatches, 1/2/large indels, identities)
 *                                                          clusterDown.cpp:745-753
 *   indexKmers(clusters, kLen, runCutOff, kmersByPosition, expandKmerPos, expandKmerSize)
 *                                                          clusterDown.cpp:412-417, 505-510
 *   kmerInfo(seq, kLen, revComp) / compareKmers            src/SeekDeepPrograms/SeekDeepProgram/extractor/extractorPairedEnd.cpp:819-824
 *   collapser::runFullClustering(clusters, iters, ...)     clusterDown.cpp:488-518
 *
 * Each entry point below names the reference interface it replaces.  INTEGRATION.md shows the
 * C++ shim (`GpuAligner`) a SeekDeep maintainer would add on the reference side.
 *
 * Rules: plain pointers and sizes only; caller-allocated outputs; the library never frees
 * caller memory; all calls on one ctx are ordered on that ctx's CUDA stream; ctxs are
 * independent (one per host thread, mirroring njhseq's one-aligner-per-thread AlignerPool,
 * src/SeekDeepPrograms/SeekDeepProgram/processClusters/processClusters.cpp:145-155).
 * Errors are int return codes (0 = ok); no exceptions cross the ABI.  There is no CPU
 * fallback: without a CUDA device every compute entry point fails with SDGPU_E_CUDA.
 */
#ifndef SDGPU_H_
#define SDGPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDGPU_ABI_VERSION 3 /* 3: sdq_opts gained the k-mer binning fields, sdq_stats three counters; new entry points only add */
#define SDGPU_MAT_DIM 128 /* substitution matrix is indexed by 7-bit ASCII, like njhseq's mat_[127][127] */

/* return codes */
enum {
  SDGPU_OK = 0,
  SDGPU_E_ARG = 1,         /* bad argument (null pointer, index out of range, ...) */
  SDGPU_E_CUDA = 2,        /* CUDA runtime error or no device */
  SDGPU_E_UNSUPPORTED = 3, /* input outside what the kernels implement (e.g. sequence too long) */
  SDGPU_E_ALPHABET = 4,    /* a base outside the IUPAC nucleotide alphabet */
  SDGPU_E_OVERFLOW = 5,    /* caller buffer too small (e.g. gap list capacity) */
  SDGPU_E_STATE = 6        /* call order violation (e.g. checkKmer without an index) */
};

/* The ten gap penalties of njhseq's gapScoringParameters (positive numbers are penalties).
 * Reference: extractorSetUp.cpp:138-150 names all ten; qluster sets gap=5,1 gapLeft=5,1
 * gapRight=0,0 (clusterDownSetUp.cpp:402-404).  "Query" gaps are gaps placed in the read
 * (second argument, DP columns); "Ref" gaps are gaps placed in the ref (first argument, rows). */
typedef struct sdgpu_gap_pars {
  int32_t gapOpen, gapExtend;
  int32_t gapLeftQueryOpen, gapLeftQueryExtend;
  int32_t gapRightQueryOpen, gapRightQueryExtend;
  int32_t gapLeftRefOpen, gapLeftRefExtend;
  int32_t gapRightRefOpen, gapRightRefExtend;
} sdgpu_gap_pars;

/* Everything the 7-argument aligner constructor takes (clusterDown.cpp:420-426) except
 * maxSize (the library sizes its own scratch) and kMaps (built by sdgpu_build_kmer_index). */
typedef struct sdgpu_params {
  sdgpu_gap_pars gaps;
  int32_t scoreMat[SDGPU_MAT_DIM * SDGPU_MAT_DIM]; /* [refChar * 128 + readChar], parts_.scoring_.mat_ */
  uint32_t primaryQual;      /* qScorePars_.primaryQual_   (25 for --illumina, clusterDownSetUp.cpp:285) */
  uint32_t secondaryQual;    /* qScorePars_.secondaryQual_ (20) */
  uint32_t qualThresWindow;  /* qScorePars_.qualThresWindow_ (2) */
  uint32_t countEndGaps;     /* colOpts_.alignOpts_.countEndGaps_ */
  uint32_t weighHomopolymers;/* colOpts_.iTOpts_.weighHomopolyer_ */
  uint32_t reserved;
} sdgpu_params;

/* One gap run of the global alignment, exactly njhseq's gapInfo(pos_, size_, gapInA_), listed
 * in traceback order (from the alignment's end towards its start). */
typedef struct sdgpu_gap {
  uint32_t pos;    /* position in the gapped sequence's ungapped coordinates where '-' go */
  uint32_t size;
  uint32_t gapInA; /* 1: '-' inserted into the ref (A); 0: into the read (B) */
} sdgpu_gap;

/* njhseq `comparison` + the scalar part of `DistanceComp` (clusterDown.cpp:745-753). */
typedef struct sdgpu_comparison {
  int32_t alnScore;
  uint32_t hqMismatches;
  uint32_t lqMismatches;
  uint32_t lowKmerMismatches;
  uint32_t highQualityMatches;
  uint32_t lowQualityMatches;
  uint32_t numGaps;           /* distances_.alignmentGaps_.size(): counted (non-end unless countEndGaps) gap runs */
  uint32_t basesInAln;        /* distances_.basesInAln_ */
  uint32_t overLappingEvents; /* distances_.overLappingEvents_ */
  uint32_t alnLen;            /* columns of the gapped alignment */
  uint32_t nGapRuns;          /* all gap runs of the traceback, end gaps included */
  uint32_t status;            /* 0 = ok; SDGPU_ST_* bits */
  double oneBaseIndel;
  double twoBaseIndel;
  double largeBaseIndel;
  double eventBasedIdentity;
  double eventBasedIdentityHq;
  double refCoverage;         /* distances_.ref_.coverage_ */
  double queryCoverage;       /* distances_.query_.coverage_ */
} sdgpu_comparison;

#define SDGPU_ST_GAPLIST_TRUNCATED 1u /* more gap runs than the caller's per-pair capacity */

/* flags of sdgpu_align_profile == the three bools of aligner::profileAlignment */
#define SDGPU_CHECK_KMER 1u
#define SDGPU_USING_QUALITY 2u
#define SDGPU_MATCH_QUALITY 4u

/* One par-file line (etc/SeekDeepParameterFiles/illumina_lkmer1:2-23; columns documented in
 * clusterDownSetUp.cpp:63-84) or one OTU line (etc/SeekDeepParameterFiles/otu_99:2-5): njhseq's
 * IterPar, i.e. the thresholds comparison::passErrorProfile tests an errorProfile against. */
typedef struct sdgpu_iter_par {
  uint32_t stopCheck;
  uint32_t smallCutoff;
  double oneBaseIndel, twoBaseIndel, largeBaseIndel;
  double hqMismatches, lqMismatches, lowKmerMismatches;
  uint32_t onPerId;   /* --onPerId / --otu: compare identity against perId instead */
  uint32_t onHqPerId; /* !--regOtu (clusterDownSetUp.cpp:295) */
  double perId;
} sdgpu_iter_par;
#define SDGPU_MAX_PASS_LINES 64

typedef struct sdgpu_ctx sdgpu_ctx; /* opaque; one per host thread; owns one device + one stream */

/* Library / device probing (no reference counterpart). */
int sdgpu_abi_version(void);
int sdgpu_device_count(void);

/* Replaces: aligner::aligner(maxSize, gapInfo_, scoring_, kMaps, qScorePars_, countEndGaps_,
 * weighHomopolyer_), clusterDown.cpp:420-426. */
int sdgpu_create(int device, const sdgpu_params* params, sdgpu_ctx** out);
void sdgpu_destroy(sdgpu_ctx* ctx);
const char* sdgpu_last_error(sdgpu_ctx* ctx); /* ctx may be NULL: last create() error of this thread */
int sdgpu_set_params(sdgpu_ctx* ctx, const sdgpu_params* params); /* e.g. `alignerObj.weighHomopolymers_ = x` */
int sdgpu_get_params(sdgpu_ctx* ctx, sdgpu_params* out);          /* e.g. reading `alignerObj.qScorePars_` (clusterDown.cpp:429) */

/* Sequence store.  Replaces the caller-owned std::vector<cluster>/seqInfo the aligner reads
 * (seqBase_.seq_, seqBase_.qual_, seqBase_.cnt_; clusterDown.cpp:274-285).  `bases` is 7-bit
 * ASCII, compared character by character through scoreMat exactly like mat_[a][b] (so case
 * matters unless the matrix says otherwise); at most 16 distinct characters per ctx.
 * `quals` are phred values (may be NULL: all 0), `offsets` has n+1
 * entries, `cnts` may be NULL (all 1).  upload replaces the store; append adds to it and
 * returns the index of the first appended sequence in *firstIndex. */
int sdgpu_upload_seqs(sdgpu_ctx* ctx, const uint8_t* bases, const uint8_t* quals,
                      const uint64_t* offsets, const double* cnts, uint32_t n);
/* sdgpu_upload_seqs for sequences that are not packed: sequence s is the lens[s] bytes at seqPtrs[s] (qualities at
 * qualPtrs[s]; qualPtrs may be NULL: all 0).  Saves the caller a packing pass -- the collapse loop's unique reads are
 * views into the caller's FASTQ buffers (clusterDown.cpp:272-288 keeps one seqInfo per unique sequence). */
int sdgpu_upload_seqs_gather(sdgpu_ctx* ctx, const uint8_t* const* seqPtrs, const uint8_t* const* qualPtrs,
                             const uint32_t* lens, const double* cnts, uint32_t n);
int sdgpu_append_seqs(sdgpu_ctx* ctx, const uint8_t* bases, const uint8_t* quals,
                      const uint64_t* offsets, const double* cnts, uint32_t n,
                      uint32_t* firstIndex);
int sdgpu_num_seqs(sdgpu_ctx* ctx, uint32_t* n);
/* The distinct base characters the ctx has met so far (at most 16; A, C, G, T, N are always listed first): what a
 * caller sizes per-symbol tables with, e.g. the consensus column counters (njhseq charCounter). */
int sdgpu_get_alphabet(sdgpu_ctx* ctx, uint8_t chars[16], uint32_t* n);
/* Overwrites seqBase_.cnt_ of stored sequences (clusters grow as they absorb reads; the k-mer
 * index weights every k-mer by the current count, clusterDown.cpp:505-510). */
int sdgpu_update_cnts(sdgpu_ctx* ctx, const uint32_t* seqIdx, const double* cnts, uint32_t n);

/* Replaces: indexKmers(clusters, kLength_, runCutOff_, kmersByPosition_, expandKmerPos_,
 * expandKmerSize_) -> KmerMaps, clusterDown.cpp:412-417 and 505-510, followed by
 * `alignerObj.kMaps_ = recalcKmaps` (clusterDown.cpp:511).  Indexes the sequences listed in
 * seqIdx (NULL: the whole store), weighting each k-mer by round(cnt). */
int sdgpu_build_kmer_index(sdgpu_ctx* ctx, const uint32_t* seqIdx, uint32_t nIdx, uint32_t kLen,
                           uint32_t runCutOff, int kmersByPosition, int expandKmerPos,
                           uint32_t expandKmerSize);
/* Look up KmerMaps counts (test/debug hook; ≙ kMaps.kmersByPos_[pos][kmer].count_). */
int sdgpu_kmer_index_lookup(sdgpu_ctx* ctx, const uint8_t* kmers /* n*kLen ASCII */,
                            const uint32_t* positions, uint32_t n, uint32_t* countsOut);

/* Replaces: kmerInfo(seq, kLen, false) for every stored sequence (extractorPairedEnd.cpp:819,
 * PrimersAndMids.cpp:49-54).  Builds dense 4^kLen count profiles on the device (kLen <= 8). */
int sdgpu_build_kmer_profiles(sdgpu_ctx* ctx, uint32_t kLen);
/* Replaces: a.compareKmers(b) -> std::pair<uint32_t,double> and, with revCompB != 0,
 * a.compareKmersRevComp(b) where b = kmerInfo(seq, kLen, true) (extractorPairedEnd.cpp:819-824: the k-mers of a
 * against the k-mers of b's reverse complement) for nPairs (a[i], b[i]) index pairs into the store.  shared is a
 * uint32 count and frac the exact double shared / (min(lenA, lenB) - kLen + 1): both identical to the reference's,
 * not merely within the 1e-6 bar.  The reverse-complement profiles are derived from the forward ones on first
 * use; an alphabet holding a character without an IUPAC complement in the same alphabet is refused
 * (SDGPU_E_ALPHABET). */
int sdgpu_kmer_compare(sdgpu_ctx* ctx, const uint32_t* a, const uint32_t* b, uint32_t nPairs, int revCompB,
                       uint32_t* sharedOut, double* fracOut);
/* All-pairs form used by the qluster prefilter: every read in reads[] against every cluster
 * in clusters[]; outputs are row-major [nReads][nClusters]. */
int sdgpu_kmer_compare_all(sdgpu_ctx* ctx, const uint32_t* reads, uint32_t nReads,
                           const uint32_t* clusters, uint32_t nClusters, uint32_t* sharedOut,
                           double* fracOut);

/* k-mer binning in front of the collapse (qluster --useKmerBinning / --kmerCutOff / --kCompareLen,
 * src/SeekDeepPrograms/SeekDeepProgram/clusterDown/clusterDownSetUp.cpp:352-354; njhseq collapser::runFullClustering
 * bins its clusters by k-mer similarity to a bin's first read before it aligns anything).
 * sdgpu_build_kmer_lists replaces kmerInfo(seq, kLen, false) for every stored sequence at any 1 <= kLen <= 16 (the
 * dense profiles above stop at 16 384 bins): the sequence's own k-mers, sorted, on the device.
 * sdgpu_kmer_bin_masks replaces seed.compareKmers(read) for every (read, seed) of reads[] x seeds[] (nSeeds <= 64):
 * bit c of maskOut[r] is set <=> the shared fraction -- the exact double shared / (min(lenSeed, lenRead) - kLen + 1)
 * -- is > cutOff; only 8 bytes per read return.  sharedOut (may be NULL) receives the shared counts, row-major
 * [nReads][nSeeds].  Sequences appended after the last sdgpu_build_kmer_lists call are not covered (SDGPU_E_ARG). */
int sdgpu_build_kmer_lists(sdgpu_ctx* ctx, uint32_t kLen);
int sdgpu_kmer_bin_masks(sdgpu_ctx* ctx, const uint32_t* seeds, uint32_t nSeeds, const uint32_t* reads, uint32_t nReads,
                         double cutOff, uint64_t* maskOut, uint32_t* sharedOut);

/* Replaces, per pair i: alignerObj.alignCacheGlobal(seq[refIdx[i]], seq[readIdx[i]]) followed
 * by alignerObj.profileAlignment(ref, read, checkKmer, usingQuality, doingMatchQuality)
 * (clusterDown.cpp:739-741), returning aligner::comp_ in out[i].  If gapsOut != NULL the gap
 * runs of pair i (alnInfoGlobal::gapInfos_, traceback order) are written to
 * gapsOut[i*gapCap .. i*gapCap + min(nGapRuns, gapCap)); a longer list sets
 * SDGPU_ST_GAPLIST_TRUNCATED in out[i].status. */
int sdgpu_align_profile(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx,
                        uint32_t nPairs, uint32_t flags, sdgpu_comparison* out,
                        sdgpu_gap* gapsOut, uint32_t gapCap);

/* Gap lists only, packed: what baseCluster::calculateConsensus needs of alignCacheGlobal(consensus, member) -- the
 * alignment, i.e. alnInfoGlobal::gapInfos_ -- for nPairs pairs.  Members of a cluster mostly align to its consensus
 * without a gap, so only the pairs that have gap runs come back, as records of 32-bit words in no particular order:
 * [pair index, nRuns | 0x80000000 if the list was cut at gapCap, then (pos, size, gapInA) for each of the
 * min(nRuns, gapCap) runs, in traceback order].  A pair without a record aligned gaplessly.  *recs stays valid until
 * the next call on ctx. */
int sdgpu_align_gap_lists(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs, uint32_t gapCap,
                          const uint32_t** recs, uint64_t* nWords);

/* Per-event part of aligner::comp_: distances_.mismatches_, distances_.lowKmerMismatches_ and
 * distances_.alignmentGaps_, the maps qluster's --writeOutFinalAllByAllComparison dumps column by column
 * (clusterDown.cpp:759-800: refBase, refBasePos, refQual, seqBase, seqBasePos, seqQual, kMerFreqByPos,
 * highQuality for every mismatch; ref_/size/refPos_ for every gap).  sdgpu_align_events runs alignCacheGlobal +
 * profileAlignment like sdgpu_align_profile and, from the traced alignment still on the device, lists the events
 * of pair i in alignment-column order into eventsOut[i*eventCap ...]; nEventsOut[i] = how many the pair has (more
 * than eventCap: the list is cut).  A gap's gapedSequence_ / qualities_ are the `size` bases and qualities of the
 * OTHER sequence starting at seqPos (gap in the ref) or refPos (gap in the read). */
#define SDGPU_EV_MISMATCH_HQ 1u      /* in mismatches_, highQuality(qScorePars_) true */
#define SDGPU_EV_MISMATCH_LQ 2u      /* in mismatches_, highQuality false */
#define SDGPU_EV_MISMATCH_LOWKMER 3u /* in lowKmerMismatches_ */
#define SDGPU_EV_GAP_IN_REF 4u       /* alignmentGaps_ entry with ref_ true ("gap-insertion"): '-' in the ref */
#define SDGPU_EV_GAP_IN_READ 5u      /* ... ref_ false ("gap-deletion"): '-' in the read */
#define SDGPU_EV_END_GAP 0x10u       /* or-ed onto a gap kind: an end gap that profileAlignment did not count
                                        (countEndGaps_ off), hence not in alignmentGaps_ */
typedef struct sdgpu_event {
  uint32_t kind;        /* SDGPU_EV_* */
  uint32_t alnPos;      /* alignment column = key of the map (gap: its first column) */
  uint32_t refPos;      /* mismatch::refBasePos / gap::refPos_: ref bases in front of the column */
  uint32_t seqPos;      /* mismatch::seqBasePos / gap::seqPos_ */
  uint32_t size;        /* gap::size_; 1 for a mismatch */
  uint32_t kmerFreqRef; /* mismatches under SDGPU_CHECK_KMER: KmerMaps count of the k-mer centred on the ref base ... */
  uint32_t kmerFreqSeq; /* ... and on the read base (the by-position table when the index is by position); else 0 */
  uint8_t refBase, seqBase, refQual, seqQual; /* mismatch: both sides; gap: first gapped base / quality on the side that has bases, 0 on the other */
} sdgpu_event;
int sdgpu_align_events(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs, uint32_t flags,
                       sdgpu_comparison* out, sdgpu_event* eventsOut, uint32_t eventCap, uint32_t* nEventsOut);

/* Replaces: iteration.errors_.passErrorProfile(alignerObj.comp_) / the OTU identity test
 * (ControlBencher.hpp:120; clusterDownSetUp.cpp:289-312) for up to 64 par-file lines at once.
 * sdgpu_set_pass_table uploads the lines; sdgpu_align_pass runs alignCacheGlobal +
 * profileAlignment per pair like sdgpu_align_profile but returns only a verdict mask per pair
 * (bit l set <=> the pair's errorProfile passes line l): 8 bytes per pair cross PCIe instead of
 * a whole comparison struct.  The collapse loop (sdgpu_qluster.h) runs on this. */
int sdgpu_set_pass_table(sdgpu_ctx* ctx, const sdgpu_iter_par* lines, uint32_t nLines);
int sdgpu_align_pass(sdgpu_ctx* ctx, const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs,
                     uint32_t flags, uint64_t* passMaskOut);

/* The collapse loop's own pair shape: every read of readIdx[] against every cluster of
 * clusterIdx[] (collapser::collapseWithParameters compares each live read with the first
 * stopCheck clusters of the sorted list; clusterDown.cpp:488-518).  Per (read r, cluster c) with
 * readIdx[r] != clusterIdx[c] the call runs alignCacheGlobal(cluster, read) + profileAlignment +
 * passErrorProfile against every line of the pass table, exactly like sdgpu_align_pass -- but the
 * pair list is generated on the device and only the pairs whose verdict mask is NOT zero come back,
 * as (read ordinal, cluster ordinal, mask) records in no particular order.  A pair that is absent
 * from the records failed every line.  *hits stays valid until the next call on ctx. */
typedef struct sdgpu_grid_hit {
  uint32_t read;    /* ordinal into readIdx[] */
  uint32_t cluster; /* ordinal into clusterIdx[] */
  uint64_t mask;    /* bit l set <=> the pair's errorProfile passes line l */
} sdgpu_grid_hit;
int sdgpu_align_pass_grid(sdgpu_ctx* ctx, const uint32_t* clusterIdx, uint32_t nClusters, const uint32_t* readIdx,
                          uint32_t nReads, uint32_t flags, const sdgpu_grid_hit** hits, uint64_t* nHits);
/* Streaming form: the grid runs in chunks of whole reads (about 2 M pairs each, double-buffered); onChunk is called on the
 * calling thread, in read order, as soon as a chunk has finished -- with that chunk's records and readsDone, the number of
 * leading entries of readIdx[] whose pairs are now all judged -- while the GPU already works on the next chunk.  The
 * collapse loop files verdicts and runs its sequential pass over the finished reads inside the callback
 * (collapser::findMatch consumes a read's comparisons in list order, clusterDown.cpp:488-518).  The records are only
 * valid during the call; the callback must not call into ctx.  onChunk == NULL: records are kept for
 * sdgpu_align_pass_grid. */
typedef void (*sdgpu_grid_chunk_fn)(void* user, const sdgpu_grid_hit* hits, uint64_t nHits, uint32_t readsDone);
int sdgpu_align_pass_grid_stream(sdgpu_ctx* ctx, const uint32_t* clusterIdx, uint32_t nClusters, const uint32_t* readIdx,
                                 uint32_t nReads, uint32_t flags, sdgpu_grid_chunk_fn onChunk, void* user);

/* Chimera probe: the per-(parent, candidate) part of collapser::markChimeras
 * (clusterDown.cpp:566-577; options clusterDownSetUp.cpp:372-381).  Per pair i the call runs
 * alignCacheGlobal(parent, candidate) + profileAlignment(parent, candidate, checkKmer,
 * usingQuality, matchQuality), takes the mismatches markChimeras keeps (all of them if
 * keepLowQualityMismatches, else every class but low-quality), and profiles the alignment
 * columns before the first kept mismatch ("front") and after the last one ("back") against
 * chiOverlap (chiOpts_.chiOverlap_, passErrorProfile semantics; stopCheck/smallCutoff/onPerId
 * ignored).  Only this 32-byte record per pair returns to the host. */
typedef struct sdgpu_chimera_probe {
  int32_t alnScore;
  uint32_t keptMismatches; /* mismatches markChimeras looks at */
  uint32_t numGaps;        /* counted gap runs of the whole alignment */
  uint32_t firstPos;       /* candidate (read) base position of the first kept mismatch */
  uint32_t lastPos;        /* ... of the last kept mismatch */
  uint32_t firstAlnPos;    /* alignment columns of the two */
  uint32_t lastAlnPos;
  uint32_t flags;          /* SDGPU_CHI_* */
} sdgpu_chimera_probe;
#define SDGPU_CHI_FRONT_PASS 1u /* columns [0, firstAlnPos) pass chiOverlap */
#define SDGPU_CHI_BACK_PASS 2u  /* columns (lastAlnPos, alnLen) pass chiOverlap */
int sdgpu_chimera_probe_pairs(sdgpu_ctx* ctx, const uint32_t* parentIdx, const uint32_t* candIdx,
                              uint32_t nPairs, uint32_t flags, const sdgpu_iter_par* chiOverlap,
                              int keepLowQualityMismatches, sdgpu_chimera_probe* out);

/* All-vs-all alignment of the sequences listed in seqIdx (NULL: the whole store): the pair list
 * is the upper triangle (i < j) in row-major order, ref = seqIdx[i], read = seqIdx[j]; the call
 * computes pairs [firstPair, firstPair + nPairs) of it, so disjoint blocks can run on different
 * GPUs with no exchange (BASELINE config 3).  Replaces the all-by-all loop of
 * clusterDown.cpp:727-805 (alignCacheGlobal + profileAlignment per pair, the same columns) and is
 * the kernel of processClusters' population comparisons (processClusters.cpp:226).  Pair lists are
 * generated on the device; 28 bytes per pair come back.  The three indel tallies and the identity of the summary are
 * the comparison's doubles rounded to float (relative error <= 6e-8, inside the 1e-6 bar of the contract; the
 * integer fields are exact); a caller that needs the doubles asks sdgpu_align_profile for the pairs it cares about. */
typedef struct sdgpu_pair_summary {
  int32_t alnScore;
  uint16_t hqMismatches, lqMismatches, lowKmerMismatches, numGaps;
  float oneBaseIndel, twoBaseIndel, largeBaseIndel, eventBasedIdentity;
} sdgpu_pair_summary;
int sdgpu_all_vs_all(sdgpu_ctx* ctx, const uint32_t* seqIdx, uint32_t n, uint32_t flags, uint64_t firstPair,
                     uint64_t nPairs, sdgpu_pair_summary* out);

/* Asynchronous, double-buffered form of sdgpu_align_pass: two slots (0, 1) per ctx.  submit copies
 * the pair lists into pinned staging and enqueues H2D + kernel + D2H on the ctx stream, then
 * returns; wait blocks until that slot's masks are on the host and hands back a pointer to them
 * (valid until the slot is submitted again).  Lets the caller prepare / consume one batch while
 * the GPU works on the other. */
int sdgpu_align_pass_submit(sdgpu_ctx* ctx, int slot, const uint32_t* refIdx, const uint32_t* readIdx,
                            uint32_t nPairs, uint32_t flags);
int sdgpu_align_pass_wait(sdgpu_ctx* ctx, int slot, const uint64_t** masks, uint32_t* nPairs);

/* Device-resident variant for benchmarking the kernels alone: pair lists and outputs are
 * device pointers (from sdgpu_dev_alloc); nothing crosses PCIe.  Asynchronous on the ctx stream. */
int sdgpu_align_profile_dev(sdgpu_ctx* ctx, const uint32_t* dRefIdx, const uint32_t* dReadIdx,
                            uint32_t nPairs, uint32_t flags, sdgpu_comparison* dOut);
int sdgpu_kmer_compare_all_dev(sdgpu_ctx* ctx, const uint32_t* dReads, uint32_t nReads,
                               const uint32_t* dClusters, uint32_t nClusters, uint32_t* dSharedOut,
                               double* dFracOut); /* device pointers; indices are not range-checked */
int sdgpu_dev_alloc(sdgpu_ctx* ctx, uint64_t bytes, void** dptr);
int sdgpu_dev_free(sdgpu_ctx* ctx, void* dptr);
int sdgpu_memcpy_h2d(sdgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int sdgpu_memcpy_d2h(sdgpu_ctx* ctx, void* dst, const void* src, uint64_t bytes);

/* Stream-ordered timing on the ctx's own stream (torch.cuda.Event only sees torch's stream). */
int sdgpu_sync(sdgpu_ctx* ctx);
int sdgpu_timer_start(sdgpu_ctx* ctx);
int sdgpu_timer_stop(sdgpu_ctx* ctx, float* msOut); /* records, synchronizes, returns elapsed ms */
/* Counters: kernels launched by this ctx and DP cells computed since creation. */
int sdgpu_counters(sdgpu_ctx* ctx, uint64_t* kernelLaunches, uint64_t* cellUpdates);
/* The alignment kernel has two exact forward passes: int32 (one pair per warp) and packed fp16
 * (two pairs per warp; used when the launch's bases are within A,C,G,T,N and every DP value is an
 * integer of magnitude <= 2047, which fp16 represents exactly).  mode 0 = choose automatically,
 * 1 = int32 only, 2 = packed only (fails with SDGPU_E_UNSUPPORTED when not applicable).
 * sdgpu_last_kernel reports which one the last launch used (1 or 2).  For tests / profiling. */
int sdgpu_set_kernel(sdgpu_ctx* ctx, int mode);
int sdgpu_last_kernel(sdgpu_ctx* ctx);
/* Banded first pass.  mode 0 (default): every launch of at least 64 pairs first runs a 64-diagonal
 * banded forward pass; a pair whose banded score exceeds the best score any alignment leaving the
 * band could reach (maxSub x the diagonal moves left to it, gap penalties being >= 0) provably has
 * its full-matrix optimum, traceback and errorProfile inside the band and is finished there; all
 * other pairs are recomputed by the full-matrix kernel in the same call.  Results are identical to
 * mode 1 (full-matrix kernels only) by construction; sdgpu_set_kernel(1|2) also turns the band off. */
int sdgpu_set_band(sdgpu_ctx* ctx, int mode);

/* Screen pass in front of the verdict-mask calls (sdgpu_align_pass*, sdgpu_align_pass_grid).  For a pair
 * of equal-length sequences the gapless alignment is scored and profiled in one linear scan, and a
 * score-only banded DP under a perturbed scoring (every score x K, every gap open cheaper by 1, K larger
 * than any path's move count) decides whether that gapless alignment is the UNIQUE optimum: it is iff the
 * perturbed optimum equals K x its score and the score exceeds what any path leaving the band could reach.
 * Then traceback and errorProfile are the gapless ones whatever the tie rules, and the verdict mask of
 * the scan is exact; every other pair goes on to the bit-recording passes unchanged.  mode 0 (default):
 * on whenever applicable; 1: off.  Results are identical either way. */
int sdgpu_set_screen(sdgpu_ctx* ctx, int mode);

/* Measurement bookkeeping for bench.py: with profiling enabled every alignment-kernel launch is
 * bracketed by CUDA events on the ctx stream; sdgpu_profile_read sums them (optionally resets). */
typedef struct sdgpu_profile {
  double alignKernelMs;        /* sum of per-launch device time of the alignment kernel */
  uint64_t alignKernelLaunches;
  uint64_t kernelLaunches;     /* all kernels this library launched on the ctx */
  uint64_t cellUpdates;        /* sum of lenA x lenB over the pairs those launches finished (the GCUPS numerator) */
  uint64_t h2dBytes, d2hBytes; /* bytes the library copied across PCIe */
  uint64_t evaluatedCells;     /* DP cells the kernels really evaluated (band cells + full matrices of the rest) */
  uint64_t bandCells;          /* ... of which inside banded passes */
  uint64_t bandCertifiedCells; /* sum of lenA x lenB over the pairs finished by the banded pass */
  double bandKernelMs;         /* part of alignKernelMs spent in banded passes */
  double screenKernelMs;       /* part of alignKernelMs spent in the screen pass (scan + score-only DP) */
  uint64_t screenPairs;        /* pairs the screen pass looked at */
  uint64_t screenCertified;    /* ... of which it finished (gapless alignment proved to be the unique optimum) */
  uint64_t screenCells;        /* DP cells the score-only pass evaluated (included in evaluatedCells) */
  uint64_t screenCertCells;    /* sum of lenA x lenB over the pairs the screen finished (included in cellUpdates) */
  double screenScanMs;         /* part of screenKernelMs spent in the scan kernel (the rest is the score-only DP) */
} sdgpu_profile;
int sdgpu_profile_enable(sdgpu_ctx* ctx, int on);
int sdgpu_profile_read(sdgpu_ctx* ctx, sdgpu_profile* out, int reset);

/* Roofline denominator for the alignment kernels: measured thread-level INT32 instructions per second on this
 * GPU, of one kind (0 add, 1 max(+sub), 2 fused add-max VIADDMNMX, 3 max3(+sub)) or of a two-pipe mix with both
 * instructions counted (4: max on the ALU pipe + multiply-add on the FMA pipe, 5: fused add-max + multiply-add --
 * the mix the DP kernels issue). */
int sdgpu_measure_int32_peak(sdgpu_ctx* ctx, int which, double* instrPerSec);

#ifdef __cplusplus
}
#endif
#endif /* SDGPU_H_ */
