/*
 * sd_oracle.h — C API of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This directory is the checker for the CUDA hot path: a plain,
 * single-threaded CPU restatement of the njhseq algorithms SeekDeep's qluster calls
 * (aligner::alignCacheGlobal / alignCalc::runNeedleSave, aligner::profileAlignment,
 * kmerInfo::compareKmers, indexKmers, collapser::runFullClustering, collapser::markChimeras).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Nothing under seekdeep_b200/ links, includes or calls it.
 *
 * PARITY UNPINNED: the reference repo ships no tests, fixtures or golden vectors for this
 * path and njhseq (github.com/nickjhathaway/njhseq, pinned transitively to v3.0.1 through
 * configure.py:10 `seqServer:v3.0.1`) is not vendored and cannot be fetched or built here.
 * The restatement follows njhseq's published algorithm as documented in DESIGN.md
 * ("Oracle: restated rules"), anchored on the reference's own call sites (cited per function).
 * POD types are shared with include/sdgpu.h so results compare field by field.
 */
#ifndef SD_ORACLE_H_
#define SD_ORACLE_H_
#include <stdint.h>

#include "../include/sdgpu_policy.h" /* the unverifiable rules, one name each, shared with the product */
#include "../include/sdgpu.h"
#include "../include/sdgpu_stitch.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdo_kmaps sdo_kmaps; /* KmerMaps restatement */

/* alignCalc::runNeedleSave: score + gap runs (traceback order).  Returns number of gap runs
 * (may exceed gapCap; only gapCap are written). */
int sdo_align_global(const sdgpu_params* p, const uint8_t* a, uint32_t lenA, const uint8_t* b,
                     uint32_t lenB, int32_t* scoreOut, sdgpu_gap* gaps, uint32_t gapCap);

/* alignCacheGlobal + profileAlignment for one pair.  alnA/alnB (may be NULL) receive the
 * gapped strings ('-' for gaps), capacity lenA+lenB+1, NUL-terminated. */
int sdo_align_profile(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* a,
                      const uint8_t* qa, uint32_t lenA, const uint8_t* b, const uint8_t* qb,
                      uint32_t lenB, uint32_t flags, sdgpu_comparison* out, sdgpu_gap* gaps,
                      uint32_t gapCap, char* alnA, char* alnB);

/* Batched form over a packed sequence store (same layout as sdgpu_upload_seqs). */
/* per-event lists of comp_ (mismatches_, lowKmerMismatches_, alignmentGaps_; clusterDown.cpp:759-800) */
int sdo_align_events(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* a, const uint8_t* qa, uint32_t lenA,
                     const uint8_t* b, const uint8_t* qb, uint32_t lenB, uint32_t flags, sdgpu_comparison* cmpOut,
                     sdgpu_event* out, uint32_t cap);
int sdo_align_profile_batch(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* bases,
                            const uint8_t* quals, const uint64_t* offsets,
                            const uint32_t* refIdx, const uint32_t* readIdx, uint32_t nPairs,
                            uint32_t flags, sdgpu_comparison* out, sdgpu_gap* gaps,
                            uint32_t gapCap);

/* indexKmers (clusterDown.cpp:412-417). */
sdo_kmaps* sdo_index_kmers(const uint8_t* bases, const uint64_t* offsets, const double* cnts,
                           const uint32_t* seqIdx, uint32_t nIdx, uint32_t kLen,
                           uint32_t runCutOff, int kmersByPosition, int expandKmerPos,
                           uint32_t expandKmerSize);
void sdo_free_kmaps(sdo_kmaps* k);
uint32_t sdo_kmaps_count(const sdo_kmaps* k, const uint8_t* kmer, uint32_t pos);

/* kmerInfo::compareKmers (extractorPairedEnd.cpp:820-824). */
int sdo_kmer_compare(const uint8_t* a, uint32_t lenA, const uint8_t* b, uint32_t lenB,
                     uint32_t kLen, uint32_t* sharedOut, double* fracOut);

/* kmerInfo(b, kLen, true) + a.compareKmersRevComp(b) (extractorPairedEnd.cpp:819-824) */
int sdo_kmer_compare_revcomp(const uint8_t* a, uint32_t lenA, const uint8_t* b, uint32_t lenB,
                             uint32_t kLen, uint32_t* sharedOut, double* fracOut);

/* ---- paired-end stitching: PairedReadProcessor::processPairedEnd (PairedReadProcessor.cpp:287-664) ---- */
typedef struct sdo_stitch_result sdo_stitch_result;
sdo_stitch_result* sdo_stitch_pairs(const sdgpu_params* p, const sdst_params* s, const uint8_t* r1Bases, const uint8_t* r1Quals,
                                    const uint64_t* r1Offsets, const uint8_t* r2Bases, const uint8_t* r2Quals,
                                    const uint64_t* r2Offsets, uint32_t n);
void sdo_stitch_free(sdo_stitch_result* r);
uint64_t sdo_stitch_retried(const sdo_stitch_result* r);
uint32_t sdo_stitch_status(const sdo_stitch_result* r, uint32_t i);
void sdo_stitch_overlap(const sdo_stitch_result* r, uint32_t i, sdgpu_overlap* out);
uint32_t sdo_stitch_len(const sdo_stitch_result* r, uint32_t i);
void sdo_stitch_seq(const sdo_stitch_result* r, uint32_t i, uint8_t* seq, uint8_t* qual);

/* ---- qluster (clusterDown.cpp:205-563 call sequence) ---- */
typedef struct sdo_iter_par { /* one par-file line, etc/SeekDeepParameterFiles/illumina_lkmer1 */
  uint32_t stopCheck;
  uint32_t smallCutoff;
  double oneBaseIndel, twoBaseIndel, largeBaseIndel;
  double hqMismatches, lqMismatches, lowKmerMismatches;
  uint32_t onPerId; /* 1: OTU mode, `perId` is the identity threshold */
  uint32_t onHqPerId;
  double perId;
} sdo_iter_par;

typedef struct sdo_qluster_opts {
  uint32_t kLen;            /* kmerOpts_.kLength_ */
  uint32_t runCutOff;       /* kmerOpts_.runCutOff_ */
  uint32_t kmersByPosition; /* kmerOpts_.kmersByPosition_ */
  uint32_t expandKmerPos;
  uint32_t expandKmerSize;
  uint32_t checkKmers;      /* kmerOpts_.checkKmers_ */
  uint32_t startWithSingles;
  uint32_t leaveOutSinglets;
  uint32_t dontRecalLowFreq;/* dontRecalLowFreqMismatchAndReRun */
  uint32_t smallReadSize;
  /* chimera marking, chiOpts_ (clusterDownSetUp.cpp:372-381; clusterDown.cpp:566-577) */
  uint32_t markChimeras;         /* checkChimeras_ */
  uint32_t chiKeepLowQual;       /* keepLowQaulityMismatches_ */
  uint32_t chiRunCutOff;         /* candidates with cnt_ <= this are skipped */
  uint32_t chiOverlapSizeCutoff; /* overLapSizeCutoff_ */
  uint32_t chiPosSpacing;        /* posSpacing_ */
  uint32_t converge;             /* colOpts_.clusOpts_.converge_ (--converge, clusterDownSetUp.cpp:394): repeat a par line while it merges */
  double chiParentFreqs;         /* parentFreqs_ */
  sdo_iter_par chiOverlap;       /* chiOverlap_ */
  /* colOpts_.kmerBinOpts_ (clusterDownSetUp.cpp:352-354) + pars.binIteratorMap (clusterDown.cpp:489) */
  double kmerCutOff;
  uint32_t useKmerBinning;
  uint32_t kCompareLen;
  const sdo_iter_par* binPars;   /* NULL: the iterations of the run itself */
  uint64_t nBinPars;
} sdo_qluster_opts;

typedef struct sdo_qluster_result sdo_qluster_result;

/* Runs the collapse on n input reads (packed like sdgpu_upload_seqs; quals required).
 * initPars/pars ≙ pars.intialParameters / pars.iteratorMap. */
sdo_qluster_result* sdo_qluster(const sdgpu_params* p, const sdo_qluster_opts* o,
                                const sdo_iter_par* initPars, uint32_t nInit,
                                const sdo_iter_par* pars, uint32_t nPars, const uint8_t* bases,
                                const uint8_t* quals, const uint64_t* offsets, uint32_t n);
/* doPopulationClustering's collapse (processClusters.cpp:226): input = clusters with counts, no dedupe, no k-mer checks */
sdo_qluster_result* sdo_population_cluster(const sdgpu_params* p, const sdo_qluster_opts* o,
                                           const sdo_iter_par* popPars, uint32_t nPars, const uint8_t* bases,
                                           const uint8_t* quals, const uint64_t* offsets, const double* cnts, uint32_t n);
void sdo_qluster_free(sdo_qluster_result* r);
uint32_t sdo_qluster_num_clusters(const sdo_qluster_result* r);
uint32_t sdo_qluster_num_unique(const sdo_qluster_result* r);
uint64_t sdo_qluster_num_alignments(const sdo_qluster_result* r);
uint64_t sdo_qluster_num_cells(const sdo_qluster_result* r);
/* cluster c: consensus sequence/qualities (length returned), total count */
uint32_t sdo_qluster_cluster_len(const sdo_qluster_result* r, uint32_t c);
double sdo_qluster_cluster_cnt(const sdo_qluster_result* r, uint32_t c);
void sdo_qluster_cluster_seq(const sdo_qluster_result* r, uint32_t c, uint8_t* seq, uint8_t* qual);
/* 1 if cluster c was marked chimeric; parents[2] / positions[2] = (front parent, back parent)
 * cluster indices and the candidate base positions of the two crossover mismatches */
int sdo_qluster_cluster_chimeric(const sdo_qluster_result* r, uint32_t c, uint32_t* parents, uint32_t* positions);
/* collapser::markChimeras' per-pair probe over a packed store (checks sdgpu_chimera_probe_pairs) */
int sdo_chimera_probe_batch(const sdgpu_params* p, const sdo_kmaps* kmaps, const uint8_t* bases, const uint8_t* quals,
                            const uint64_t* offsets, const uint32_t* parentIdx, const uint32_t* candIdx,
                            uint32_t nPairs, uint32_t flags, const sdo_iter_par* chiOverlap,
                            int keepLowQualityMismatches, sdgpu_chimera_probe* out);
/* membership: for each input read, the final cluster index (UINT32_MAX: dropped as small read) */
void sdo_qluster_membership(const sdo_qluster_result* r, uint32_t* clusterOfRead);

#ifdef __cplusplus
}
#endif
#endif
