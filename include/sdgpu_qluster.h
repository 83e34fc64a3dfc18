/*
 * sdgpu_qluster.h — C ABI of the host-side collapse loop that drives the sdgpu kernels.
 *
 * Replaces, for one sample, the part of SeekDeep's qluster body that sits on the hot path:
 * src/SeekDeepPrograms/SeekDeepProgram/clusterDown/clusterDown.cpp:230-563 —
 *   exact-sequence dedupe (:272-288), per-base median quality (:293-359), std::sort (:389),
 *   indexKmers (:412-417), singlet split (:467-484), collapser::runFullClustering x3
 *   (:488-490, :492-499, :502-519) and the final sortReadVector "totalCount" (:560).
 * The loop keeps njhseq's sequential, size-ordered greedy semantics (first passing cluster wins,
 * consensus rebuilt at the end of each par-file iteration); the GPU is fed speculative batches of
 * the candidate (cluster, read) pairs each iteration can ask for, and the results are served to
 * the unchanged sequential logic from a pair cache — the same "lookup-or-compute" contract
 * aligner::alignCacheGlobal already has (clusterDown.cpp:694, 739).
 *
 * With opts.markChimeras the run ends with collapser::markChimeras (:566-577) on the sorted
 * clusters, every (parent, candidate) probe of it computed in one sdgpu_chimera_probe_pairs batch.
 *
 * Out of scope here (left to the caller, as in the reference): FASTQ parsing, Illumina sample
 * filtering / down-sampling (:83-181), cluster renaming, output files.
 */
#ifndef SDGPU_QLUSTER_H_
#define SDGPU_QLUSTER_H_

#include "sdgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One par-file line (etc/SeekDeepParameterFiles/illumina_lkmer1:2-23; column meaning documented in
 * clusterDownSetUp.cpp:63-84) or one OTU line (etc/SeekDeepParameterFiles/otu_99:2-5). */
typedef sdgpu_iter_par sdq_iter_par;

/* The qluster options that shape the hot path (setUpPars.hpp:112-187 + seqSetUp kmerOpts_). */
typedef struct sdq_opts {
  uint32_t kLen;             /* kmerOpts_.kLength_ */
  uint32_t runCutOff;        /* kmerOpts_.runCutOff_ (after processRunCutoff, clusterDown.cpp:397) */
  uint32_t kmersByPosition;  /* kmerOpts_.kmersByPosition_ */
  uint32_t expandKmerPos;    /* pars_.expandKmerPos_ */
  uint32_t expandKmerSize;   /* pars_.expandKmerSize_ */
  uint32_t checkKmers;       /* kmerOpts_.checkKmers_ */
  uint32_t startWithSingles; /* --startWithSingles */
  uint32_t leaveOutSinglets; /* --leaveOutSinglets */
  uint32_t dontRecalLowFreq; /* --dontRecalLowFreqMismatchAndReRun */
  uint32_t smallReadSize;    /* reads with len <= this are dropped (clusterDown.cpp:269) */
  /* chimera marking: pars_.chiOpts_ (clusterDownSetUp.cpp:372-381; clusterDown.cpp:566-577) */
  uint32_t markChimeras;         /* checkChimeras_ (qluster: on unless --noMarkChimeras) */
  uint32_t chiKeepLowQual;       /* keepLowQaulityMismatches_ (--chiKeepLowQaulityMismatches) */
  uint32_t chiRunCutOff;         /* candidates with cnt_ <= this are not examined (njhseq default 1) */
  uint32_t chiOverlapSizeCutoff; /* overLapSizeCutoff_: bases a parent must explain (njhseq default 5) */
  uint32_t chiPosSpacing;        /* posSpacing_ (--chiPosSpacing) */
  uint32_t converge;             /* colOpts_.clusOpts_.converge_ (--converge, clusterDownSetUp.cpp:394): repeat a par line while it merges */
  double chiParentFreqs;         /* parentFreqs_ (--parFreqs, 2) */
  sdq_iter_par chiOverlap;       /* chiOverlap_ (njhseq: oneBaseIndel 2, twoBaseIndel 1, lowKmer 1;
                                    qluster sets largeBaseIndel .99, clusterDown.cpp:572) */
  /* k-mer binning in front of every clustering run: colOpts_.kmerBinOpts_ (clusterDownSetUp.cpp:352-354).  The clusters
   * are binned greedily by compareKmers against each bin's first cluster (a cluster joins the first bin whose seed
   * shares more than kmerCutOff of its k-mers, else it opens a bin), every bin of two or more is collapsed on its own
   * with binPars (pars.binIteratorMap, clusterDown.cpp:489), and the survivors then go through the run's own
   * iterations.  The k-mer comparisons run on the GPU (sdgpu_build_kmer_lists / sdgpu_kmer_bin_masks). */
  double kmerCutOff;             /* kmerCutOff_ (--kmerCutOff) */
  uint32_t useKmerBinning;       /* useKmerBinning_ (--useKmerBinning) */
  uint32_t kCompareLen;          /* kCompareLen_ (--kCompareLen), 1..16 */
  const sdq_iter_par* binPars;   /* pars.binIteratorMap (--binPar); NULL: the iterations of the run itself */
  uint64_t nBinPars;
} sdq_opts;

typedef struct sdq_result sdq_result;

/* Statistics of one run (the reference logs only the alignment count, clusterDown.cpp:1008). */
typedef struct sdq_stats {
  uint64_t inputReads, uniqueSeqs, finalClusters;
  uint64_t pairsComputed;   /* (cluster, read) comparisons computed on the GPU */
  uint64_t pairsConsumed;   /* comparisons the sequential loop actually looked at */
  uint64_t consensusAligns; /* member-vs-consensus alignments */
  uint64_t cellUpdates;     /* DP cells of all of the above */
  uint64_t gpuBatches;      /* alignment kernel launches */
  uint64_t onDemandPairs;   /* pairs computed outside the speculative batches */
  uint64_t chimeraPairs;    /* (parent, candidate) probes of markChimeras */
  uint64_t chimericClusters;
  uint64_t passingPairs;    /* computed comparisons whose verdict mask was not zero (the only ones that cross PCIe) */
  uint64_t kmerPairs;       /* (seed, cluster) k-mer comparisons of --useKmerBinning */
  uint64_t kmerBins;        /* bins those comparisons formed, summed over the clustering runs */
  uint64_t streamedGrids;   /* (read x window) grids whose chunks were consumed by the sequential pass while the GPU worked on the next */
  double secondsGpu;        /* wall time inside sdgpu_align_profile calls */
  double secondsTotal;
  double msResident;        /* CUDA-event time on the ctx stream from "unique reads resident in HBM" to the end */
} sdq_stats;

/* Runs the collapse for one sample on ctx's GPU.  Reads are packed like sdgpu_upload_seqs
 * (quals required).  The ctx's sequence store and k-mer index are replaced. */
int sdq_run(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit,
            const sdq_iter_par* pars, uint32_t nPars, const uint8_t* bases, const uint8_t* quals,
            const uint64_t* offsets, uint32_t n, sdq_result** out);
void sdq_free(sdq_result* r);
const char* sdq_last_error(void);
int sdq_get_stats(const sdq_result* r, sdq_stats* out);
uint32_t sdq_num_clusters(const sdq_result* r);
uint32_t sdq_cluster_len(const sdq_result* r, uint32_t c);
double sdq_cluster_cnt(const sdq_result* r, uint32_t c);
int sdq_cluster_seq(const sdq_result* r, uint32_t c, uint8_t* seq, uint8_t* qual);
/* All final clusters in one call (the result a caller writes out, clusterDown.cpp:886-913): consensus bases
 * and qualities concatenated in cluster order, offsets[nClusters + 1], cnts[nClusters] (seqBase_.cnt_).
 * sdq_clusters_total_len gives the size of the two byte buffers.  Any output pointer may be NULL.
 * (sdq_cluster_len / _cnt return 0 for a NULL result or an index out of range.) */
uint64_t sdq_clusters_total_len(const sdq_result* r);
int sdq_clusters_packed(const sdq_result* r, uint8_t* bases, uint8_t* quals, uint64_t* offsets, double* cnts);
/* 1 if markChimeras flagged cluster c (seqBase_.markAsChimeric()); parents[2] = cluster indices of
 * the parent explaining its front and the one explaining its back, positions[2] = the candidate
 * base positions of the two crossover mismatches (the columns of chiParentsInfo.txt, :578-581) */
int sdq_cluster_chimeric(const sdq_result* r, uint32_t c, uint32_t* parents, uint32_t* positions);
/* final cluster index of every input read (UINT32_MAX: dropped as a small read) */
int sdq_membership(const sdq_result* r, uint32_t* clusterOfRead);

/* FASTQ ingest in front of the collapse (the reference reads through njhseq's SeqInput::readNextRead,
 * clusterDown.cpp:207-232).  sdq_read_fastq parses a plain or gzip FASTQ file (four-line records, Phred+33) into the
 * packed (bases, quals, offsets) form sdq_run takes -- lower-case bases upper-cased, qualities as numbers -- and
 * sdq_run_fastq feeds it to sdq_run.  Malformed input (multi-line records, length mismatch, characters outside
 * Phred+33) is refused with SDGPU_E_UNSUPPORTED and a message naming the record (sdq_fastq_last_error).  Per-read
 * trimming and the Illumina sample-number filter (clusterDown.cpp:83-181, :233-270) stay with the caller. */
typedef struct sdq_reads sdq_reads;
int sdq_read_fastq(const char* path, sdq_reads** out);
void sdq_reads_free(sdq_reads* r);
uint32_t sdq_reads_count(const sdq_reads* r);
const uint8_t* sdq_reads_bases(const sdq_reads* r);
const uint8_t* sdq_reads_quals(const sdq_reads* r);
const uint64_t* sdq_reads_offsets(const sdq_reads* r); /* count + 1 entries */
const char* sdq_reads_name(const sdq_reads* r, uint32_t i); /* without the '@'; NULL when out of range */
const char* sdq_fastq_last_error(void);
int sdq_run_fastq(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit,
                  const sdq_iter_par* pars, uint32_t nPars, const char* fastqPath, sdq_result** out);

/* Population clustering across samples.  Replaces the collapse inside
 * sampColl.doPopulationClustering(sampColl.createPopInput(), alignerObj, collapserObj, pars.popIteratorMap)
 * (src/SeekDeepPrograms/SeekDeepProgram/processClusters/processClusters.cpp:226): the final clusters of every sample
 * -- each keeping its own count, nothing deduplicated -- go through the par-file iterations once, with k-mer checks
 * off (collapserObj.opts_.kmerOpts_.checkKmers_ = false, :127), no singlet round, no k-mer re-count, no chimera
 * marking.  sdq_membership of the result gives, for every input cluster in sample-major order, the population
 * cluster it ended in; consensus sequences, counts and statistics come back like sdq_run's.  The per-sample steps
 * around it (replicate merging, the exclusion filters, renaming; :158-196) are bookkeeping left to the caller. */
int sdq_population_cluster(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* popPars, uint32_t nPars,
                           const sdq_result* const* samples, uint32_t nSamples, sdq_result** out);
/* ... with the input clusters packed by the caller (like sdgpu_upload_seqs; cnts may be NULL: all 1) */
int sdq_population_cluster_packed(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* popPars, uint32_t nPars,
                                  const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, const double* cnts,
                                  uint32_t nClusters, sdq_result** out);

/* Many independent samples (one qluster each, like qlusterCmds.txt + runMultipleCommands,
 * SeekDeepUtilsRunner.cpp:294-356) on ONE GPU: nWorkers host threads, each with its own ctx on
 * `device`, pull samples from a shared queue, so one sample's host-only phases overlap another's
 * alignment batches.  out[s] receives sample s's result (free each with sdq_free). */
int sdq_run_many(int device, const sdgpu_params* params, const sdq_opts* opts, const sdq_iter_par* initPars,
                 uint32_t nInit, const sdq_iter_par* pars, uint32_t nPars, uint32_t nSamples,
                 const uint8_t* const* bases, const uint8_t* const* quals, const uint64_t* const* offsets,
                 const uint32_t* nReads, uint32_t nWorkers, sdq_result** out);

#ifdef __cplusplus
}
#endif
#endif /* SDGPU_QLUSTER_H_ */
