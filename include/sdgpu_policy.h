/*
 * sdgpu_policy.h — the rules this library restates from njhseq WITHOUT being able to check them here.
 *
 * SeekDeep's hot path lives in njhseq, which is not in /root/reference (SURVEY.md Appendix A marks every such rule
 * with a warning sign).  Each rule has ONE name here.  The first group are compile-time switches consulted by the
 * kernels (seekdeep_b200/csrc), the host loop and the CPU oracle alike: the day upstream turns out to differ, flip
 * the value, rebuild, re-freeze the fixtures -- no kernel surgery.  The second group is structural: the rule is
 * woven into a kernel's formulation, so the header only names the places that implement it.
 *
 * The screen pass (screen_kernel.cu) is the one component whose exactness does not depend on the alignment tie rules
 * at all: it finishes a pair only when the gapless alignment is the UNIQUE optimum.
 */
#ifndef SDGPU_POLICY_H_
#define SDGPU_POLICY_H_

/* ---- switchable ------------------------------------------------------------------------------------------------ */

/* seqInfo::checkQual: a base is high quality when its quality is strictly above primaryQual and its neighbours'
 * strictly above secondaryQual (1), or at-or-above (0).  Used by: qualFlagKernel (kmer_kernels.cu), qualSignature
 * (qluster_host.cpp), overlap trimming (stitch_host.cpp reads the same flags), oracle checkQual. */
#define SDGPU_POLICY_CHECKQUAL_STRICT 1
#if SDGPU_POLICY_CHECKQUAL_STRICT
#define SDGPU_QUAL_ABOVE(q, thres) ((q) > (thres))
#else
#define SDGPU_QUAL_ABOVE(q, thres) ((q) >= (thres))
#endif

/* seqInfo::checkQual's trailing window stops before the last base of the read when it reaches the read's end (1),
 * or inspects it (0).  Same users as above. */
#define SDGPU_POLICY_CHECKQUAL_SKIPS_LAST_BASE 1
#define SDGPU_QUAL_WINDOW_END(pos, window, len) \
  (((pos) + (window) + 1 < (len)) ? (pos) + (window) + 1 : (SDGPU_POLICY_CHECKQUAL_SKIPS_LAST_BASE ? (len)-1 : (len)))

/* profileAlignment's low-k-mer test: a high-quality mismatch is a lowKmerMismatch when the k-mer centred on the
 * mismatching base is low-frequency in the ref OR in the read (1), or in the read only (0).  Used by: finishPair
 * and the chimera probe (align_kernel.cu), the scan kernel (screen_kernel.cu), overlapKernel (stitch_kernel.cu),
 * alignEventsKernel (events_kernel.cu), oracle profileAlignment / sdo_align_events. */
#define SDGPU_POLICY_LOWKMER_EITHER_SIDE 1
#if SDGPU_POLICY_LOWKMER_EITHER_SIDE
#define SDGPU_LOWKMER_FLAGS(refFlags, readFlags) ((refFlags) | (readFlags))
#else
#define SDGPU_LOWKMER_FLAGS(refFlags, readFlags) (readFlags)
#endif

/* kmerCluster::compareRead in the k-mer binning of --useKmerBinning: a cluster joins a bin when the shared fraction
 * is strictly above kmerCutOff (1) or at-or-above (0).  Used by: kmerBinKernel (kmer_list_kernels.cu), oracle
 * runFullClustering. */
#define SDGPU_POLICY_KMER_BIN_STRICT 1
#if SDGPU_POLICY_KMER_BIN_STRICT
#define SDGPU_KMER_BIN_PASSES(frac, cutOff) ((frac) > (cutOff))
#else
#define SDGPU_KMER_BIN_PASSES(frac, cutOff) ((frac) >= (cutOff))
#endif

/* ---- structural (named here, implemented in the places listed) ------------------------------------------------------
 *
 * SDGPU_POLICY_TIE_ORDER        needleMaximum: up >= left >= diag, 'B' (up == left) treated as 'U'.
 *                               align_kernel.cu (sign-bit decisions of the three forward passes, warp traceback,
 *                               alignBandKernel), oracle needleMaximum / runNeedleSave traceback.
 * SDGPU_POLICY_BORDER_INIT      first row / column: running end-gap penalty in all three states (not -inf).
 *                               align_kernel.cu (strip set-up), screen_kernel.cu (initial state of every diagonal),
 *                               oracle runNeedleSave.
 * SDGPU_POLICY_QUERY_IS_COLUMNS "Query" gap penalties apply to gaps in the read (second argument, DP columns).
 *                               DevScoring field names (sdgpu_internal.cuh), oracle runNeedleSave.
 * SDGPU_POLICY_MISMATCH_DEF     a column is a mismatch when mat_[a][b] < 0 (degenerate-base matches score >= 0).
 *                               every errorProfile stage tests the sign of the score-table entry; oracle idem.
 * SDGPU_POLICY_LOWKMER_CENTRED  the k-mer looked up for a base is the one centred on it, clamped to the sequence.
 *                               sd_kmer_count_at (sdgpu_internal.cuh), oracle kmerStartFor.
 * SDGPU_POLICY_HOMOPOLYMER_W    weight of a homopolymer indel = |gapSide - baseSide| / (gapSide + baseSide).
 *                               finishPair (align_kernel.cu), oracle handleGapCounting.
 * SDGPU_POLICY_KMER_FRAC        compareKmers fraction = shared / (min(lenA, lenB) - k + 1).
 *                               kmerCompareKernel / kmerCompareGridKernel / kmerBinKernel, oracle compareKmers.
 * SDGPU_POLICY_SORT_KEY         readVecSorter "totalCount": count desc, average error asc, sequence, first read.
 *                               posLess / clusLess (qluster_host.cpp), oracle clusLess.
 * SDGPU_POLICY_STOPCHECK        a read counts every live candidate it looks at (itself skipped) and stops after stopCheck;
 *                               candidates are the clusters with count > smallCutoff.  collapseWithParameters in
 *                               qluster_host.cpp and in the oracle.
 * SDGPU_POLICY_CONSENSUS_VOTE   first strictly-greatest symbol in ASCII order wins a column; 40 % coverage rule; an
 *                               insertion is kept when its best symbol outnumbers the non-inserting members.
 *                               calculateConsensusBatch (qluster_host.cpp), oracle calculateConsensus.
 * SDGPU_POLICY_OVERLAP_TIE      alignRegGlobalNoInternalGaps: the smallest shift among equally scoring diagonals.
 *                               overlapKernel (stitch_kernel.cu), oracle alignNoInternalGaps.
 */

#endif /* SDGPU_POLICY_H_ */
