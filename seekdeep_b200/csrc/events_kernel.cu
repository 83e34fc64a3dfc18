// events_kernel.cu — the per-event part of aligner::comp_ (distances_.mismatches_, lowKmerMismatches_, alignmentGaps_;
// dumped by qluster's --writeOutFinalAllByAllComparison, clusterDown.cpp:759-800).
//
// The alignment kernels return scalars and gap lists; the event maps are a walk over the traced alignment, which the
// gap list determines.  This kernel does that walk on the device for a batch whose gap lists are still there: one
// thread per pair (the output is used on the final clusters of a sample: thousands of pairs, not millions), columns in
// order, mismatches classified on the same flag bytes the errorProfile stages use, k-mer counts looked up in the index.
#include "sdgpu_internal.cuh"

namespace {

__global__ void __launch_bounds__(128) alignEventsKernel(DevScoring sc, DevSeqs seqs, DevKmerIndex idx, const uint32_t* refIdx,
                                                          const uint32_t* readIdx, uint32_t nPairs, uint32_t flags,
                                                          const sdgpu_comparison* cmp, const sdgpu_gap* gaps, uint32_t gapCap,
                                                          sdgpu_event* events, uint32_t eventCap, uint32_t* nEvents,
                                                          const uint8_t* charOf) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPairs) return;
  const bool checkKmer = flags & SDGPU_CHECK_KMER, usingQuality = flags & SDGPU_USING_QUALITY;
  const uint32_t ia = refIdx[p], ib = readIdx[p];
  const uint64_t offA = seqs.offsets[ia], offB = seqs.offsets[ib];
  const int la = int(seqs.offsets[ia + 1] - offA), lb = int(seqs.offsets[ib + 1] - offB);
  const uint8_t *A = seqs.codes + offA, *B = seqs.codes + offB;
  const uint8_t *qA = seqs.quals + offA, *qB = seqs.quals + offB;
  const uint8_t *fA = seqs.flags + offA, *fB = seqs.flags + offB;
  const uint32_t nRuns = cmp[p].nGapRuns;
  if (nRuns > gapCap) {  // the host re-runs the batch with room for every run
    nEvents[p] = 0xFFFFFFFFu;
    return;
  }
  const sdgpu_gap* gl = gaps + size_t(p) * gapCap;
  const uint32_t alnLen = cmp[p].alnLen;
  sdgpu_event* out = events + size_t(p) * eventCap;
  uint32_t n = 0;
  int pa = 0, pb = 0;
  uint32_t col = 0;
  auto emit = [&](const sdgpu_event& e) {
    if (n < eventCap) out[n] = e;
    ++n;
  };
  auto matched = [&](int upto, bool inA) {  // diagonal columns until the gapped sequence reaches position `upto`
    while ((inA ? pa : pb) < upto && pa < la && pb < lb) {
      const uint8_t a = A[pa], b = B[pb];
      if (sc.score[a * SDGPU_MAX_SYMBOLS + b] < 0) {
        const unsigned fa = fA[pa], fb = fB[pb];
        sdgpu_event e{};
        const bool good = !usingQuality || (fa & fb & SD_FLAG_QUAL_OK);
        if (!good) {
          e.kind = SDGPU_EV_MISMATCH_LQ;
        } else if (checkKmer && (SDGPU_LOWKMER_FLAGS(fa, fb) & SD_FLAG_LOW_KMER)) {
          e.kind = SDGPU_EV_MISMATCH_LOWKMER;
        } else {
          e.kind = SDGPU_EV_MISMATCH_HQ;
        }
        e.alnPos = col;
        e.refPos = uint32_t(pa);
        e.seqPos = uint32_t(pb);
        e.size = 1;
        if (checkKmer) {
          sd_kmer_count_at(idx, A, la, pa, e.kmerFreqRef);
          sd_kmer_count_at(idx, B, lb, pb, e.kmerFreqSeq);
        }
        e.refBase = charOf[a];
        e.seqBase = charOf[b];
        e.refQual = qA[pa];
        e.seqQual = qB[pb];
        emit(e);
      }
      ++pa;
      ++pb;
      ++col;
    }
  };
  for (uint32_t g = nRuns; g-- > 0;) {  // the list is in traceback order: last entry = first gap of the alignment
    const sdgpu_gap gp = gl[g];
    matched(int(gp.pos), gp.gapInA != 0);
    sdgpu_event e{};
    const bool endGap = col == 0 || col + gp.size >= alnLen;
    e.kind = (gp.gapInA ? SDGPU_EV_GAP_IN_REF : SDGPU_EV_GAP_IN_READ) | ((endGap && !sc.countEndGaps) ? SDGPU_EV_END_GAP : 0u);
    e.alnPos = col;
    e.refPos = uint32_t(pa);
    e.seqPos = uint32_t(pb);
    e.size = gp.size;
    if (gp.gapInA) {  // '-' in the ref: the read's bases sit in the gap
      if (pb < lb) {
        e.seqBase = charOf[B[pb]];
        e.seqQual = qB[pb];
      }
      pb += int(gp.size);
    } else {
      if (pa < la) {
        e.refBase = charOf[A[pa]];
        e.refQual = qA[pa];
      }
      pa += int(gp.size);
    }
    col += gp.size;
    emit(e);
  }
  matched(la, true);  // the diagonal run behind the last gap
  nEvents[p] = n;
}

// Packed gap lists of a batch: only the pairs whose alignment has gap runs leave a record -- [pair, nRuns (bit 31: the
// list was cut at gapCap), then (pos, size, gapInA) per kept run] -- at a slot reserved with one atomic.
__global__ void __launch_bounds__(256) packGapListsKernel(const sdgpu_comparison* cmp, const sdgpu_gap* gaps, uint32_t gapCap, uint32_t nPairs,
                                                           uint32_t* packed, unsigned long long* nWords) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nPairs) return;
  const uint32_t n = cmp[p].nGapRuns;
  if (n == 0) return;
  const uint32_t m = n < gapCap ? n : gapCap;
  const unsigned long long at = atomicAdd(nWords, 2ull + 3ull * m);
  packed[at] = p;
  packed[at + 1] = n | (n > gapCap ? 0x80000000u : 0u);
  const sdgpu_gap* gl = gaps + size_t(p) * gapCap;
  for (uint32_t g = 0; g < m; ++g) {
    packed[at + 2 + 3 * g] = gl[g].pos;
    packed[at + 3 + 3 * g] = gl[g].size;
    packed[at + 4 + 3 * g] = gl[g].gapInA;
  }
}

}  // namespace

int sd_launch_pack_gaps(sdgpu_ctx* ctx, const sdgpu_comparison* dCmp, const sdgpu_gap* dGaps, uint32_t gapCap, uint32_t nPairs,
                        uint32_t* dPacked, unsigned long long* dNWords) {
  if (nPairs == 0) return SDGPU_OK;
  packGapListsKernel<<<(nPairs + 255) / 256, 256, 0, ctx->stream>>>(dCmp, dGaps, gapCap, nPairs, dPacked, dNWords);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}

int sd_launch_events(sdgpu_ctx* ctx, const uint32_t* dRef, const uint32_t* dRead, uint32_t nPairs, uint32_t flags,
                     const sdgpu_comparison* dCmp, const sdgpu_gap* dGaps, uint32_t gapCap, sdgpu_event* dEvents, uint32_t eventCap,
                     uint32_t* dNEvents) {
  if (nPairs == 0) return SDGPU_OK;
  if (!ctx->dCharOf) SD_CUDA(ctx, cudaMalloc(&ctx->dCharOf, SDGPU_MAX_SYMBOLS));
  SD_CUDA(ctx, cudaMemcpyAsync(ctx->dCharOf, ctx->charOf, SDGPU_MAX_SYMBOLS, cudaMemcpyHostToDevice, ctx->stream));
  alignEventsKernel<<<(nPairs + 127) / 128, 128, 0, ctx->stream>>>(ctx->scoring, sd_dev_seqs(ctx), sd_dev_index(ctx), dRef, dRead, nPairs,
                                                                  flags, dCmp, dGaps, gapCap, dEvents, eventCap, dNEvents, ctx->dCharOf);
  SD_CUDA(ctx, cudaGetLastError());
  ++ctx->launches;
  return SDGPU_OK;
}
