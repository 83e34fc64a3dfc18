// stitch_host.cpp — host side of paired-end stitching (C ABI in include/sdgpu_stitch.h).
//
// Mirrors PairedReadProcessor::processPairedEnd (src/SeekDeep/objects/IlluminaUtils/PairedReadProcessor.cpp:287-664)
// for a batch of pairs.  What the reference does per pair with one stateful aligner --
//   alignRegGlobalNoInternalGaps + profileAlignment (:333-336), the accept rule (:448-452), the low-quality-window
//   trim and second attempt (:358-412), the front / back overhang cases and addToConsensus (:43-78, :453-628)
// -- runs here as: one GPU batch over all pairs, the window search and trimming for the pairs that failed (host,
// threaded), one GPU batch over their trimmed copies, then the overhang cases and the consensus (host, threaded).
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sdgpu_stitch.h"

namespace {

thread_local std::string g_stErr;

unsigned allowedCpus() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) return unsigned(CPU_COUNT(&set));
  const unsigned hw = std::thread::hardware_concurrency();
  return hw ? hw : 1;
}

template <class Fn>
void parallelFor(size_t n, Fn fn, size_t minPerThread = 256) {
  size_t T = std::min<size_t>(allowedCpus(), 8);
  T = std::min(T, std::max<size_t>(1, n / minPerThread));
  if (T <= 1) {
    fn(size_t(0), n);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (n + T - 1) / T;
  for (size_t t = 0; t < T; ++t) {
    const size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back([=] { fn(lo, hi); });
  }
  for (auto& x : th) x.join();
}

struct Read {  // a (possibly trimmed) mate: a view into the caller's buffers
  const uint8_t* seq = nullptr;
  const uint8_t* qual = nullptr;
  uint32_t len = 0;
};

bool accepted(const sdgpu_overlap& o, const sdst_params& P) {  // PairedReadProcessor.cpp:448-452
  const double percentId = 1 - P.errorAllowed;
  return o.eventBasedIdentityHq >= percentId && o.basesInAln >= P.minOverlap &&
         o.hqMismatches + o.lqMismatches <= P.hardMismatchCutOff && o.hqMismatches <= P.hqMismatchCutOff &&
         o.lqMismatches <= P.lqMismatchCutOff;
}

// getFirstFailedWindow / getLastFailedWindow (PairedReadProcessor.cpp:358-389)
uint32_t firstFailedWindow(const Read& r, const sdst_params& P) {
  uint32_t ret = std::numeric_limits<uint32_t>::max();
  if (r.len > P.windowSize) {
    for (uint32_t pos = 0; pos < r.len - P.windowSize; pos += P.windowStep) {
      double sum = 0;
      for (uint32_t q = pos; q < pos + P.windowSize; ++q) sum += r.qual[q];
      if (sum / P.windowSize < P.avgQualCutOff) {
        ret = pos;
        break;
      }
    }
  }
  return ret;
}
uint32_t lastFailedWindow(const Read& r, const sdst_params& P) {
  uint32_t ret = 0;
  const uint32_t end = r.len;
  if (end > P.windowSize + 1) {
    for (uint32_t pos = 0; pos < end - P.windowSize - 1; pos += P.windowStep) {
      const uint32_t truePos = end - P.windowSize - pos - 1;
      double sum = 0;
      for (uint32_t q = truePos; q < truePos + P.windowSize; ++q) sum += r.qual[q];
      if (sum / P.windowSize < P.avgQualCutOff) {
        ret = truePos;
        break;
      }
    }
  }
  return ret;
}

}  // namespace

struct sdst_result {
  std::vector<uint32_t> status;
  std::vector<sdgpu_overlap> overlaps;
  std::vector<uint64_t> offsets;
  std::vector<uint8_t> bases, quals;
  uint64_t counts[7] = {0, 0, 0, 0, 0, 0, 0};
};

extern "C" {

const char* sdst_last_error(void) { return g_stErr.c_str(); }
void sdst_free(sdst_result* r) { delete r; }
uint64_t sdst_total_len(const sdst_result* r) { return r ? r->bases.size() : 0; }

int sdst_fetch(const sdst_result* r, uint32_t* status, sdgpu_overlap* overlaps, uint8_t* bases, uint8_t* quals, uint64_t* offsets) {
  if (!r) return SDGPU_E_ARG;
  if (status) std::memcpy(status, r->status.data(), r->status.size() * sizeof(uint32_t));
  if (overlaps) std::memcpy(overlaps, r->overlaps.data(), r->overlaps.size() * sizeof(sdgpu_overlap));
  if (bases) std::memcpy(bases, r->bases.data(), r->bases.size());
  if (quals) std::memcpy(quals, r->quals.data(), r->quals.size());
  if (offsets) std::memcpy(offsets, r->offsets.data(), r->offsets.size() * sizeof(uint64_t));
  return SDGPU_OK;
}

int sdst_counts(const sdst_result* r, uint64_t counts[7]) {
  if (!r || !counts) return SDGPU_E_ARG;
  std::memcpy(counts, r->counts, sizeof(r->counts));
  return SDGPU_OK;
}

int sdst_process_pairs(sdgpu_ctx* ctx, const sdst_params* pars, const uint8_t* r1Bases, const uint8_t* r1Quals,
                       const uint64_t* r1Offsets, const uint8_t* r2Bases, const uint8_t* r2Quals, const uint64_t* r2Offsets,
                       uint32_t n, sdst_result** out) {
  if (!ctx || !pars || !out || (n && (!r1Bases || !r1Quals || !r1Offsets || !r2Bases || !r2Quals || !r2Offsets))) {
    g_stErr = "sdst_process_pairs: null argument";
    return SDGPU_E_ARG;
  }
  *out = nullptr;
  const sdst_params P = *pars;
  if (P.windowSize == 0 || P.windowStep == 0) {
    g_stErr = "sdst_process_pairs: windowSize and windowStep must be positive";
    return SDGPU_E_ARG;
  }
  auto res = new sdst_result();
  auto fail = [&](int rc, const char* msg) {
    g_stErr = msg ? msg : sdgpu_last_error(ctx);
    delete res;
    return rc;
  };
  res->status.assign(n, SDST_NOOVERLAP);
  res->overlaps.assign(n, sdgpu_overlap{});
  res->offsets.assign(size_t(n) + 1, 0);
  if (n == 0) {
    *out = res;
    return SDGPU_OK;
  }
  sdgpu_params gp;
  if (sdgpu_get_params(ctx, &gp)) return fail(SDGPU_E_CUDA, nullptr);
  // r1Trim / r2Trim (:325-332): r1 loses its last bases; the mate is stored reverse-complemented, so it loses its first ones
  std::vector<Read> R1(n), R2(n);
  for (uint32_t i = 0; i < n; ++i) {
    uint32_t l1 = uint32_t(r1Offsets[i + 1] - r1Offsets[i]), l2 = uint32_t(r2Offsets[i + 1] - r2Offsets[i]);
    uint32_t skip2 = 0;
    if (P.r1Trim) l1 = l1 > P.r1Trim ? l1 - P.r1Trim : 0;
    if (P.r2Trim) {
      skip2 = std::min(l2, P.r2Trim);
      l2 -= skip2;
    }
    if (l1 == 0 || l2 == 0) return fail(SDGPU_E_ARG, "sdst_process_pairs: a mate is empty (after r1Trim / r2Trim)");
    R1[i] = Read{r1Bases + r1Offsets[i], r1Quals + r1Offsets[i], l1};
    R2[i] = Read{r2Bases + r2Offsets[i] + skip2, r2Quals + r2Offsets[i] + skip2, l2};
  }
  // store: r1 of every pair, then r2 of every pair
  {
    std::vector<const uint8_t*> sp(size_t(n) * 2), qp(size_t(n) * 2);
    std::vector<uint32_t> ln(size_t(n) * 2);
    for (uint32_t i = 0; i < n; ++i) {
      sp[i] = R1[i].seq, qp[i] = R1[i].qual, ln[i] = R1[i].len;
      sp[n + i] = R2[i].seq, qp[n + i] = R2[i].qual, ln[n + i] = R2[i].len;
    }
    if (sdgpu_upload_seqs_gather(ctx, sp.data(), qp.data(), ln.data(), nullptr, n * 2)) return fail(SDGPU_E_CUDA, nullptr);
  }
  std::vector<uint32_t> a(n), b(n);
  for (uint32_t i = 0; i < n; ++i) a[i] = i, b[i] = n + i;
  // profileAlignment(r1, r2, false, true, false)
  if (sdgpu_align_no_internal_gaps(ctx, a.data(), b.data(), n, SDGPU_USING_QUALITY, res->overlaps.data())) return fail(SDGPU_E_CUDA, nullptr);
  // second attempt on low-quality-window-trimmed copies for the pairs that failed (:391-412)
  std::vector<uint32_t> retry;
  if (P.trimLowQualWindows) {
    std::vector<uint8_t> want(n, 0);
    std::vector<Read> T1(n), T2(n);
    parallelFor(n, [&](size_t lo, size_t hi) {
      for (size_t i = lo; i < hi; ++i) {
        if (accepted(res->overlaps[i], P)) continue;
        const uint32_t w1 = firstFailedWindow(R1[i], P), w2 = lastFailedWindow(R2[i], P);
        const bool cut1 = w1 != std::numeric_limits<uint32_t>::max() && w1 > 1;
        const bool cut2 = w2 != 0 && w2 + P.windowSize + 2 < R2[i].len;
        if (!cut1 && !cut2) continue;
        Read t1 = R1[i], t2 = R2[i];
        if (cut1) t1.len = w1;  // getSubRead(0, firstMateWindow)
        if (cut2) {             // getSubRead(secondMateWindow + windowSize)
          const uint32_t s = w2 + P.windowSize;
          t2 = Read{R2[i].seq + s, R2[i].qual + s, R2[i].len - s};
        }
        if (double(t1.len) / R1[i].len >= P.percentAfterTrimCutOff && double(t2.len) / R2[i].len >= P.percentAfterTrimCutOff) {
          T1[i] = t1, T2[i] = t2;
          want[i] = 1;
        }
      }
    });
    for (uint32_t i = 0; i < n; ++i)
      if (want[i]) retry.push_back(i);
    if (!retry.empty()) {
      const uint32_t m = uint32_t(retry.size());
      // the trimmed copies join the store behind the 2n originals (appended packed: there are few of them)
      std::vector<uint8_t> tb, tq;
      std::vector<uint64_t> to{0};
      for (int side = 0; side < 2; ++side)
        for (uint32_t x = 0; x < m; ++x) {
          const Read& t = side == 0 ? T1[retry[x]] : T2[retry[x]];
          tb.insert(tb.end(), t.seq, t.seq + t.len);
          tq.insert(tq.end(), t.qual, t.qual + t.len);
          to.push_back(tb.size());
        }
      uint32_t first = 0;
      if (sdgpu_append_seqs(ctx, tb.data(), tq.data(), to.data(), nullptr, 2 * m, &first)) return fail(SDGPU_E_CUDA, nullptr);
      std::vector<uint32_t> ra(m), rb(m);
      for (uint32_t x = 0; x < m; ++x) ra[x] = first + x, rb[x] = first + m + x;
      std::vector<sdgpu_overlap> ro(m);
      if (sdgpu_align_no_internal_gaps(ctx, ra.data(), rb.data(), m, SDGPU_USING_QUALITY, ro.data())) return fail(SDGPU_E_CUDA, nullptr);
      for (uint32_t x = 0; x < m; ++x) {  // alignerObj now holds the trimmed pair's alignment
        res->overlaps[retry[x]] = ro[x];
        R1[retry[x]] = T1[retry[x]];
        R2[retry[x]] = T2[retry[x]];
      }
      res->counts[6] = m;
    }
  }
  // overhang cases and consensus (:448-628).  The reference walks alignment columns and lets addToConsensus decide; in
  // two of its cases the walked range reaches past the overlap into columns where one mate is '-' (R1ALLINR2: the
  // mate's bases behind r1's end; R2ALLINR1: r1's bases in front of the mate's start), and there a base is kept only
  // if its quality reaches primaryQual_ (:52-62).  emit() produces the combined read of pair i, or just its length.
  const int32_t* mat = gp.scoreMat;
  const uint32_t primaryQual = gp.primaryQual;
  auto emit = [&](size_t i, uint8_t* cs, uint8_t* cq) -> uint32_t {
    const sdgpu_overlap& o = res->overlaps[i];
    const int d = o.shift;
    const Read &r1 = R1[i], &r2 = R2[i];
    const int l1 = int(r1.len), l2 = int(r2.len);
    const uint32_t st = res->status[i];
    uint32_t w = 0;
    auto put = [&](uint8_t base, uint8_t q) {
      if (cs) cs[w] = base, cq[w] = q;
      ++w;
    };
    const int i0 = d > 0 ? d : 0, i1 = std::min(l1, l2 + d);  // r1 positions of the overlap
    if (st == SDST_R1ENDSINR2)
      for (int p = 0; p < d; ++p) put(r1.seq[p], r1.qual[p]);  // r1 beginning, as it is
    if (st == SDST_R2ALLINR1)
      for (int p = 0; p < d; ++p)                                // gap in r2: r1's base if it is sure enough
        if (r1.qual[p] >= primaryQual) put(r1.seq[p], r1.qual[p]);
    for (int p = i0; p < i1; ++p) {  // both mates present
      const uint8_t b1 = r1.seq[p], b2 = r2.seq[p - d], q1 = r1.qual[p], q2 = r2.qual[p - d];
      const bool lower = std::islower(b1) || std::islower(b2);
      uint8_t cb;
      if (mat[size_t(b1 & 127) * SDGPU_MAT_DIM + (b2 & 127)] > 0) {  // match: r1's base, the higher quality
        cb = b1;
      } else {  // mismatch: the base of the higher quality, r1 on ties
        cb = q1 >= q2 ? b1 : b2;
      }
      if (lower) cb = uint8_t(std::tolower(cb));
      put(cb, q1 >= q2 ? q1 : q2);
    }
    if (st == SDST_R1ENDSINR2)
      for (int p = l1 - d; p < l2; ++p) put(r2.seq[p], r2.qual[p]);  // r2 ending, as it is
    if (st == SDST_R1ALLINR2)
      for (int p = l1 - d; p < l2; ++p)                              // gap in r1: the mate's base if it is sure enough
        if (r2.qual[p] >= primaryQual) put(r2.seq[p], r2.qual[p]);
    return w;
  };
  std::vector<uint32_t> outLen(n, 0);
  parallelFor(n, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      const sdgpu_overlap& o = res->overlaps[i];
      if (!accepted(o, P)) continue;
      const int d = o.shift, l1 = int(R1[i].len), l2 = int(R2[i].len);
      // alignObjectA_ = '-' x max(0,-d) + r1 + ..., alignObjectB_ = '-' x max(0,d) + r2 + ...: which mate stands alone at each end
      const int front = d > 0 ? 1 : (d < 0 ? 2 : 0);  // 0 none, 1 r1 overhangs, 2 r2 overhangs
      const int tail = l1 - (l2 + d);
      const int back = tail > 0 ? 1 : (tail < 0 ? 2 : 0);
      uint32_t st;
      if (front == 0 && back == 0) {
        st = SDST_PERFECTOVERLAP;
      } else if ((front == 0 || front == 1) && (back == 0 || back == 2)) {
        st = SDST_R1ENDSINR2;
      } else if ((front == 0 || front == 2) && (back == 0 || back == 1)) {
        st = SDST_R1BEGINSINR2;
      } else if (front == 2 && back == 2) {
        st = SDST_R1ALLINR2;
      } else {
        st = SDST_R2ALLINR1;
      }
      res->status[i] = st;
      outLen[i] = emit(i, nullptr, nullptr);
    }
  });
  for (uint32_t i = 0; i < n; ++i) {
    res->offsets[i + 1] = res->offsets[i] + outLen[i];
    ++res->counts[res->status[i]];
  }
  res->bases.resize(res->offsets[n]);
  res->quals.resize(res->offsets[n]);
  parallelFor(n, [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i)
      if (res->status[i] != SDST_NOOVERLAP) emit(i, res->bases.data() + res->offsets[i], res->quals.data() + res->offsets[i]);
  });
  *out = res;
  return SDGPU_OK;
}

}  // extern "C"
