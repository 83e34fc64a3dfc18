// qluster_host.cpp — host side of the collapse loop (C ABI in include/sdgpu_qluster.h).
//
// Mirrors the call sequence of SeekDeep's qluster body (clusterDown.cpp:230-563) and njhseq's
// collapser::runFullClustering / collapseWithParameters / findMatch / cluster::compare /
// baseCluster::calculateConsensus as the reference drives them (clusterDown.cpp:488-518).
// The sequential greedy logic is unchanged; what differs is where comparisons come from:
//   * all par-file lines of the run are uploaded once (sdgpu_set_pass_table); for every pair the
//     GPU returns a 64-bit verdict mask (bit l: errorProfile passes line l), so one alignment
//     answers the same pair for every later iteration;
//   * before an iteration's sequential pass, every (cluster, read) pair the pass can ask for —
//     each live read against the first stopCheck(+1) eligible clusters, the "window" — that the
//     read is not covered for yet goes to the GPU as a (reads x clusters) grid: only the two index
//     lists travel up (sdgpu_align_pass_grid generates the pairs on the device) and only the pairs
//     whose verdict mask is not zero travel back.  A read that was covered for the previous window
//     only needs the window's new entries;
//   * non-zero verdicts live in per-read rows keyed by the cluster's sequence-store index; a pair
//     the read is covered for and whose key is absent failed every line.  Coverage is a pair of
//     serial numbers: the read has been judged against every window since `coverStart`, the cluster
//     version was last in a window at `lastIn`.  A store index names one immutable (sequence,
//     quality) version, so a cached verdict is a pure function of its key (+ the k-mer index, on
//     whose rebuild all rows are dropped) — the same transparency njhseq's previousErrorChecks_ /
//     alnHolder_ caches have.  A cluster whose consensus moved
//     only in quality values keeps its device version as long as seqInfo::checkQual gives the
//     same answer at every base: qualities enter an errorProfile only through checkQual;
//   * consensus rebuilding keeps the per-column counters of a cluster while its base sequence
//     stays the same, aligns only the newly absorbed members (one GPU batch per iteration, gap
//     lists returned) and re-reads the consensus off the counters (integer sums: exact).
// Host threads: the main thread runs the sequential logic; short-lived worker threads share the
// per-item independent phases (read hashing, per-unique median qualities, packing, cluster pool
// set-up, consensus counter updates).  None of them changes an order the results depend on.
#include <sched.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/sdgpu_qluster.h"
#include "../../include/sdgpu_policy.h"

namespace {

// NVTX range over a host phase of the collapse loop (prep / listing + grids / sequential pass / consensus / sort / index),
// so that a timeline shows which phases leave the GPU idle
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

thread_local std::string g_err;

double nowSec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Per-read verdict row: cluster store index -> pass mask.  Only non-zero masks are kept, so a row holds a
// handful of entries: two live in the object itself, the rest in a small vector; look-ups are linear scans.
class Row {
 public:
  const uint64_t* find(uint32_t k) const {
    for (uint32_t i = 0; i < n_ && i < INLINE; ++i)
      if (keys_[i] == k) return &vals_[i];
    for (const auto& e : more_)
      if (e.first == k) return &e.second;
    return nullptr;
  }
  void insert(uint32_t k, uint64_t v) {
    or_ |= v;
    for (uint32_t i = 0; i < n_ && i < INLINE; ++i)
      if (keys_[i] == k) {
        vals_[i] = v;
        return;
      }
    for (auto& e : more_)
      if (e.first == k) {
        e.second = v;
        return;
      }
    if (n_ < INLINE) {
      keys_[n_] = k;
      vals_[n_] = v;
    } else {
      more_.emplace_back(k, v);
    }
    ++n_;
  }
  void clear() {
    std::vector<std::pair<uint32_t, uint64_t>>().swap(more_);
    n_ = 0;
    or_ = 0;
  }
  uint32_t size() const { return n_; }
  uint64_t orMask() const { return or_; }  // superset of every verdict bit held in the row
  template <class F>
  void forEach(F f) const {
    for (uint32_t i = 0; i < n_ && i < INLINE; ++i) f(keys_[i], vals_[i]);
    for (const auto& e : more_) f(e.first, e.second);
  }

 private:
  static constexpr uint32_t INLINE = 2;
  uint64_t or_ = 0;
  uint32_t n_ = 0;
  uint32_t keys_[INLINE] = {0, 0};
  uint64_t vals_[INLINE] = {0, 0};
  std::vector<std::pair<uint32_t, uint64_t>> more_;
};

// Splits [0, n) over a few short-lived host threads (the prep phases before the GPU has work:
// hashing, per-unique median qualities, packing, cluster pool set-up are all per-item independent).
// host threads this process may use: the CPUs it is allowed on (ranks sharing a node are pinned to their
// own core slices), not the machine's core count
static unsigned allowedCpus() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof(set), &set) == 0) {
    const int n = CPU_COUNT(&set);
    if (n > 0) return unsigned(n);
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return hw ? hw : 1;
}

template <class Fn>
static void parallelFor(size_t n, Fn fn, size_t minItemsPerThread = 512) {
  const unsigned hw = allowedCpus();
  size_t T = std::min<size_t>(hw ? hw : 1, 8);
  T = std::min(T, std::max<size_t>(1, n / minItemsPerThread));
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

static const bool g_verboseTiming = getenv("SDQ_TIMING") && atoi(getenv("SDQ_TIMING")) > 1;

// Items of very different cost (clusters absorbing 1 .. thousands of members): threads pull indices
// from a shared counter.
template <class Fn>
static void parallelForDynamic(size_t n, Fn fn, size_t maxThreads = 4, size_t minItemsPerThread = 1) {
  const unsigned hw = allowedCpus();
  size_t T = std::min<size_t>(hw ? hw : 1, maxThreads);
  T = std::min(T, std::max<size_t>(1, n / minItemsPerThread));  // (starting a thread costs as much as tens of small items)
  if (n < 8) T = 1;
  if (T <= 1) {
    for (size_t i = 0; i < n; ++i) fn(i);
    return;
  }
  std::atomic<size_t> nextIdx{0};
  std::vector<std::thread> th;
  for (size_t t = 0; t < T; ++t)
    th.emplace_back([&] {
      for (size_t i = nextIdx.fetch_add(1); i < n; i = nextIdx.fetch_add(1)) fn(i);
    });
  for (auto& x : th) x.join();
}

// A sequence or quality string that lives somewhere else: in the caller's read buffers (unique reads are views
// of their first read), in the run's quality arena (per-base medians) or in a cluster's own consensus storage.
struct View {
  const uint8_t* p = nullptr;
  uint32_t n = 0;
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  const uint8_t* data() const { return p; }
  const uint8_t* begin() const { return p; }
  const uint8_t* end() const { return p + n; }
  uint8_t operator[](size_t i) const { return p[i]; }
  bool equals(const void* q, size_t m) const { return m == n && (n == 0 || std::memcmp(p, q, n) == 0); }
  int compare(const View& o) const {  // std::string::compare on the bytes
    const int c = std::memcmp(p, o.p, std::min(n, o.n));
    return c != 0 ? c : (n < o.n ? -1 : (n > o.n ? 1 : 0));
  }
};

struct Uniq {  // one exact-unique input sequence; its store index is its position in `uniq`
  View seq, qual;
  double cnt = 0;
  uint32_t inputBegin = 0, inputEnd = 0;  // its input reads: Runner::inputList[inputBegin, inputEnd), ascending
};

// The unique sequences a cluster holds.  Almost every cluster of a run is a singleton that is absorbed or stays alone,
// so the first member lives in the object and only clusters that absorb others allocate.
class Members {
 public:
  struct Iter {
    const Members* m;
    size_t i;
    uint32_t operator*() const { return (*m)[i]; }
    Iter& operator++() {
      ++i;
      return *this;
    }
    bool operator!=(const Iter& o) const { return i != o.i; }
  };
  void init(uint32_t u) {
    first_ = u;
    has_ = true;
    more_.clear();
  }
  size_t size() const { return has_ ? 1 + more_.size() : 0; }
  uint32_t operator[](size_t i) const { return i == 0 ? first_ : more_[i - 1]; }
  void append(const Members& o) {
    for (size_t i = 0; i < o.size(); ++i) more_.push_back(o[i]);  // (push_back grows geometrically; an exact reserve per call would not)
  }
  Iter begin() const { return Iter{this, 0}; }
  Iter end() const { return Iter{this, size()}; }

 private:
  uint32_t first_ = 0;
  bool has_ = false;
  std::vector<uint32_t> more_;
};

// njhseq charCounter restricted to what consensus building uses, over the few symbols that occur
constexpr int NSYM = 18;
struct Col {
  uint32_t cnt[NSYM];
  uint32_t qual[NSYM];
  Col() {
    std::memset(cnt, 0, sizeof(cnt));
    std::memset(qual, 0, sizeof(qual));
  }
  uint32_t total() const {
    uint32_t t = 0;
    for (uint32_t v : cnt) t += v;
    return t;
  }
};

struct ConsState {  // column counters of one cluster against one base sequence
  std::string baseSeq;
  uint32_t baseStore = 0;
  size_t nDone = 0;  // members already counted
  std::vector<Col> counters;
  std::map<uint32_t, std::map<uint32_t, Col>> insertions;
  std::map<uint32_t, Col> beginningGap;
};

// The per-iteration loops walk every live cluster but mostly look at a handful of scalars: those come
// first, together with the row's summary words, so that one 64-byte line per cluster serves them.
struct alignas(64) Clus {
  double cnt = 0;
  double avgErr = 0;
  double sumLenCnt = 0;    // sum over the members of length x count (integers: exact in any order), for the average member length
  uint32_t storeIdx = 0;   // device sequence-store entry (verdict-equivalent to (seq, qual))
  uint32_t firstId = 0;    // cluster::firstReadName_ stand-in (tie-break of the sort)
  bool remove = false, needCons = false;
  bool chimeric = false;   // seqBase_.markAsChimeric()
  Row row;                 // verdicts of this cluster *as the read* against larger clusters
  View seq, qual;          // current consensus: its seed read until a consensus rebuild changes it, then own storage
  std::string ownSeq;
  std::vector<uint8_t> ownQual;
  Members members;
  uint32_t chiParent[2] = {0, 0}, chiPos[2] = {0, 0};
  std::vector<uint8_t> devSig;  // checkQual outcome per base of the device version (empty: exact copy)
  std::unique_ptr<ConsState> cons;
};

// seqInfo::checkQual for every base (strictly-greater thresholds, trailing window stops before
// the last base) — the only way qualities enter profileAlignment.
template <class Q>
void qualSignature(const Q& q, const sdgpu_params& P, std::vector<uint8_t>& sig) {
  const int len = int(q.size()), out = int(P.qualThresWindow);
  sig.assign(q.size(), 0);
  for (int pos = 0; pos < len; ++pos) {
    bool ok = SDGPU_QUAL_ABOVE(q[pos], P.primaryQual);
    const int lower = pos > out ? pos - out : 0;
    for (int i = lower; ok && i < pos; ++i) ok = SDGPU_QUAL_ABOVE(q[i], P.secondaryQual);
    const int higher = SDGPU_QUAL_WINDOW_END(pos, out, len);
    for (int i = pos + 1; ok && i < higher; ++i) ok = SDGPU_QUAL_ABOVE(q[i], P.secondaryQual);
    sig[pos] = ok ? 1 : 0;
  }
}

struct ErrLut {
  double v[256];
  ErrLut() {
    for (int q = 0; q < 256; ++q) v[q] = std::pow(10.0, -(double(q) / 10.0));
  }
};

template <class Q>
double averageErrorRate(const Q& q) {
  static const ErrLut lut;  // same pow() values, evaluated once
  if (q.size() == 0) return 0;
  double s = 0;
  for (size_t i = 0; i < q.size(); ++i) s += lut.v[q[i]];
  return s / double(q.size());
}

bool clusLess(const Clus& a, const Clus& b) {  // readVecSorter "totalCount"
  if (a.cnt != b.cnt) return a.cnt > b.cnt;
  if (a.avgErr != b.avgErr) return a.avgErr < b.avgErr;
  if (const int c = a.seq.compare(b.seq)) return c < 0;
  return a.firstId < b.firstId;
}

// The working set of clusters in sorted order.  A Clus is heavy (strings, vectors, a verdict row),
// so the storage never moves (one pool per run) and sorting / dropping only shuffles pointers.
class ClusVec {
 public:
  struct Iter {
    Clus* const* p;
    Clus& operator*() const { return **p; }
    Iter& operator++() {
      ++p;
      return *this;
    }
    bool operator!=(const Iter& o) const { return p != o.p; }
  };
  Clus& operator[](size_t i) { return *v_[i]; }
  const Clus& operator[](size_t i) const { return *v_[i]; }
  size_t size() const { return v_.size(); }
  void push_back(Clus* c) { v_.push_back(c); }
  Iter begin() const { return Iter{v_.data()}; }
  Iter end() const { return Iter{v_.data() + v_.size()}; }
  std::vector<Clus*>& ptrs() { return v_; }

 private:
  std::vector<Clus*> v_;
};

bool sameLine(const sdq_iter_par& a, const sdq_iter_par& b) {  // thresholds only: stopCheck/smallCutoff do not affect verdicts
  return a.oneBaseIndel == b.oneBaseIndel && a.twoBaseIndel == b.twoBaseIndel && a.largeBaseIndel == b.largeBaseIndel &&
         a.hqMismatches == b.hqMismatches && a.lqMismatches == b.lqMismatches && a.lowKmerMismatches == b.lowKmerMismatches &&
         a.onPerId == b.onPerId && a.onHqPerId == b.onHqPerId && a.perId == b.perId;
}

class Runner {
 public:
  Runner(sdgpu_ctx* ctx, const sdq_opts& o) : ctx_(ctx), O(o) {
    std::memset(slotOf, -1, sizeof(slotOf));
  }
  sdgpu_ctx* ctx_;
  const sdq_opts& O;
  sdgpu_params P;  // the ctx's scoring / quality parameters (for qualSignature)
  std::vector<Uniq> uniq;
  std::vector<uint32_t> inputList;   // input read indices grouped by unique sequence
  std::vector<uint8_t> qualArena;    // per-base median qualities of the unique sequences with several reads
  std::vector<sdq_iter_par> lines;  // distinct par-file lines == bits of a verdict mask
  sdq_stats st{};
  uint32_t storeSize = 0;
  uint64_t serial = 0;                // iteration counter across the three clustering runs
  std::vector<uint32_t> prevWindow;   // store indices of the previous iteration's window
  std::vector<uint8_t> inPrevWindow;  // by store index
  std::vector<uint32_t> winPos;       // by store index: position + 1 in the current iteration's window, 0 = not in it
  uint64_t lastListSerial = 0;        // serial of the latest iteration that listed (judged) every live read
  // Compact copies, by position in the sorted cluster list, of what the per-iteration loops look at for EVERY
  // cluster: rebuilt whenever the list is re-sorted, updated in place at the few points an iteration changes
  // them.  The loops then touch a heavy Clus object only for the few reads that have a passing verdict.
  struct Pos {
    double cnt, avgErr;  // sort keys as of the last sort (readVecSorter "totalCount": count desc, average error asc)
    Clus* c;
    uint64_t cover;      // first of the unbroken run of iterations in which this read was judged against the whole window (0: none)
    uint64_t orMask;     // row.orMask()
    uint32_t store;      // storeIdx
    uint8_t remove;
    uint8_t changed;     // absorbed reads since the last sort: its keys / device version may have moved
  };
  std::vector<Pos> PR;                // parallel to the cluster list
  bool posValid = false;              // P matches the list and only `remove` / `changed` entries moved since the last sort
  std::vector<uint64_t> lastIn;       // by store index: latest iteration serial whose window held this version
  std::vector<uint64_t> onDemandAt;   // by store index: iteration serial in which it was judged on demand
  static constexpr uint32_t BATCH = 1u << 21;
  double tSpec = 0, tPass = 0, tCons = 0, tInsert = 0, tSort = 0, tIndex = 0, tSubmit = 0;
  double tConsAlign = 0, tConsCount = 0, tConsBuild = 0, tConsUpload = 0;  // parts of tCons
  double tBin = 0;  // k-mer binning (lists + masks + settling)
  uint64_t streamMinPairs = getenv("SDQ_STREAM_MIN_PAIRS") ? uint64_t(atoll(getenv("SDQ_STREAM_MIN_PAIRS"))) : (3ull << 21);  // (test hook)
  uint64_t streamedGrids = 0, nStreamCb = 0, streamedReads = 0;
  double tStreamCb = 0;  // time inside the streaming callbacks (filing + the streamed part of the pass)
  uint64_t keptVersions = 0, skippedScans = 0;
  bool orderDirty = true;
  // symbols of the consensus counters (bases + '-'), ASCII order kept for njhseq's tie rule
  int8_t slotOf[128];
  int nSlots = 0;
  std::vector<std::pair<char, int>> slotOrder;

  int fail(int rc) {
    g_err = sdgpu_last_error(ctx_);
    return rc;
  }

  uint32_t flags() const { return (O.checkKmers ? SDGPU_CHECK_KMER : 0u) | SDGPU_USING_QUALITY; }

  int lineOf(const sdq_iter_par& it) {
    for (size_t i = 0; i < lines.size(); ++i)
      if (sameLine(lines[i], it)) return int(i);
    lines.push_back(it);
    return int(lines.size()) - 1;
  }

  int slot(char ch) {
    const int c = uint8_t(ch) & 127;
    if (slotOf[c] < 0) {
      if (nSlots >= NSYM) return -1;
      slotOf[c] = int8_t(nSlots++);
      slotOrder.emplace_back(char(c), int(slotOf[c]));
      std::sort(slotOrder.begin(), slotOrder.end());
    }
    return slotOf[c];
  }
  bool add(Col& col, char b, uint32_t q, uint32_t w) {
    const int s = slot(b);
    if (s < 0) return false;
    col.cnt[s] += w;
    col.qual[s] += q * w;
    return true;
  }
  void best(const Col& col, char& b, uint32_t& q, uint32_t& bc) const {  // first strictly-greatest in ASCII order
    bc = 0;
    b = ' ';
    q = 0;
    for (const auto& so : slotOrder)
      if (col.cnt[so.second] > bc) {
        bc = col.cnt[so.second];
        b = so.first;
        q = col.qual[so.second];
      }
  }

  // One (reads x clusters) grid on the GPU; the non-zero verdict masks are filed into the reads' rows.
  // readPos[r] is the position in `cs` of the read behind readStore[r].
  int gridPass(ClusVec& cs, const std::vector<uint32_t>& clusStore, const std::vector<uint32_t>& readStore,
               const std::vector<uint32_t>& readPos) {
    if (clusStore.empty() || readStore.empty()) return 0;
    const sdgpu_grid_hit* hits = nullptr;
    uint64_t nHits = 0;
    const double t0 = nowSec();
    int rc = sdgpu_align_pass_grid(ctx_, clusStore.data(), uint32_t(clusStore.size()), readStore.data(), uint32_t(readStore.size()),
                                   flags(), &hits, &nHits);
    st.secondsGpu += nowSec() - t0;
    if (rc) return fail(rc);
    ++st.gpuBatches;
    st.pairsComputed += uint64_t(clusStore.size()) * readStore.size();
    const double t1 = nowSec();
    for (uint64_t i = 0; i < nHits; ++i) {
      const uint32_t rp = readPos[hits[i].read];
      cs[rp].row.insert(clusStore[hits[i].cluster], hits[i].mask);
      PR[rp].orMask |= hits[i].mask;
    }
    tInsert += nowSec() - t1;
    st.passingPairs += nHits;
    return 0;
  }
  // has `read` been judged against cluster version `s` (so that an absent row entry means "fails every line")?
  bool covered(uint64_t coverStart, uint32_t s) const {
    return (coverStart != 0 && lastIn[s] >= coverStart) || onDemandAt[s] == serial;
  }

  // what the streaming grid call hands back to the collapse loop, chunk by chunk (sdgpu_grid_chunk_fn)
  struct StreamCtx {
    Runner* R;
    ClusVec* cs;
    const std::vector<uint32_t>* window;   // store indices of the grid's clusters
    const std::vector<uint32_t>* readPos;  // list position of every read of the grid, in pass order
    std::function<void(size_t)>* fastStep;
    bool* windowIntact;
    size_t* streamed;
    uint64_t nHits;
    static void onChunk(void* user, const sdgpu_grid_hit* hits, uint64_t n, uint32_t readsDone) {
      StreamCtx& S = *static_cast<StreamCtx*>(user);
      Runner& R = *S.R;
      double t0 = nowSec();
      for (uint64_t i = 0; i < n; ++i) {
        const uint32_t rp = (*S.readPos)[hits[i].read];
        (*S.cs)[rp].row.insert((*S.window)[hits[i].cluster], hits[i].mask);
        R.PR[rp].orMask |= hits[i].mask;
      }
      S.nHits += n;
      double t1 = nowSec();
      R.tInsert += t1 - t0;
      // the reads of the finished chunks take their turn, as long as the window is intact (once a window cluster has
      // been absorbed a read may need a candidate judged on demand: that is the pass proper's business)
      while (*S.streamed < readsDone && *S.windowIntact) {
        const uint32_t rp = (*S.readPos)[*S.streamed];
        if (!R.PR[rp].remove) (*S.fastStep)(rp);
        ++*S.streamed;
      }
      R.tPass += nowSec() - t1;
      R.tStreamCb += nowSec() - t0;
      ++R.nStreamCb;
    }
  };

  // njhseq collapser::collapseWithParameters + findMatch for one par-file line
  int collapseWithParameters(ClusVec& cs, const sdq_iter_par& it, int line) {
    ++serial;
    // (the list was compacted and re-sorted after the last iteration that merged anything: nothing in it is removed,
    //  and counts descend, so the candidate clusters of the smallCutoff rule are a prefix of it)
    const size_t alive = PR.size();
    if (alive < 2) return 0;
    const uint64_t bit = 1ull << line;
    const uint32_t eligEnd = uint32_t(std::partition_point(PR.begin(), PR.end(), [&](const Pos& x) { return x.cnt > double(it.smallCutoff); }) -
                                      PR.begin());
    // ---- speculative batch: the window = first stopCheck(+1: a read skips itself) candidates
    double tt0 = nowSec();
    std::unique_ptr<NvtxRange> gridRange(new NvtxRange("sdq:list+grids"));
    const uint32_t wn = uint32_t(std::min<uint64_t>(eligEnd, uint64_t(it.stopCheck) + 1));  // the window: positions [0, wn)
    size_t amountAdded = 0;
    bool windowIntact = true;   // no window cluster absorbed yet: every read sees the static window
    size_t streamed = 0;        // leading entries of the whole-window read list the streaming callback has dealt with
    size_t resumeAt = cs.size();  // the pass proper starts below this position
    // One read's turn while the window is intact: the candidates it looks at are window positions -- all alive, all
    // judged -- so its first passing candidate is the smallest window position among its own (few) non-zero verdicts,
    // provided it lies within the first stopCheck candidates the read counts (itself skipped).
    std::function<void(size_t)> fastStep = [&](size_t rp) {
      if (!(PR[rp].orMask & bit)) {  // judged against every window cluster, no verdict has this line's bit
        ++skippedScans;
        return;
      }
      Clus& read = cs[rp];
      uint32_t best = UINT32_MAX;
      read.row.forEach([&](uint32_t k, uint64_t v) {
        if (!(v & bit) || k >= winPos.size()) return;
        const uint32_t wp = winPos[k];
        if (wp != 0 && wp - 1 != rp && wp - 1 < best) best = wp - 1;
      });
      const uint32_t looked = best == UINT32_MAX ? std::min<uint32_t>(wn - (rp < wn ? 1u : 0u), it.stopCheck) : best + 1 - (rp < best ? 1u : 0u);
      if (best != UINT32_MAX && looked > it.stopCheck) best = UINT32_MAX;  // beyond the read's stopCheck-th candidate
      st.pairsConsumed += std::min<uint32_t>(looked, it.stopCheck);
      if (best == UINT32_MAX) return;
      Clus& clus = cs[best];
      clus.cnt += read.cnt;
      clus.sumLenCnt += read.sumLenCnt;
      clus.members.append(read.members);
      clus.needCons = true;
      read.remove = true;
      PR[rp].remove = 1;
      PR[best].changed = 1;
      read.row.clear();
      read.cons.reset();
      if (rp < wn) windowIntact = false;
      ++amountAdded;
    };
    {
      if (inPrevWindow.size() < storeSize) inPrevWindow.resize(storeSize, 0);
      if (lastIn.size() < storeSize) lastIn.resize(storeSize, 0);
      if (onDemandAt.size() < storeSize) onDemandAt.resize(storeSize, 0);
      std::vector<uint32_t> window, fresh;  // store indices: the whole window / its entries the previous one lacked
      for (uint32_t w = 0; w < wn; ++w) {
        const uint32_t sIdx = PR[w].store;
        window.push_back(sIdx);
        if (!inPrevWindow[sIdx]) fresh.push_back(sIdx);
      }
      // reads covered for the previous window only need the fresh entries; the others the whole window
      std::vector<uint32_t> fullStore, fullPos, incStore, incPos;
      const bool unbroken = lastListSerial + 1 == serial;  // the previous iteration judged every live read
      for (uint32_t rp = uint32_t(cs.size()); rp-- > 0;) {   // same order as the pass below
        if (PR[rp].remove) continue;
        if (unbroken && PR[rp].cover != 0) {
          incStore.push_back(PR[rp].store);
          incPos.push_back(rp);
        } else {
          fullStore.push_back(PR[rp].store);
          fullPos.push_back(rp);
          PR[rp].cover = serial;
        }
      }
      lastListSerial = serial;
      for (uint32_t s : prevWindow) inPrevWindow[s] = 0;
      prevWindow = window;
      if (winPos.size() < storeSize) winPos.resize(storeSize, 0);
      for (uint32_t w = 0; w < wn; ++w) {
        const uint32_t s = window[w];
        inPrevWindow[s] = 1;
        lastIn[s] = serial;
        winPos[s] = w + 1;
      }
      const double tList = nowSec() - tt0;
      // A large grid of whole-window reads (the iteration in which the singletons meet the clusters) is consumed chunk by
      // chunk: verdicts are filed and the pass below runs over the finished reads -- in list order, which is the pass's
      // order -- while the GPU works on the next chunk.
      int rc = 0;
      const bool stream = incStore.empty() && uint64_t(fullStore.size()) * window.size() >= streamMinPairs;
      if (stream) {
        StreamCtx sc{this, &cs, &window, &fullPos, &fastStep, &windowIntact, &streamed, 0};
        const double t0 = nowSec();
        rc = sdgpu_align_pass_grid_stream(ctx_, window.data(), uint32_t(window.size()), fullStore.data(), uint32_t(fullStore.size()), flags(),
                                          &StreamCtx::onChunk, &sc);
        st.secondsGpu += nowSec() - t0;
        if (rc) return fail(rc);
        ++st.gpuBatches;
        st.pairsComputed += uint64_t(window.size()) * fullStore.size();
        st.passingPairs += sc.nHits;
        if (streamed > 0) resumeAt = fullPos[streamed - 1];  // every live read above this position has had its turn
        ++streamedGrids;
        streamedReads += streamed;
        ++st.streamedGrids;
      } else {
        rc = gridPass(cs, window, fullStore, fullPos);
        if (!rc) rc = gridPass(cs, fresh, incStore, incPos);
      }
      tSpec += nowSec() - tt0;  // listing + GPU + filing (+ the streamed part of the pass)
      gridRange.reset();
      if (g_verboseTiming && nowSec() - tt0 > 0.01)
        fprintf(stderr, "[sdq]   line %d: %zu live, window %u (%zu new), %zu reads x window + %zu x new, spec %.3f s (listing %.3f)%s\n", line,
                alive, wn, fresh.size(), fullStore.size(), incStore.size(), nowSec() - tt0, tList, stream ? " streamed" : "");
      if (rc) return rc;
    }
    tt0 = nowSec();
    NvtxRange passRange("sdq:sequential pass");
    // ---- the sequential pass (smallest cluster first; first passing candidate absorbs it)
    std::vector<uint32_t> odClus(1), odStore, odPos;
    for (size_t rp = resumeAt; rp-- > 0;) {
      if (rp >= 12 && (PR[rp - 12].orMask & bit)) {  // the cluster object a coming turn will touch (its row spans two lines)
        __builtin_prefetch(PR[rp - 12].c);
        __builtin_prefetch(reinterpret_cast<const char*>(PR[rp - 12].c) + 64);
      }
      if (PR[rp].remove) continue;
      if (windowIntact) {
        fastStep(rp);
        continue;
      }
      Clus& read = cs[rp];
      uint32_t count = 0;
      for (uint32_t cp = 0; cp < eligEnd; ++cp) {
        if (PR[cp].remove) continue;
        if (cp == rp) continue;
        ++count;
        if (count > it.stopCheck) break;
        const uint32_t clusStore = PR[cp].store;
        const uint64_t* v = read.row.find(clusStore);
        if (!v && !covered(PR[rp].cover, clusStore)) {
          Clus& clus = cs[cp];
          // a candidate that slid into the window because earlier ones were absorbed: the reads
          // still to come will meet it too, so judge it against all of them in one grid
          odClus[0] = clus.storeIdx;
          odStore.clear();
          odPos.clear();
          for (size_t r2 = rp + 1; r2-- > 0;) {
            if (PR[r2].remove || r2 == cp || covered(PR[r2].cover, clus.storeIdx)) continue;
            odStore.push_back(PR[r2].store);
            odPos.push_back(uint32_t(r2));
          }
          st.onDemandPairs += odStore.size();
          int rc = gridPass(cs, odClus, odStore, odPos);
          if (rc) return rc;
          onDemandAt[clus.storeIdx] = serial;
          v = read.row.find(clusStore);
        }
        ++st.pairsConsumed;
        if (v && (*v & bit)) {
          Clus& clus = cs[cp];
          clus.cnt += read.cnt;
          clus.sumLenCnt += read.sumLenCnt;
          clus.members.append(read.members);
          clus.needCons = true;
          read.remove = true;
          PR[rp].remove = 1;
          PR[cp].changed = 1;
          read.row.clear();
          read.cons.reset();
          if (rp < wn) windowIntact = false;
          ++amountAdded;
          break;
        }
      }
    }
    for (uint32_t s : prevWindow) winPos[s] = 0;  // (prevWindow is this iteration's window by now)
    tPass += nowSec() - tt0;
    if (g_verboseTiming && nowSec() - tt0 > 0.001)
      fprintf(stderr, "[sdq]   line %d: pass %.4f s over %zu positions (from %zu), %zu absorbed, window intact %d, on-demand pairs so far %llu\n", line,
              nowSec() - tt0, PR.size(), resumeAt, amountAdded, int(windowIntact), (unsigned long long)st.onDemandPairs);
    if (amountAdded > 0) {
      orderDirty = true;
      tt0 = nowSec();
      int rc = calculateConsensusBatch(cs);
      tCons += nowSec() - tt0;
      return rc;
    }
    return 0;
  }

  // baseCluster::calculateConsensus for every cluster that absorbed reads, one GPU batch.
  int calculateConsensusBatch(ClusVec& cs) {
    NvtxRange range("sdq:consensus");
    struct Job {
      uint32_t clus;
      size_t firstPair, firstMember;
    };
    std::vector<Job> jobs;
    std::vector<uint32_t> a, b;
    for (uint32_t ci = 0; ci < cs.size(); ++ci) {
      if (PR[ci].remove || !PR[ci].changed) continue;  // (the clusters that absorbed reads in this iteration)
      Clus& cl = cs[ci];
      if (!cl.needCons) continue;
      cl.needCons = false;
      if (cl.members.size() <= 1) continue;
      // (the members' counts add up to the cluster's; both sums are sums of integers, so keeping them up to date at
      //  every absorption gives the very doubles a walk over the members would)
      const double avgLen = cl.sumLenCnt / cl.cnt;
      uint32_t baseStore = cl.storeIdx;
      View baseSeq = cl.seq;
      if (std::fabs(double(cl.seq.size()) - avgLen) > 0.1 * avgLen) {  // restart from the longest member
        size_t biggest = 0;
        for (uint32_t u : cl.members)
          if (uniq[u].seq.size() > biggest) {
            biggest = uniq[u].seq.size();
            baseStore = u;
            baseSeq = uniq[u].seq;
          }
      }
      if (!cl.cons || !baseSeq.equals(cl.cons->baseSeq.data(), cl.cons->baseSeq.size())) {  // new base: counters start over
        cl.cons.reset(new ConsState());
        cl.cons->baseSeq.assign(reinterpret_cast<const char*>(baseSeq.data()), baseSeq.size());
        cl.cons->counters.assign(baseSeq.size(), Col());
      }
      cl.cons->baseStore = baseStore;
      Job j{ci, a.size(), cl.cons->nDone};
      for (size_t m = cl.cons->nDone; m < cl.members.size(); ++m) {
        a.push_back(baseStore);
        b.push_back(cl.members[m]);
      }
      jobs.push_back(j);
    }
    if (jobs.empty()) return 0;
    double tc0 = nowSec();
    // alignments with gap lists; pairs whose list overflowed are re-run with a larger capacity
    const size_t n = a.size();
    uint32_t gapCap = 8;
    std::vector<std::vector<sdgpu_gap>> gaps(n);
    std::vector<uint32_t> todo(n);
    for (size_t i = 0; i < n; ++i) todo[i] = uint32_t(i);
    while (!todo.empty()) {
      std::vector<uint32_t> next;
      const size_t chunk = std::max<size_t>(1024, (BATCH / 4) / (gapCap / 8));
      for (size_t off = 0; off < todo.size(); off += chunk) {
        const uint32_t m = uint32_t(std::min<size_t>(chunk, todo.size() - off));
        std::vector<uint32_t> pa(m), pb(m);
        for (uint32_t i = 0; i < m; ++i) {
          pa[i] = a[todo[off + i]];
          pb[i] = b[todo[off + i]];
        }
        const uint32_t* recs = nullptr;
        uint64_t nWords = 0;
        const double t0 = nowSec();
        int rc = sdgpu_align_gap_lists(ctx_, pa.data(), pb.data(), m, gapCap, &recs, &nWords);  // only pairs with gap runs come back
        st.secondsGpu += nowSec() - t0;
        if (rc) return fail(rc);
        ++st.gpuBatches;
        st.consensusAligns += m;
        for (uint64_t w = 0; w < nWords;) {
          const uint32_t i = recs[w], nRuns = recs[w + 1] & 0x7fffffffu;
          const uint32_t kept = std::min(nRuns, gapCap);
          if (recs[w + 1] >> 31) {
            next.push_back(todo[off + i]);
          } else {
            std::vector<sdgpu_gap>& g = gaps[todo[off + i]];
            g.resize(kept);
            for (uint32_t k = 0; k < kept; ++k) g[k] = sdgpu_gap{recs[w + 2 + 3 * k], recs[w + 3 + 3 * k], recs[w + 4 + 3 * k]};
          }
          w += 2 + 3ull * kept;
        }
      }
      todo.swap(next);
      gapCap *= 8;
    }
    tConsAlign += nowSec() - tc0;
    tc0 = nowSec();
    // column counters -> new consensus (ConsensusHelper::increaseCounters / buildConsensus)
    // (clusters are independent of each other here: a few host threads share the jobs; what must
    //  stay in job order -- the list of new device versions -- is collected afterwards)
    std::vector<uint32_t> changed;
    enum : uint8_t { OUT_SAME = 0, OUT_UPLOAD = 1, OUT_KEPT = 2, OUT_BAD_SYMBOL = 3 };
    std::vector<uint8_t> outcome(jobs.size(), OUT_SAME);
    // (a) counting: members are independent and the sums are integers, so a cluster's new members are cut into
    //     chunks that any thread may take; ungapped members (the rule) are tallied in thread-local columns and
    //     added to the cluster's counters under its lock, gapped ones go through the lock directly
    struct Chunk {
      uint32_t job;
      size_t m0, m1;
    };
    std::vector<Chunk> chunks;
    for (uint32_t jx = 0; jx < jobs.size(); ++jx) {
      const size_t mEnd = cs[jobs[jx].clus].members.size();
      for (size_t m = jobs[jx].firstMember; m < mEnd; m += 512) chunks.push_back(Chunk{jx, m, std::min(mEnd, m + 512)});
    }
    std::vector<std::mutex> jobMu(jobs.size());
    std::vector<uint8_t> jobBad(jobs.size(), 0);
    parallelForDynamic(chunks.size(), [&](size_t cx) {
      const Chunk& ch = chunks[cx];
      const Job& j = jobs[ch.job];
      struct InsRec {
        uint32_t pos, offset;  // in front of base position `pos`, the offset-th inserted base (1-based)
        uint8_t sym, qual;
        uint32_t w;
        uint8_t leading;       // part of a gap at the very start of the alignment (ConsensusHelper's beginningGap)
      };
      std::vector<InsRec> ins;
      // Ungapped members differ from the base sequence in a few columns only.  Per chunk: wAll = their total weight,
      // qAll[i] = sum of quality x weight over ALL of them (one multiply-add sweep per member); the columns where a
      // member does not carry the base's symbol are tallied symbol by symbol in `local` and remembered in wOff / qOff,
      // so that the base symbol of column i receives wAll - wOff[i] and qAll[i] - qOff[i].  Integer sums: exact.
      std::vector<Col> local;
      std::vector<uint32_t> qAll, wOff, qOff;
      uint32_t wAll = 0;
      bool symOk = true;
      Clus& cl = cs[j.clus];
      ConsState& S = *cl.cons;
      const std::string& base = S.baseSeq;
      const size_t L = base.size();
      for (size_t m = ch.m0; m < ch.m1; ++m) {
        if (m + 3 < ch.m1) {  // members lie all over the caller's read buffers: have a later one's bytes on their way
          const Uniq& nx = uniq[cl.members[m + 3]];
          for (uint32_t o = 0; o < nx.seq.n; o += 64) {
            __builtin_prefetch(nx.seq.data() + o);
            __builtin_prefetch(nx.qual.data() + o);
          }
        }
        const Uniq& r = uniq[cl.members[m]];
        const uint32_t w = uint32_t(std::llround(r.cnt));
        const std::vector<sdgpu_gap>& gl = gaps[j.firstPair + (m - j.firstMember)];
        if (gl.empty() && r.seq.size() == L) {  // ungapped: column i is base position i
          if (local.empty()) {
            local.assign(L, Col());
            qAll.assign(L, 0);
            wOff.assign(L, 0);
            qOff.assign(L, 0);
          }
          wAll += w;
          const uint8_t* q = r.qual.data();
          const uint8_t* sq = r.seq.data();
          uint32_t* qa = qAll.data();
          for (size_t i = 0; i < L; ++i) qa[i] += uint32_t(q[i]) * w;
          size_t i = 0;
          for (; i + 8 <= L; i += 8) {  // eight columns at a time: where does the member leave the base sequence?
            uint64_t x, y;
            std::memcpy(&x, sq + i, 8);
            std::memcpy(&y, base.data() + i, 8);
            if (x == y) continue;
            for (size_t k = i; k < i + 8; ++k)
              if (sq[k] != uint8_t(base[k])) {
                symOk &= add(local[k], char(sq[k]), q[k], w);
                wOff[k] += w;
                qOff[k] += uint32_t(q[k]) * w;
              }
          }
          for (; i < L; ++i)
            if (sq[i] != uint8_t(base[i])) {
              symOk &= add(local[i], char(sq[i]), q[i], w);
              wOff[i] += w;
              qOff[i] += uint32_t(q[i]) * w;
            }
          continue;
        }
        // gapped member: walk the alignment its gap list spells (the list is in traceback order: last entry = first
        // gap).  Columns that carry a base of the base sequence are tallied in `local` like the mismatches above (a '-'
        // of the member counts as symbol '-' with quality 0); bases the member inserts between two base positions go to
        // the cluster's insertion maps, collected here and filed under the cluster's lock when the chunk is done.
        if (local.empty()) {
          local.assign(L, Col());
          qAll.assign(L, 0);
          wOff.assign(L, 0);
          qOff.assign(L, 0);
        }
        const uint8_t* q = r.qual.data();
        const uint8_t* sq = r.seq.data();
        const uint32_t lr = uint32_t(r.seq.size());
        uint32_t pa = 0, pb = 0;
        auto diagonal = [&](uint32_t upto, bool inA) {
          while ((inA ? pa : pb) < upto && pa < L && pb < lr) {
            symOk &= add(local[pa], char(sq[pb]), q[pb], w);
            ++pa;
            ++pb;
          }
        };
        for (size_t g = gl.size(); g-- > 0;) {
          const sdgpu_gap& gp = gl[g];
          diagonal(gp.pos, gp.gapInA != 0);
          if (gp.gapInA) {  // '-' in the base sequence: the member inserts gp.size bases in front of base position pa
            for (uint32_t x = 0; x < gp.size && pb < lr; ++x, ++pb)
              ins.push_back(InsRec{pa, x + 1, sq[pb], q[pb], w, uint8_t(pa == 0 && pb == x)});
          } else {          // '-' in the member
            for (uint32_t x = 0; x < gp.size && pa < L; ++x, ++pa) symOk &= add(local[pa], '-', 0, w);
          }
        }
        diagonal(uint32_t(L), true);
      }
      if (!local.empty() || !ins.empty()) {
        std::lock_guard<std::mutex> lk(jobMu[ch.job]);
        for (const InsRec& x : ins) {
          if (x.leading) {
            symOk &= add(S.beginningGap[x.offset - 1], char(x.sym), x.qual, x.w);
          } else {
            symOk &= add(S.insertions[x.pos][x.offset], char(x.sym), x.qual, x.w);
          }
        }
        for (size_t i = 0; i < local.size(); ++i) {
          for (int x = 0; x < NSYM; ++x) {
            S.counters[i].cnt[x] += local[i].cnt[x];
            S.counters[i].qual[x] += local[i].qual[x];
          }
          const int bs = slot(base[i]);  // (registered: the base sequence is in the store)
          if (bs < 0) {
            symOk = false;
            continue;
          }
          S.counters[i].cnt[bs] += wAll - wOff[i];
          S.counters[i].qual[bs] += qAll[i] - qOff[i];
        }
      }
      if (!symOk) jobBad[ch.job] = 1;
    }, 8, 2);
    tConsCount += nowSec() - tc0;
    tc0 = nowSec();
    // (b) the new consensus of every cluster, clusters side by side
    parallelForDynamic(jobs.size(), [&](size_t jx) {
      const Job& j = jobs[jx];
      const bool symOk = !jobBad[jx];
      Clus& cl = cs[j.clus];
      ConsState& S = *cl.cons;
      S.nDone = cl.members.size();
      if (!symOk) {
        outcome[jx] = OUT_BAD_SYMBOL;
        return;
      }
      std::string ns;
      std::vector<uint8_t> nq;
      const double fortyPercent = 0.40 * cl.cnt;
      const uint32_t size = uint32_t(std::llround(cl.cnt));
      char bch;
      uint32_t q, bc;
      for (const auto& bg : S.beginningGap) {
        best(bg.second, bch, q, bc);
        const uint32_t tot = bg.second.total();
        if (bch == '-' || tot < fortyPercent) continue;
        ns.push_back(bch);
        nq.push_back(uint8_t(q / tot));
      }
      for (uint32_t pos = 0; pos < S.counters.size(); ++pos) {
        if (!S.insertions.empty()) {
          auto ins = S.insertions.find(pos);
          if (ins != S.insertions.end()) {
            for (const auto& ci : ins->second) {
              best(ci.second, bch, q, bc);
              const uint32_t tot = ci.second.total();
              const uint32_t nonInserting = size > tot ? size - tot : 0;
              if (bch == ' ' || bc <= nonInserting) continue;
              ns.push_back(bch);
              nq.push_back(uint8_t(q / bc));
            }
          }
        }
        best(S.counters[pos], bch, q, bc);
        const uint32_t tot = S.counters[pos].total();
        if (bch == '-' || bch == ' ' || tot < fortyPercent) continue;
        ns.push_back(bch);
        nq.push_back(uint8_t(q / tot));
      }
      const bool seqChanged = !cl.seq.equals(ns.data(), ns.size());
      if (seqChanged || !cl.qual.equals(nq.data(), nq.size())) {
        std::vector<uint8_t> newSig;
        qualSignature(nq, P, newSig);
        bool needUpload = seqChanged;
        if (!seqChanged) {  // same sequence: a new device version is needed only if checkQual flips somewhere
          if (cl.devSig.empty()) qualSignature(cl.qual, P, cl.devSig);
          needUpload = newSig != cl.devSig;
        }
        cl.ownSeq.swap(ns);
        cl.ownQual.swap(nq);
        cl.seq = View{reinterpret_cast<const uint8_t*>(cl.ownSeq.data()), uint32_t(cl.ownSeq.size())};
        cl.qual = View{cl.ownQual.data(), uint32_t(cl.ownQual.size())};
        if (needUpload) {
          outcome[jx] = OUT_UPLOAD;
          cl.devSig.swap(newSig);
        } else {
          outcome[jx] = OUT_KEPT;
        }
      }
      cl.avgErr = averageErrorRate(cl.qual);
    }, 4, 32);
    tConsBuild += nowSec() - tc0;
    tc0 = nowSec();
    struct UploadTick {
      double& acc;
      double t0;
      ~UploadTick() { acc += nowSec() - t0; }
    } uploadTick{tConsUpload, tc0};
    for (size_t jx = 0; jx < jobs.size(); ++jx) {
      if (outcome[jx] == OUT_BAD_SYMBOL) {
        g_err = "more than 18 distinct symbols in consensus columns";
        return SDGPU_E_ALPHABET;
      }
      if (outcome[jx] == OUT_UPLOAD) changed.push_back(jobs[jx].clus);
      if (outcome[jx] == OUT_KEPT) ++keptVersions;
    }
    // new consensus versions -> device store (append-only; a store index never changes meaning)
    if (!changed.empty()) {
      std::vector<uint8_t> bases, quals;
      std::vector<uint64_t> offs{0};
      std::vector<double> cnts;
      for (uint32_t ci : changed) {
        const Clus& cl = cs[ci];
        if (cl.seq.empty()) {
          g_err = "consensus collapsed to an empty sequence";
          return SDGPU_E_UNSUPPORTED;
        }
        bases.insert(bases.end(), cl.seq.begin(), cl.seq.end());
        quals.insert(quals.end(), cl.qual.begin(), cl.qual.end());
        offs.push_back(bases.size());
        cnts.push_back(cl.cnt);
      }
      uint32_t first = 0;
      int rc = sdgpu_append_seqs(ctx_, bases.data(), quals.data(), offs.data(), cnts.data(), uint32_t(changed.size()), &first);
      if (rc) return fail(rc);
      for (size_t i = 0; i < changed.size(); ++i) {
        Clus& cl = cs[changed[i]];
        cl.storeIdx = first + uint32_t(i);
        cl.row.clear();  // its verdicts as a read were for the old version
        PR[changed[i]].cover = 0;
        PR[changed[i]].orMask = 0;
        PR[changed[i]].store = cl.storeIdx;
      }
      storeSize = first + uint32_t(changed.size());
    }
    return 0;
  }

  // drop absorbed clusters, then readVecSorter "totalCount".  A Clus is heavy, so the list holds pointers and the
  // sort works on the compact records of P.  Between two sorts only a few records move -- the clusters that absorbed
  // reads (new count / average error) and the absorbed ones (gone) -- so the usual call compacts the sorted records,
  // sorts the few changed ones and merges the two runs; the order is total (first id breaks every tie), hence the
  // result is std::sort's.
  static bool posLess(const Pos& a, const Pos& b) {
    if (a.cnt != b.cnt) return a.cnt > b.cnt;
    if (a.avgErr != b.avgErr) return a.avgErr < b.avgErr;
    return clusLess(*a.c, *b.c);
  }
  void dropRemovedAndSort(ClusVec& cs) {
    NvtxRange range("sdq:sort");
    const double t0 = nowSec();
    std::vector<Clus*>& v = cs.ptrs();
    if (posValid && PR.size() == v.size()) {
      std::vector<Pos> moved;
      size_t w = 0;
      for (size_t i = 0; i < PR.size(); ++i) {
        if (PR[i].remove) continue;
        if (PR[i].changed) {
          Pos x = PR[i];
          x.cnt = x.c->cnt;
          x.avgErr = x.c->avgErr;
          x.changed = 0;
          moved.push_back(x);
          continue;
        }
        PR[w++] = PR[i];
      }
      PR.resize(w);
      if (!moved.empty()) {
        std::sort(moved.begin(), moved.end(), posLess);
        std::vector<Pos> out(PR.size() + moved.size());
        std::merge(PR.begin(), PR.end(), moved.begin(), moved.end(), out.begin(), posLess);
        PR.swap(out);
      }
    } else {  // a new list (start of a clustering run): build the records from the clusters and sort them all
      v.erase(std::remove_if(v.begin(), v.end(), [](const Clus* c) { return c->remove; }), v.end());
      PR.resize(v.size());
      parallelFor(v.size(), [&](size_t lo, size_t hi) {  // (one cache miss per cluster object: worth a few threads at 178 k clusters)
        for (size_t i = lo; i < hi; ++i) PR[i] = Pos{v[i]->cnt, v[i]->avgErr, v[i], 0, v[i]->row.orMask(), v[i]->storeIdx, 0, 0};
      }, 4096);
      // (the singletons joining the survivors of the first run: two sorted runs back to back -> one merge)
      const auto firstDescent = std::is_sorted_until(PR.begin(), PR.end(), posLess);
      if (firstDescent != PR.end() && std::is_sorted(firstDescent, PR.end(), posLess)) {
        std::inplace_merge(PR.begin(), firstDescent, PR.end(), posLess);
      } else if (firstDescent != PR.end()) {
        const size_t n = PR.size();
        const size_t T = n >= (1u << 15) ? std::min<size_t>(allowedCpus(), 4) : 1;
        if (T > 1) {
          std::vector<size_t> cut(T + 1);
          for (size_t t = 0; t <= T; ++t) cut[t] = n * t / T;
          parallelFor(T, [&](size_t lo, size_t hi) {
            for (size_t t = lo; t < hi; ++t) std::sort(PR.begin() + cut[t], PR.begin() + cut[t + 1], posLess);
          }, 1);
          for (size_t width = 1; width < T; width *= 2)
            for (size_t t = 0; t + width < T; t += 2 * width)
              std::inplace_merge(PR.begin() + cut[t], PR.begin() + cut[t + width], PR.begin() + cut[std::min(T, t + 2 * width)], posLess);
        } else {
          std::sort(PR.begin(), PR.end(), posLess);
        }
      }
    }
    v.resize(PR.size());
    for (size_t i = 0; i < PR.size(); ++i) v[i] = PR[i].c;
    posValid = true;
    tSort += nowSec() - t0;
  }

  // njhseq collapser::runFullClustering: with kmerBinOpts_.useKmerBinning_ (clusterDownSetUp.cpp:352-354) the clusters
  // are first binned by k-mer similarity and every bin is collapsed on its own with binIteratorMap (clusterDown.cpp:489).
  int runFullClustering(ClusVec& cs, const sdq_iter_par* its, uint32_t nIts) {
    if (O.useKmerBinning) {
      int rc = binAndCollapse(cs, its, nIts);
      if (rc) return rc;
    }
    return runClustering(cs, its, nIts);
  }

  // Greedy k-mer binning (njhseq greedyKmerSimCluster / kmerCluster::compareRead): in list order a cluster joins the
  // first bin whose seed shares more than kmerCutOff of its k-mers, else it opens a bin.  The order dependence only concerns
  // which clusters are seeds, so a round hands the GPU the next <= 64 unbinned clusters as candidate seeds against every
  // unbinned cluster (sdgpu_kmer_bin_masks: one 64-bit word per cluster comes back) and settles it here: candidate j is a
  // seed iff none of this round's earlier seeds passes it (bins of earlier rounds have already refused everything that is
  // still unbinned); the others join the earliest seed that passes them or wait for the next round.
  int binAndCollapse(ClusVec& cs, const sdq_iter_par* its, uint32_t nIts) {
    NvtxRange range("sdq:k-mer binning");
    const double t0 = nowSec();
    int rc = sdgpu_build_kmer_lists(ctx_, O.kCompareLen);
    if (rc) return fail(rc);
    std::vector<Clus*>& v = cs.ptrs();
    v.erase(std::remove_if(v.begin(), v.end(), [](const Clus* c) { return c->remove; }), v.end());
    std::vector<uint32_t> unb(v.size()), rest, seeds, readStore;
    for (uint32_t i = 0; i < unb.size(); ++i) unb[i] = i;
    std::vector<uint64_t> masks;
    std::vector<std::vector<Clus*>> bins;
    while (!unb.empty()) {
      const uint32_t nc = uint32_t(std::min<size_t>(64, unb.size()));
      seeds.resize(nc);
      readStore.resize(unb.size());
      masks.resize(unb.size());
      for (size_t x = 0; x < unb.size(); ++x) readStore[x] = v[unb[x]]->storeIdx;
      for (uint32_t j = 0; j < nc; ++j) seeds[j] = readStore[j];
      rc = sdgpu_kmer_bin_masks(ctx_, seeds.data(), nc, readStore.data(), uint32_t(readStore.size()), O.kmerCutOff, masks.data(), nullptr);
      if (rc) return fail(rc);
      st.kmerPairs += uint64_t(nc) * readStore.size();
      uint64_t actual = 0;  // this round's candidates that became seeds (only bits below x while x is a candidate)
      uint32_t binOfBit[64];
      rest.clear();
      for (size_t x = 0; x < unb.size(); ++x) {
        const uint64_t m = masks[x] & actual;
        if (m) {
          bins[binOfBit[__builtin_ctzll(m)]].push_back(v[unb[x]]);
        } else if (x < nc) {
          actual |= 1ull << x;
          binOfBit[x] = uint32_t(bins.size());
          bins.emplace_back(1, v[unb[x]]);
        } else {
          rest.push_back(unb[x]);
        }
      }
      unb.swap(rest);
    }
    st.kmerBins += bins.size();
    tBin += nowSec() - t0;
    const sdq_iter_par* binIts = O.binPars ? O.binPars : its;
    const uint32_t nBinIts = O.binPars ? uint32_t(O.nBinPars) : nIts;
    std::vector<Clus*> survivors;
    for (auto& bin : bins) {
      if (bin.size() > 1) {
        ClusVec b;
        b.ptrs().swap(bin);
        if ((rc = runClustering(b, binIts, nBinIts))) return rc;
        bin.swap(b.ptrs());
      }
      survivors.insert(survivors.end(), bin.begin(), bin.end());
    }
    v.swap(survivors);
    return 0;
  }

  int runClustering(ClusVec& cs, const sdq_iter_par* its, uint32_t nIts) {
    orderDirty = true;
    posValid = false;  // the list may have been re-composed (singletons joining)
    for (uint32_t i = 0; i < nIts; ++i) {
      if (orderDirty) dropRemovedAndSort(cs);  // counts / members only change when something merged
      orderDirty = false;
      if (cs.size() < 2) break;
      int rc = collapseWithParameters(cs, its[i], lineOf(its[i]));
      if (rc) return rc;
      // --converge: the same line again, on the re-sorted survivors, while it still merges something
      while (O.converge && orderDirty) {
        dropRemovedAndSort(cs);
        orderDirty = false;
        if (cs.size() < 2) break;
        if ((rc = collapseWithParameters(cs, its[i], lineOf(its[i])))) return rc;
      }
    }
    dropRemovedAndSort(cs);
    return 0;
  }

  int buildIndex(ClusVec& cs) {  // indexKmers over the current clusters
    NvtxRange range("sdq:indexKmers");
    const double t0 = nowSec();
    struct Tick {
      double& acc;
      double t0;
      ~Tick() { acc += nowSec() - t0; }
    } tick{tIndex, t0};
    std::vector<uint32_t> idx(cs.size());
    std::vector<double> cnts(cs.size());
    // store order (sequential reads on the device): store indices are unique, so one pass over a
    // store-sized table replaces a sort
    std::vector<double> byStore(storeSize, -1.0);
    for (uint32_t i = 0; i < cs.size(); ++i) byStore[cs[i].storeIdx] = cs[i].cnt;
    uint32_t m = 0;
    for (uint32_t sIdx = 0; sIdx < storeSize; ++sIdx)
      if (byStore[sIdx] >= 0) {
        idx[m] = sIdx;
        cnts[m++] = byStore[sIdx];
      }
    const double tA = nowSec();
    int rc = sdgpu_update_cnts(ctx_, idx.data(), cnts.data(), uint32_t(idx.size()));
    if (rc) return fail(rc);
    const double tB = nowSec();
    rc = sdgpu_build_kmer_index(ctx_, idx.data(), uint32_t(idx.size()), O.kLen, O.runCutOff, int(O.kmersByPosition),
                                int(O.expandKmerPos), O.expandKmerSize);
    if (rc) return fail(rc);
    const double tC = nowSec();
    if (g_verboseTiming) fprintf(stderr, "[sdq]   index: gather+sort %.3f update_cnts %.3f build %.3f\n", tA - t0, tB - tA, tC - tB);
    parallelFor(cs.size(), [&](size_t lo, size_t hi) {  // verdicts depend on the index through lowKmerMismatches
      for (size_t i = lo; i < hi; ++i) {
        cs[i].row.clear();
      }
    });
    for (Pos& x : PR) x.cover = 0, x.orMask = 0;
    return 0;
  }
};

}  // namespace

// collapser::markChimeras(clusters, alignerObj, chiOpts_) (clusterDown.cpp:566-577) on the final,
// count-sorted clusters.  Which (parent, candidate) pairs are examined depends only on the counts,
// so all probes go to the GPU as one batch; the crossover search over the returned 32-byte
// records stays here.
static int markChimeras(sdgpu_ctx* ctx, const sdq_opts& O, ClusVec& cs, sdq_stats& st) {
  const size_t n = cs.size();
  if (n <= 2) return SDGPU_OK;
  std::vector<uint32_t> parentIdx, candIdx;
  std::vector<uint32_t> firstPair(n + 1, 0);
  for (size_t sub = 2; sub < n; ++sub) {
    firstPair[sub] = uint32_t(parentIdx.size());
    if (cs[sub].cnt <= double(O.chiRunCutOff)) continue;
    for (size_t par = 0; par < sub; ++par) {
      if (cs[par].cnt / cs[sub].cnt < O.chiParentFreqs) continue;
      parentIdx.push_back(cs[par].storeIdx);
      candIdx.push_back(cs[sub].storeIdx);
    }
  }
  firstPair[n] = uint32_t(parentIdx.size());
  for (size_t sub = 0; sub < 2; ++sub) firstPair[sub] = 0;
  st.chimeraPairs = parentIdx.size();
  if (parentIdx.empty()) return SDGPU_OK;
  std::vector<sdgpu_chimera_probe> probes(parentIdx.size());
  // profileAlignment(parent, candidate, false, true, false)
  if (sdgpu_chimera_probe_pairs(ctx, parentIdx.data(), candIdx.data(), uint32_t(parentIdx.size()), SDGPU_USING_QUALITY,
                                &O.chiOverlap, int(O.chiKeepLowQual), probes.data()))
    return SDGPU_E_CUDA;
  std::vector<std::pair<uint32_t, uint32_t>> front, back;  // (candidate base position, parent)
  for (size_t sub = 2; sub < n; ++sub) {
    Clus& cand = cs[sub];
    if (cand.cnt <= double(O.chiRunCutOff)) continue;
    front.clear();
    back.clear();
    const uint32_t lb = uint32_t(cand.seq.size());
    uint32_t x = firstPair[sub];
    for (size_t par = 0; par < sub; ++par) {
      if (cs[par].cnt / cand.cnt < O.chiParentFreqs) continue;
      const sdgpu_chimera_probe& pr = probes[x++];
      if (pr.keptMismatches == 0) continue;
      if ((pr.flags & SDGPU_CHI_FRONT_PASS) && pr.firstPos >= O.chiOverlapSizeCutoff) front.emplace_back(pr.firstPos, uint32_t(par));
      if ((pr.flags & SDGPU_CHI_BACK_PASS) && lb - 1 - pr.lastPos >= O.chiOverlapSizeCutoff) back.emplace_back(pr.lastPos, uint32_t(par));
    }
    // multimap order: by position, insertion order among equal positions
    auto byPos = [](const std::pair<uint32_t, uint32_t>& a, const std::pair<uint32_t, uint32_t>& b) { return a.first < b.first; };
    std::stable_sort(front.begin(), front.end(), byPos);
    std::stable_sort(back.begin(), back.end(), byPos);
    for (const auto& f : front) {
      for (const auto& b : back) {
        if (cand.chimeric) break;
        if (f.second != b.second && f.first > b.first && f.first - b.first >= O.chiPosSpacing) {
          cand.chimeric = true;
          cand.chiParent[0] = f.second, cand.chiPos[0] = f.first;
          cand.chiParent[1] = b.second, cand.chiPos[1] = b.first;
          ++st.chimericClusters;
        }
      }
    }
  }
  return SDGPU_OK;
}

struct sdq_result {
  std::vector<Clus> pool;  // storage of every cluster object of the run
  ClusVec clusters;        // the final clusters, sorted
  std::vector<uint32_t> membership;
  sdq_stats stats;
};

extern "C" {

const char* sdq_last_error(void) { return g_err.c_str(); }

}  // extern "C"

namespace {
// The collapse's input: packed reads (sdq_run) or one pointer per sequence with a count each (population
// clustering: the final clusters of many samples, which are not deduplicated).
struct ReadSrc {
  const uint8_t* bases = nullptr;
  const uint8_t* quals = nullptr;
  const uint64_t* offsets = nullptr;
  const uint8_t* const* seqPtrs = nullptr;
  const uint8_t* const* qualPtrs = nullptr;
  const uint32_t* lens = nullptr;
  const double* cnts = nullptr;  // with seqPtrs: seqBase_.cnt_ of every input cluster
  const uint8_t* seq(uint32_t r) const { return seqPtrs ? seqPtrs[r] : bases + offsets[r]; }
  const uint8_t* qual(uint32_t r) const { return seqPtrs ? qualPtrs[r] : quals + offsets[r]; }
  uint32_t len(uint32_t r) const { return seqPtrs ? lens[r] : uint32_t(offsets[r + 1] - offsets[r]); }
};
}  // namespace

static int runCollapse(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit, const sdq_iter_par* pars,
                       uint32_t nPars, const ReadSrc& in, uint32_t n, sdq_result** out) {
  const bool clustersIn = in.seqPtrs != nullptr;  // every input is a cluster of its own, with its count
  *out = nullptr;
  std::unique_ptr<NvtxRange> prepRange(new NvtxRange("sdq:prep (dedupe, medians, upload)"));
  const double tStart = nowSec();
  Runner R(ctx, *opts);
  auto res = new sdq_result();
  std::thread poolThread;  // constructs res->pool while the main thread goes on (see below)
  struct Joiner {
    std::thread& t;
    ~Joiner() {
      if (t.joinable()) t.join();
    }
  } poolJoiner{poolThread};
  auto bail = [&](int rc) {
    if (poolThread.joinable()) poolThread.join();
    delete res;
    return rc;
  };
  res->membership.assign(n, UINT32_MAX);
  uint64_t launches0 = 0, cells0 = 0;
  if (sdgpu_counters(ctx, &launches0, &cells0)) return bail(R.fail(SDGPU_E_CUDA));
  if (sdgpu_get_params(ctx, &R.P)) return bail(R.fail(SDGPU_E_CUDA));
  // every distinct par-file line becomes one bit of the verdict mask
  for (uint32_t i = 0; i < nInit; ++i) R.lineOf(initPars[i]);
  for (uint32_t i = 0; i < nPars; ++i) R.lineOf(pars[i]);
  if (opts->useKmerBinning) {
    if (opts->kCompareLen == 0 || opts->kCompareLen > 16) {
      g_err = "useKmerBinning: kCompareLen must be 1..16";
      return bail(SDGPU_E_ARG);
    }
    if (opts->nBinPars && !opts->binPars) {
      g_err = "useKmerBinning: nBinPars without binPars";
      return bail(SDGPU_E_ARG);
    }
    for (uint64_t i = 0; opts->binPars && i < opts->nBinPars; ++i) R.lineOf(opts->binPars[i]);
  }
  if (R.lines.size() > SDGPU_MAX_PASS_LINES) {
    g_err = "more than 64 distinct par-file lines";
    return bail(SDGPU_E_UNSUPPORTED);
  }
  if (sdgpu_set_pass_table(ctx, R.lines.data(), uint32_t(R.lines.size()))) return bail(R.fail(SDGPU_E_CUDA));
  // exact-sequence dedupe in first-seen order (clusterDown.cpp:272-288; hashing instead of the
  // reference's O(N*U) scan, same result)
  {
    // 64-bit hash of the bases, all reads in parallel; open addressing, equal hashes confirmed byte for byte
    std::vector<uint64_t> readHash(n);
    parallelFor(n, [&](size_t lo, size_t hi) {
      for (size_t r = lo; r < hi; ++r) {
        const uint32_t len = in.len(uint32_t(r));
        const uint8_t* p = in.seq(uint32_t(r));
        uint64_t h = 0x9e3779b97f4a7c15ull ^ len;
        uint32_t i = 0;
        for (; i + 8 <= len; i += 8) {
          uint64_t w;
          std::memcpy(&w, p + i, 8);
          h = (h ^ w) * 0xff51afd7ed558ccdull;
          h ^= h >> 32;
        }
        for (; i < len; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
        readHash[r] = h ^ (h >> 29);
      }
    });
    size_t cap = 1024;
    while (cap < size_t(n) * 2 + 16) cap <<= 1;
    std::vector<uint32_t> table(cap, UINT32_MAX);
    std::vector<uint32_t> firstRead;  // unique -> its first read
    std::vector<uint32_t> uniqOfRead(n, UINT32_MAX);
    firstRead.reserve(n);
    R.uniq.reserve(n);
    for (uint32_t r = 0; r < n; ++r) {
      const uint32_t len = in.len(r);
      if (len <= opts->smallReadSize) continue;
      if (clustersIn) {  // population input: no dedupe, the cluster brings its count
        uniqOfRead[r] = uint32_t(R.uniq.size());
        firstRead.push_back(r);
        R.uniq.emplace_back();
        R.uniq.back().cnt = in.cnts ? in.cnts[r] : 1.0;
        continue;
      }
      if (r + 8 < n) __builtin_prefetch(&table[size_t(readHash[r + 8]) & (cap - 1)]);
      const uint8_t* p = in.seq(r);
      const uint64_t h = readHash[r];
      size_t slot = size_t(h) & (cap - 1);
      uint32_t found = UINT32_MAX;
      while (table[slot] != UINT32_MAX) {
        const uint32_t u = table[slot];
        const uint32_t fr = firstRead[u];
        if (readHash[fr] == h && in.len(fr) == len && std::memcmp(in.seq(fr), p, len) == 0) {
          found = u;
          break;
        }
        slot = (slot + 1) & (cap - 1);
      }
      if (found == UINT32_MAX) {
        found = uint32_t(R.uniq.size());
        table[slot] = found;
        firstRead.push_back(r);
        R.uniq.emplace_back();
      }
      R.uniq[found].cnt += 1.0;
      uniqOfRead[r] = found;
    }
    // the input reads of every unique sequence, ascending (a counting sort of the reads by unique id)
    const uint32_t nUq = uint32_t(R.uniq.size());
    uint32_t run = 0;
    uint64_t arenaBytes = 0;
    for (uint32_t u = 0; u < nUq; ++u) {
      Uniq& q = R.uniq[u];
      q.inputBegin = q.inputEnd = run;
      run += clustersIn ? 1u : uint32_t(q.cnt);
      const uint32_t fr = firstRead[u];
      q.seq = View{in.seq(fr), in.len(fr)};
      q.qual = View{in.qual(fr), q.seq.n};  // (several reads: replaced by the per-base median below)
      if (!clustersIn && q.cnt > 1.0) arenaBytes += q.seq.n;
    }
    R.inputList.resize(run);
    for (uint32_t r = 0; r < n; ++r)
      if (uniqOfRead[r] != UINT32_MAX) R.inputList[R.uniq[uniqOfRead[r]].inputEnd++] = r;
    R.qualArena.resize(arenaBytes);
    uint64_t at = 0;
    for (uint32_t u = 0; u < nUq && !clustersIn; ++u)
      if (R.uniq[u].cnt > 1.0) {
        R.uniq[u].qual = View{R.qualArena.data() + at, R.uniq[u].seq.n};
        at += R.uniq[u].seq.n;
      }
  }
  const double tDedupe = nowSec() - tStart;
  // the cluster objects (one per unique sequence, never resized again: ClusVec holds pointers into the pool) are
  // constructed on a helper thread while this one computes medians and uploads
  poolThread = std::thread([res, nPool = R.uniq.size()] { res->pool.resize(nPool); });
  // per-base median quality over the reads of a unique sequence (clusterDown.cpp:293-359, qualRep "median"); a
  // sequence seen once keeps its read's qualities.  (abundant sequences come first and cost the most: blocks of 256
  // uniques are handed out dynamically)
  parallelForDynamic((R.uniq.size() + 255) / 256, [&](size_t blk) {
    std::vector<uint8_t> col;
    std::vector<uint32_t> hist;
    const size_t lo = blk * 256, hi = std::min(R.uniq.size(), lo + 256);
    for (size_t x = lo; x < hi; ++x) {
      Uniq& u = R.uniq[x];
      const size_t len = u.seq.size(), k = size_t(u.inputEnd - u.inputBegin);
      if (k <= 1) continue;
      const uint32_t* members = R.inputList.data() + u.inputBegin;
      uint8_t* out = const_cast<uint8_t*>(u.qual.data());  // (its slice of the arena)
      // two, three or four reads (most of the multi-read sequences): the median row by row, no per-column sort
      if (k == 2) {
        const uint8_t *a = in.qual(members[0]), *b = in.qual(members[1]);
        for (size_t i = 0; i < len; ++i) out[i] = uint8_t((uint32_t(a[i]) + b[i]) / 2);
        continue;
      }
      if (k == 3) {
        const uint8_t *a = in.qual(members[0]), *b = in.qual(members[1]), *c = in.qual(members[2]);
        for (size_t i = 0; i < len; ++i) {
          const uint8_t lo = std::min(a[i], b[i]), hi = std::max(a[i], b[i]);
          out[i] = std::max(lo, std::min(hi, c[i]));
        }
        continue;
      }
      if (k == 4) {  // mean of the two middle values = (sum - smallest - largest) / 2
        const uint8_t *a = in.qual(members[0]), *b = in.qual(members[1]), *c = in.qual(members[2]), *d = in.qual(members[3]);
        for (size_t i = 0; i < len; ++i) {
          const uint32_t lo = std::min(std::min(a[i], b[i]), std::min(c[i], d[i]));
          const uint32_t hi = std::max(std::max(a[i], b[i]), std::max(c[i], d[i]));
          out[i] = uint8_t((uint32_t(a[i]) + b[i] + c[i] + d[i] - lo - hi) / 2);
        }
        continue;
      }
      if (k < 16) {
        col.resize(k);
        for (size_t i = 0; i < len; ++i) {
          for (size_t m = 0; m < k; ++m) col[m] = in.qual(members[m])[i];
          std::sort(col.begin(), col.end());
          out[i] = (k % 2) ? col[k / 2] : uint8_t((uint32_t(col[k / 2 - 1]) + col[k / 2]) / 2);
        }
        continue;
      }
      // many members: per-column histograms filled member by member (sequential reads), medians read off them
      hist.assign(len * 256, 0);
      for (size_t m = 0; m < k; ++m) {
        const uint8_t* q = in.qual(members[m]);
        for (size_t i = 0; i < len; ++i) ++hist[i * 256 + q[i]];
      }
      const size_t r1 = (k % 2) ? k / 2 : k / 2 - 1, r2 = k / 2;  // order statistics (0-based) the median needs
      for (size_t i = 0; i < len; ++i) {
        const uint32_t* h = hist.data() + i * 256;
        size_t seen = 0;
        uint32_t v1 = 0, v2 = 0;
        bool got1 = false;
        for (uint32_t v = 0; v < 256; ++v) {
          seen += h[v];
          if (!got1 && seen > r1) {
            v1 = v;
            got1 = true;
          }
          if (seen > r2) {
            v2 = v;
            break;
          }
        }
        out[i] = (k % 2) ? uint8_t(v1) : uint8_t((v1 + v2) / 2);
      }
    }
  }, 8);
  const uint32_t nU = uint32_t(R.uniq.size());
  res->stats.inputReads = n;
  res->stats.uniqueSeqs = nU;
  if (nU == 0) {
    res->stats.secondsTotal = nowSec() - tStart;
    *out = res;
    return SDGPU_OK;
  }
  const double tMedian = nowSec() - tStart;
  // the unique reads become sequence-store entries 0..nU-1, translated straight from where they lie
  {
    std::vector<const uint8_t*> sp(nU), qp(nU);
    std::vector<uint32_t> ln(nU);
    std::vector<double> cn(nU);
    for (uint32_t u = 0; u < nU; ++u) {
      sp[u] = R.uniq[u].seq.data();
      qp[u] = R.uniq[u].qual.data();
      ln[u] = R.uniq[u].seq.n;
      cn[u] = R.uniq[u].cnt;
    }
    if (sdgpu_upload_seqs_gather(ctx, sp.data(), qp.data(), ln.data(), cn.data(), nU)) return bail(R.fail(SDGPU_E_CUDA));
    R.storeSize = nU;
    if (sdgpu_timer_start(ctx)) return bail(R.fail(SDGPU_E_CUDA));  // inputs are resident from here on
    // every symbol the store knows gets its consensus-counter slot now (the upload has just walked every base), so the
    // threaded consensus rebuild only reads the slot table
    uint8_t chars[16];
    uint32_t nChars = 0;
    if (sdgpu_get_alphabet(ctx, chars, &nChars)) return bail(R.fail(SDGPU_E_CUDA));
    R.slot('-');
    for (uint32_t c = 0; c < nChars; ++c) R.slot(char(chars[c]));
  }
  const double tUpload = nowSec() - tStart;
  poolThread.join();
  ClusVec clusters;
  for (uint32_t u = 0; u < nU; ++u) clusters.push_back(&res->pool[u]);
  parallelFor(nU, [&](size_t lo, size_t hi) {
    for (size_t x = lo; x < hi; ++x) {
      const uint32_t u = uint32_t(x);
      Clus& c = res->pool[u];
      c.seq = R.uniq[u].seq;
      c.qual = R.uniq[u].qual;
      c.cnt = R.uniq[u].cnt;
      c.sumLenCnt = double(R.uniq[u].seq.size()) * R.uniq[u].cnt;
      c.firstId = u;
      c.storeIdx = u;
      c.members.init(u);
      c.avgErr = averageErrorRate(c.qual);
    }
  });
  const double tPrep = nowSec() - tStart;
  prepRange.reset();
  if (g_verboseTiming) fprintf(stderr, "[sdq]   prep: dedupe %.3f median %.3f upload %.3f pool %.3f\n", tDedupe, tMedian - tDedupe, tUpload - tMedian, tPrep - tUpload);
  R.dropRemovedAndSort(clusters);     // clusterDown.cpp:389
  int rc = R.buildIndex(clusters);    // clusterDown.cpp:412-417
  if (rc) return bail(rc);
  ClusVec singletons;
  if (!opts->startWithSingles) {  // clusterDown.cpp:467-484
    ClusVec keep;
    for (Clus& c : clusters) (c.cnt <= 1.01 ? singletons : keep).push_back(&c);
    clusters.ptrs().swap(keep.ptrs());
  }
  if ((rc = R.runFullClustering(clusters, initPars, nInit))) return bail(rc);  // :488-490
  if (!opts->startWithSingles && !opts->leaveOutSinglets) {                  // :492-499
    for (Clus& c : singletons) clusters.push_back(&c);
    if ((rc = R.runFullClustering(clusters, pars, nPars))) return bail(rc);
  }
  if (!opts->dontRecalLowFreq) {  // :502-519
    if ((rc = R.buildIndex(clusters))) return bail(rc);
    if ((rc = R.runFullClustering(clusters, pars, nPars))) return bail(rc);
  }
  const double tTail0 = nowSec();
  R.dropRemovedAndSort(clusters);  // :560
  parallelFor(clusters.size(), [&](size_t lo, size_t hi) {  // clusters are independent here; their reads are disjoint
    for (size_t c = lo; c < hi; ++c) {
      Clus& cl = clusters[c];
      cl.row.clear();
      cl.cons.reset();
      for (uint32_t u : cl.members)
        for (uint32_t x = R.uniq[u].inputBegin; x < R.uniq[u].inputEnd; ++x) res->membership[R.inputList[x]] = uint32_t(c);
      // the result outlives the caller's read buffers and this run's quality arena: every final cluster owns its bytes
      if (cl.seq.data() != reinterpret_cast<const uint8_t*>(cl.ownSeq.data())) {
        cl.ownSeq.assign(reinterpret_cast<const char*>(cl.seq.data()), cl.seq.size());
        cl.seq = View{reinterpret_cast<const uint8_t*>(cl.ownSeq.data()), uint32_t(cl.ownSeq.size())};
      }
      if (cl.qual.data() != cl.ownQual.data()) {
        cl.ownQual.assign(cl.qual.begin(), cl.qual.end());
        cl.qual = View{cl.ownQual.data(), uint32_t(cl.ownQual.size())};
      }
    }
  }, 1024);
  if (opts->markChimeras && markChimeras(ctx, *opts, clusters, R.st)) return bail(R.fail(SDGPU_E_CUDA));  // :566-577
  uint64_t launches1 = 0, cells1 = 0;
  float msResident = 0;
  if (sdgpu_timer_stop(ctx, &msResident)) return bail(R.fail(SDGPU_E_CUDA));
  if (sdgpu_counters(ctx, &launches1, &cells1)) return bail(R.fail(SDGPU_E_CUDA));
  res->clusters = clusters;
  R.st.msResident = msResident;
  res->stats = R.st;
  res->stats.inputReads = n;
  res->stats.uniqueSeqs = nU;
  res->stats.finalClusters = res->clusters.size();
  res->stats.cellUpdates = cells1 - cells0;
  res->stats.secondsTotal = nowSec() - tStart;
  if (getenv("SDQ_TIMING")) {
    fprintf(stderr, "[sdq] total %.3f prep %.3f gpu %.3f (submit %.3f, %llu batches) spec %.3f insert %.3f pass %.3f cons(incl gpu) %.3f sort %.3f index %.3f lines %zu\n",
            res->stats.secondsTotal, tPrep, R.st.secondsGpu, R.tSubmit, (unsigned long long)R.st.gpuBatches, R.tSpec, R.tInsert, R.tPass, R.tCons, R.tSort, R.tIndex, R.lines.size());
    fprintf(stderr, "[sdq] tail (final sort, membership, chimeras) %.3f\n", nowSec() - tTail0);
    fprintf(stderr, "[sdq] consensus: member alignments %.3f counting %.3f rebuild %.3f new versions %.3f\n", R.tConsAlign, R.tConsCount, R.tConsBuild,
            R.tConsUpload);
    fprintf(stderr, "[sdq] consensus updates kept on the device version: %llu, store entries: %u (unique reads %u), scans skipped %llu, streamed grids %llu (%llu callbacks, %.3f s inside, %llu reads passed there)\n",
            (unsigned long long)R.keptVersions, R.storeSize, nU, (unsigned long long)R.skippedScans, (unsigned long long)R.streamedGrids,
            (unsigned long long)R.nStreamCb, R.tStreamCb, (unsigned long long)R.streamedReads);
  }
  *out = res;
  return SDGPU_OK;
}

extern "C" {

int sdq_run(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit, const sdq_iter_par* pars,
            uint32_t nPars, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n,
            sdq_result** out) {
  if (!ctx || !opts || !out || (n && (!bases || !quals || !offsets)) || (nInit && !initPars) || (nPars && !pars)) {
    g_err = "sdq_run: null argument";
    return SDGPU_E_ARG;
  }
  ReadSrc in;
  in.bases = bases, in.quals = quals, in.offsets = offsets;
  return runCollapse(ctx, opts, initPars, nInit, pars, nPars, in, n, out);
}

// sampColl.doPopulationClustering(sampColl.createPopInput(), alignerObj, collapserObj, popIteratorMap)
// (processClusters.cpp:226): the final clusters of every sample, each with its own count and without any
// deduplication, go through collapser::runClustering once -- no k-mer checks (collapserObj.opts_.kmerOpts_.checkKmers_
// = false, :127), no singlet round, no k-mer re-count, no chimera marking.  membership of the result: for every input
// cluster, sample-major, the population cluster it ended in.
int sdq_population_cluster(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* popPars, uint32_t nPars,
                           const sdq_result* const* samples, uint32_t nSamples, sdq_result** out) {
  if (!ctx || !opts || !out || (nPars && !popPars) || (nSamples && !samples)) {
    g_err = "sdq_population_cluster: null argument";
    return SDGPU_E_ARG;
  }
  std::vector<const uint8_t*> sp, qp;
  std::vector<uint32_t> ln;
  std::vector<double> cn;
  for (uint32_t s = 0; s < nSamples; ++s) {
    if (!samples[s]) {
      g_err = "sdq_population_cluster: null sample result";
      return SDGPU_E_ARG;
    }
    for (const Clus& c : samples[s]->clusters) {
      sp.push_back(c.seq.data());
      qp.push_back(c.qual.data());
      ln.push_back(uint32_t(c.seq.size()));
      cn.push_back(c.cnt);
    }
  }
  sdq_opts o = *opts;
  o.checkKmers = 0;
  o.startWithSingles = 1;
  o.dontRecalLowFreq = 1;
  o.markChimeras = 0;
  ReadSrc in;
  in.seqPtrs = sp.data(), in.qualPtrs = qp.data(), in.lens = ln.data(), in.cnts = cn.data();
  return runCollapse(ctx, &o, popPars, nPars, popPars, nPars, in, uint32_t(sp.size()), out);
}

// the same with the input clusters packed by the caller (bases / quals / offsets like sdgpu_upload_seqs, cnts per cluster)
int sdq_population_cluster_packed(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* popPars, uint32_t nPars,
                                  const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, const double* cnts,
                                  uint32_t nClusters, sdq_result** out) {
  if (!ctx || !opts || !out || (nPars && !popPars) || (nClusters && (!bases || !quals || !offsets))) {
    g_err = "sdq_population_cluster_packed: null argument";
    return SDGPU_E_ARG;
  }
  std::vector<const uint8_t*> sp(nClusters), qp(nClusters);
  std::vector<uint32_t> ln(nClusters);
  for (uint32_t c = 0; c < nClusters; ++c) {
    sp[c] = bases + offsets[c];
    qp[c] = quals + offsets[c];
    ln[c] = uint32_t(offsets[c + 1] - offsets[c]);
  }
  sdq_opts o = *opts;
  o.checkKmers = 0;
  o.startWithSingles = 1;
  o.dontRecalLowFreq = 1;
  o.markChimeras = 0;
  ReadSrc in;
  in.seqPtrs = sp.data(), in.qualPtrs = qp.data(), in.lens = ln.data(), in.cnts = cnts;
  return runCollapse(ctx, &o, popPars, nPars, popPars, nPars, in, nClusters, out);
}

void sdq_free(sdq_result* r) { delete r; }

int sdq_get_stats(const sdq_result* r, sdq_stats* out) {
  if (!r || !out) return SDGPU_E_ARG;
  *out = r->stats;
  return SDGPU_OK;
}

uint32_t sdq_num_clusters(const sdq_result* r) { return r ? uint32_t(r->clusters.size()) : 0; }
uint32_t sdq_cluster_len(const sdq_result* r, uint32_t c) { return (r && c < r->clusters.size()) ? uint32_t(r->clusters[c].seq.size()) : 0u; }
double sdq_cluster_cnt(const sdq_result* r, uint32_t c) { return (r && c < r->clusters.size()) ? r->clusters[c].cnt : 0.0; }

uint64_t sdq_clusters_total_len(const sdq_result* r) {
  uint64_t n = 0;
  if (r)
    for (const Clus& c : r->clusters) n += c.seq.size();
  return n;
}

int sdq_clusters_packed(const sdq_result* r, uint8_t* bases, uint8_t* quals, uint64_t* offsets, double* cnts) {
  if (!r) return SDGPU_E_ARG;
  uint64_t off = 0;
  const size_t n = r->clusters.size();
  for (size_t c = 0; c < n; ++c) {
    const Clus& cl = r->clusters[c];
    if (offsets) offsets[c] = off;
    if (bases) std::memcpy(bases + off, cl.seq.data(), cl.seq.size());
    if (quals) std::memcpy(quals + off, cl.qual.data(), cl.qual.size());
    if (cnts) cnts[c] = cl.cnt;
    off += cl.seq.size();
  }
  if (offsets) offsets[n] = off;
  return SDGPU_OK;
}

int sdq_cluster_seq(const sdq_result* r, uint32_t c, uint8_t* seq, uint8_t* qual) {
  if (!r || c >= r->clusters.size()) return SDGPU_E_ARG;
  const Clus& cl = r->clusters[c];
  if (seq) std::memcpy(seq, cl.seq.data(), cl.seq.size());
  if (qual) std::memcpy(qual, cl.qual.data(), cl.qual.size());
  return SDGPU_OK;
}

int sdq_cluster_chimeric(const sdq_result* r, uint32_t c, uint32_t* parents, uint32_t* positions) {
  if (!r || c >= r->clusters.size()) return 0;
  const Clus& cl = r->clusters[c];
  if (cl.chimeric && parents) parents[0] = cl.chiParent[0], parents[1] = cl.chiParent[1];
  if (cl.chimeric && positions) positions[0] = cl.chiPos[0], positions[1] = cl.chiPos[1];
  return cl.chimeric ? 1 : 0;
}

int sdq_membership(const sdq_result* r, uint32_t* clusterOfRead) {
  if (!r || !clusterOfRead) return SDGPU_E_ARG;
  std::memcpy(clusterOfRead, r->membership.data(), r->membership.size() * sizeof(uint32_t));
  return SDGPU_OK;
}

// Many independent samples on one GPU: `nWorkers` host threads, each with its own ctx (stream,
// scratch, sequence store) on `device`, pull samples from a shared counter.  While one worker is
// in a host-only phase (dedupe, the sequential pass, consensus, sorting) another one's alignment
// batches keep the GPU busy.  Same model as the reference's runMultipleCommands over
// qlusterCmds.txt (SeekDeepUtilsRunner.cpp:294-356): samples never interact.
int sdq_run_many(int device, const sdgpu_params* params, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit,
                 const sdq_iter_par* pars, uint32_t nPars, uint32_t nSamples, const uint8_t* const* bases,
                 const uint8_t* const* quals, const uint64_t* const* offsets, const uint32_t* nReads, uint32_t nWorkers,
                 sdq_result** out) {
  if (!params || !opts || !out || (nSamples && (!bases || !quals || !offsets || !nReads))) {
    g_err = "sdq_run_many: null argument";
    return SDGPU_E_ARG;
  }
  if (nWorkers == 0) nWorkers = 1;
  if (nWorkers > nSamples) nWorkers = nSamples ? nSamples : 1;
  for (uint32_t s = 0; s < nSamples; ++s) out[s] = nullptr;
  std::atomic<uint32_t> next{0};
  std::atomic<int> firstRc{0};
  std::mutex errMutex;
  std::string firstErr;
  auto worker = [&]() {
    sdgpu_ctx* ctx = nullptr;
    int rc = sdgpu_create(device, params, &ctx);
    if (rc) {
      std::lock_guard<std::mutex> g(errMutex);
      if (!firstRc.exchange(rc)) firstErr = sdgpu_last_error(nullptr);
      return;
    }
    for (;;) {
      const uint32_t s = next.fetch_add(1);
      if (s >= nSamples || firstRc.load()) break;
      rc = sdq_run(ctx, opts, initPars, nInit, pars, nPars, bases[s], quals[s], offsets[s], nReads[s], &out[s]);
      if (rc) {
        std::lock_guard<std::mutex> g(errMutex);
        if (!firstRc.exchange(rc)) firstErr = g_err;  // g_err is thread-local: copy it out
        break;
      }
    }
    sdgpu_destroy(ctx);
  };
  std::vector<std::thread> threads;
  for (uint32_t w = 0; w < nWorkers; ++w) threads.emplace_back(worker);
  for (auto& t : threads) t.join();
  if (firstRc.load()) {
    for (uint32_t s = 0; s < nSamples; ++s) {
      delete out[s];
      out[s] = nullptr;
    }
    g_err = firstErr;
  }
  return firstRc.load();
}

}  // extern "C"
