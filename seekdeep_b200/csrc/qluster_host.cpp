// qluster_host.cpp — host side of the collapse loop (C ABI in include/sdgpu_qluster.h).
//
// Mirrors the call sequence of SeekDeep's qluster body (clusterDown.cpp:230-563) and njhseq's
// collapser::runFullClustering / collapseWithParameters / findMatch / cluster::compare /
// baseCluster::calculateConsensus as the reference drives them (clusterDown.cpp:488-518).
// The sequential greedy logic is unchanged; what differs is where comparisons come from:
//   * before an iteration's sequential pass, every (cluster, read) pair the pass can ask for
//     (each live read against the first `stopCheck` eligible clusters) that is not cached yet
//     goes to the GPU in large batches (sdgpu_align_profile: DP + traceback + errorProfile);
//   * the pass then reads comparisons from a pair cache keyed by the two sequence-store
//     indices.  A store index names one immutable (sequence, quality) version, so a cached
//     comparison is a pure function of its key (+ the k-mer index, on whose rebuild the cache is
//     dropped) — the same transparency njhseq's previousErrorChecks_ / alnHolder_ caches have;
//   * consensus rebuilding aligns all members of all changed clusters in one batch with gap
//     lists returned, builds the column counters on the host, and appends changed consensus
//     versions to the device sequence store.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/sdgpu_qluster.h"

namespace {

thread_local std::string g_err;

double nowSec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// What the sequential pass needs from a comparison (passErrorProfile / identity thresholds).
struct Verdict {
  double one, two, large, id, idHq;
  uint32_t hq, lq, lk;
  uint32_t valid;
};

// Open-addressing map (storeIdxA << 32 | storeIdxB) -> Verdict.
class PairCache {
 public:
  PairCache() { rehash(1u << 16); }
  const Verdict* find(uint64_t key) const {
    size_t h = hash(key) & mask_;
    while (keys_[h] != EMPTY) {
      if (keys_[h] == key) return &vals_[h];
      h = (h + 1) & mask_;
    }
    return nullptr;
  }
  void insert(uint64_t key, const Verdict& v) {
    if ((size_ + 1) * 10 > (mask_ + 1) * 6) rehash((mask_ + 1) * 2);
    size_t h = hash(key) & mask_;
    while (keys_[h] != EMPTY) {
      if (keys_[h] == key) {
        vals_[h] = v;
        return;
      }
      h = (h + 1) & mask_;
    }
    keys_[h] = key;
    vals_[h] = v;
    ++size_;
  }
  void clear() {
    std::fill(keys_.begin(), keys_.end(), EMPTY);
    size_ = 0;
  }
  // keep only entries whose two store indices are still alive
  void prune(const std::vector<uint8_t>& live) {
    std::vector<uint64_t> k;
    std::vector<Verdict> v;
    k.swap(keys_);
    v.swap(vals_);
    size_t cap = 1u << 16;
    size_t keep = 0;
    for (size_t i = 0; i < k.size(); ++i)
      if (k[i] != EMPTY && live[k[i] >> 32] && live[k[i] & 0xffffffffu]) ++keep;
    while (cap * 6 < (keep + 1) * 10 * 2) cap *= 2;
    keys_.assign(cap, EMPTY);
    vals_.resize(cap);
    mask_ = cap - 1;
    size_ = 0;
    for (size_t i = 0; i < k.size(); ++i)
      if (k[i] != EMPTY && live[k[i] >> 32] && live[k[i] & 0xffffffffu]) insert(k[i], v[i]);
  }
  size_t size() const { return size_; }

 private:
  static constexpr uint64_t EMPTY = ~0ull;
  static uint64_t hash(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
  }
  void rehash(size_t cap) {
    std::vector<uint64_t> k(cap, EMPTY);
    std::vector<Verdict> v(cap);
    k.swap(keys_);
    v.swap(vals_);
    mask_ = cap - 1;
    size_ = 0;
    for (size_t i = 0; i < k.size(); ++i)
      if (k[i] != EMPTY) insert(k[i], v[i]);
  }
  std::vector<uint64_t> keys_;
  std::vector<Verdict> vals_;
  size_t mask_ = 0, size_ = 0;
};

struct Uniq {  // one exact-unique input sequence; its store index is its position in `uniq`
  std::string seq;
  std::vector<uint8_t> qual;
  double cnt = 0;
  std::vector<uint32_t> inputIdx;
};

struct Clus {
  std::string seq;
  std::vector<uint8_t> qual;
  double cnt = 0;
  uint32_t firstId = 0;  // cluster::firstReadName_ stand-in (tie-break of the sort)
  uint32_t storeIdx = 0; // device sequence-store entry holding (seq, qual)
  std::vector<uint32_t> members;
  bool remove = false, needCons = false;
  double avgErr = 0;
};

double averageErrorRate(const std::vector<uint8_t>& q) {
  if (q.empty()) return 0;
  double s = 0;
  for (uint8_t v : q) s += std::pow(10.0, -(double(v) / 10.0));
  return s / double(q.size());
}

bool clusLess(const Clus& a, const Clus& b) {  // readVecSorter "totalCount"
  if (a.cnt != b.cnt) return a.cnt > b.cnt;
  if (a.avgErr != b.avgErr) return a.avgErr < b.avgErr;
  if (a.seq != b.seq) return a.seq < b.seq;
  return a.firstId < b.firstId;
}

bool passErrorCheck(const sdq_iter_par& it, const Verdict& g) {  // comparison::passErrorProfile
  if (it.onPerId) return (it.onHqPerId ? g.idHq : g.id) >= it.perId;
  return it.oneBaseIndel >= g.one && it.twoBaseIndel >= g.two && it.largeBaseIndel >= g.large &&
         it.hqMismatches >= double(g.hq) && it.lqMismatches >= double(g.lq) && it.lowKmerMismatches >= double(g.lk);
}

struct CharCounter {
  uint32_t cnt[128];
  uint32_t qual[128];
  CharCounter() {
    std::memset(cnt, 0, sizeof(cnt));
    std::memset(qual, 0, sizeof(qual));
  }
  void add(char b, uint32_t q, uint32_t w) {
    cnt[uint8_t(b) & 127] += w;
    qual[uint8_t(b) & 127] += q * w;
  }
  uint32_t total() const {
    uint32_t t = 0;
    for (uint32_t v : cnt) t += v;
    return t;
  }
  void best(char& b, uint32_t& q) const {
    uint32_t bc = 0;
    b = ' ';
    q = 0;
    for (int ch = 0; ch < 128; ++ch)
      if (cnt[ch] > bc) {
        bc = cnt[ch];
        b = char(ch);
        q = qual[ch];
      }
  }
};

class Runner {
 public:
  Runner(sdgpu_ctx* ctx, const sdq_opts& o) : ctx_(ctx), O(o) {}
  sdgpu_ctx* ctx_;
  const sdq_opts& O;
  std::vector<Uniq> uniq;
  PairCache cache;
  sdq_stats st{};
  // host mirror of the store contents appended after the unique reads (consensus versions)
  uint32_t storeSize = 0;
  static constexpr uint32_t BATCH = 1u << 21;
  std::vector<sdgpu_comparison> cmpBuf;
  std::vector<sdgpu_gap> gapBuf;

  int fail(int rc) {
    g_err = sdgpu_last_error(ctx_);
    return rc;
  }

  uint32_t flags() const { return (O.checkKmers ? SDGPU_CHECK_KMER : 0u) | SDGPU_USING_QUALITY; }

  // comparisons for (a[i], b[i]) store-index pairs -> cache
  int computeToCache(const std::vector<uint32_t>& a, const std::vector<uint32_t>& b) {
    const size_t n = a.size();
    for (size_t off = 0; off < n; off += BATCH) {
      const uint32_t m = uint32_t(std::min<size_t>(BATCH, n - off));
      if (cmpBuf.size() < m) cmpBuf.resize(m);
      const double t0 = nowSec();
      int rc = sdgpu_align_profile(ctx_, a.data() + off, b.data() + off, m, flags(), cmpBuf.data(), nullptr, 0);
      st.secondsGpu += nowSec() - t0;
      if (rc) return fail(rc);
      ++st.gpuBatches;
      st.pairsComputed += m;
      for (uint32_t i = 0; i < m; ++i) {
        const sdgpu_comparison& c = cmpBuf[i];
        Verdict v{c.oneBaseIndel, c.twoBaseIndel, c.largeBaseIndel, c.eventBasedIdentity, c.eventBasedIdentityHq,
                  c.hqMismatches, c.lqMismatches, c.lowKmerMismatches, 1u};
        cache.insert((uint64_t(a[off + i]) << 32) | b[off + i], v);
      }
    }
    return 0;
  }

  static uint64_t key(const Clus& clus, const Clus& read) { return (uint64_t(clus.storeIdx) << 32) | read.storeIdx; }

  // candidate scan shared by speculation and the sequential pass: eligible clusters in sorted
  // order, skipping `rp` itself (njhseq findMatch: remove / smallCutoff / same-name skips)
  int collapseWithParameters(std::vector<Clus>& cs, const sdq_iter_par& it) {
    size_t alive = 0;
    for (const Clus& c : cs) alive += !c.remove;
    if (alive < 2) return 0;
    std::vector<uint32_t> elig;
    for (uint32_t cp = 0; cp < cs.size(); ++cp)
      if (!cs[cp].remove && cs[cp].cnt > double(it.smallCutoff)) elig.push_back(cp);
    // ---- speculative batch: everything the pass below can ask for and the cache lacks
    {
      std::vector<uint32_t> a, b;
      for (uint32_t rp = 0; rp < cs.size(); ++rp) {
        if (cs[rp].remove) continue;
        uint32_t count = 0;
        for (uint32_t e = 0; e < elig.size() && count < it.stopCheck; ++e) {
          const uint32_t cp = elig[e];
          if (cp == rp) continue;
          ++count;
          if (!cache.find(key(cs[cp], cs[rp]))) {
            a.push_back(cs[cp].storeIdx);
            b.push_back(cs[rp].storeIdx);
          }
        }
      }
      int rc = computeToCache(a, b);
      if (rc) return rc;
    }
    // ---- the sequential pass (smallest cluster first; first passing candidate absorbs it)
    size_t amountAdded = 0;
    std::vector<uint32_t> da, db;
    for (size_t rp = cs.size(); rp-- > 0;) {
      Clus& read = cs[rp];
      if (read.remove) continue;
      uint32_t count = 0;
      for (uint32_t e = 0; e < elig.size(); ++e) {
        const uint32_t cp = elig[e];
        Clus& clus = cs[cp];
        if (clus.remove) continue;
        if (cp == rp) continue;
        ++count;
        if (count > it.stopCheck) break;
        const Verdict* v = cache.find(key(clus, read));
        if (!v) {  // a candidate that slid into the window because earlier ones were absorbed
          da.assign(1, clus.storeIdx);
          db.assign(1, read.storeIdx);
          int rc = computeToCache(da, db);
          if (rc) return rc;
          ++st.onDemandPairs;
          v = cache.find(key(clus, read));
        }
        ++st.pairsConsumed;
        if (passErrorCheck(it, *v)) {
          clus.cnt += read.cnt;
          clus.members.insert(clus.members.end(), read.members.begin(), read.members.end());
          clus.needCons = true;
          read.remove = true;
          ++amountAdded;
          break;
        }
      }
    }
    if (amountAdded > 0) return calculateConsensusBatch(cs);
    return 0;
  }

  // baseCluster::calculateConsensus for every cluster that absorbed reads, one GPU batch.
  int calculateConsensusBatch(std::vector<Clus>& cs) {
    struct Job {
      uint32_t clus;
      uint32_t baseStore;  // store index of the sequence members are aligned to
      const std::string* baseSeq;
      const std::vector<uint8_t>* baseQual;
      size_t firstPair;
    };
    std::vector<Job> jobs;
    std::vector<uint32_t> a, b;
    for (uint32_t ci = 0; ci < cs.size(); ++ci) {
      Clus& cl = cs[ci];
      if (cl.remove || !cl.needCons) continue;
      cl.needCons = false;
      if (cl.members.size() <= 1) continue;
      double totalLen = 0, totalCnt = 0;
      for (uint32_t u : cl.members) {
        totalLen += uniq[u].seq.size() * uniq[u].cnt;
        totalCnt += uniq[u].cnt;
      }
      const double avgLen = totalLen / totalCnt;
      Job j{ci, cl.storeIdx, &cl.seq, &cl.qual, a.size()};
      if (std::fabs(double(cl.seq.size()) - avgLen) > 0.1 * avgLen) {  // restart from the longest member
        size_t biggest = 0;
        for (uint32_t u : cl.members)
          if (uniq[u].seq.size() > biggest) {
            biggest = uniq[u].seq.size();
            j.baseStore = u;
            j.baseSeq = &uniq[u].seq;
            j.baseQual = &uniq[u].qual;
          }
      }
      for (uint32_t u : cl.members) {
        a.push_back(j.baseStore);
        b.push_back(u);
      }
      jobs.push_back(j);
    }
    if (jobs.empty()) return 0;
    // alignments with gap lists; pairs whose list overflowed are re-run with a larger capacity
    const size_t n = a.size();
    uint32_t gapCap = 8;
    std::vector<std::vector<sdgpu_gap>> gaps(n);
    std::vector<uint32_t> todo(n);
    for (size_t i = 0; i < n; ++i) todo[i] = uint32_t(i);
    while (!todo.empty()) {
      std::vector<uint32_t> next;
      for (size_t off = 0; off < todo.size(); off += BATCH / 4) {
        const uint32_t m = uint32_t(std::min<size_t>(BATCH / 4, todo.size() - off));
        std::vector<uint32_t> pa(m), pb(m);
        for (uint32_t i = 0; i < m; ++i) {
          pa[i] = a[todo[off + i]];
          pb[i] = b[todo[off + i]];
        }
        if (cmpBuf.size() < m) cmpBuf.resize(m);
        if (gapBuf.size() < size_t(m) * gapCap) gapBuf.resize(size_t(m) * gapCap);
        const double t0 = nowSec();
        int rc = sdgpu_align_profile(ctx_, pa.data(), pb.data(), m, 0u, cmpBuf.data(), gapBuf.data(), gapCap);
        st.secondsGpu += nowSec() - t0;
        if (rc) return fail(rc);
        ++st.gpuBatches;
        st.consensusAligns += m;
        for (uint32_t i = 0; i < m; ++i) {
          if (cmpBuf[i].status & SDGPU_ST_GAPLIST_TRUNCATED) {
            next.push_back(todo[off + i]);
          } else {
            gaps[todo[off + i]].assign(gapBuf.begin() + size_t(i) * gapCap,
                                       gapBuf.begin() + size_t(i) * gapCap + cmpBuf[i].nGapRuns);
          }
        }
      }
      todo.swap(next);
      gapCap *= 8;
    }
    // column counters -> new consensus (ConsensusHelper::increaseCounters / buildConsensus)
    std::vector<uint32_t> changed;
    std::string ga, gb;
    std::vector<uint8_t> gq;
    for (const Job& j : jobs) {
      Clus& cl = cs[j.clus];
      const std::string base = *j.baseSeq;  // copies: cl.seq may be replaced below
      std::vector<CharCounter> counters(base.size());
      std::map<uint32_t, std::map<uint32_t, CharCounter>> insertions;
      std::map<uint32_t, CharCounter> beginningGap;
      for (size_t m = 0; m < cl.members.size(); ++m) {
        const Uniq& r = uniq[cl.members[m]];
        const uint32_t w = uint32_t(std::llround(r.cnt));
        ga = base;
        gb = r.seq;
        gq = r.qual;
        for (const sdgpu_gap& g : gaps[j.firstPair + m]) {  // descending positions
          if (g.gapInA) {
            ga.insert(g.pos, g.size, '-');
          } else {
            gb.insert(g.pos, g.size, '-');
            gq.insert(gq.begin() + g.pos, g.size, uint8_t(0));
          }
        }
        uint32_t offSet = 0, currentOffset = 1, start = 0;
        const uint32_t len = uint32_t(ga.size());
        if (len && ga[0] == '-') {
          while (start < len && ga[start] == '-') ++start;
          for (uint32_t i = 0; i < start; ++i) beginningGap[i].add(gb[i], gq[i], w);
          offSet = start;
        }
        for (uint32_t i = start; i < len; ++i) {
          if (ga[i] == '-') {
            insertions[i - offSet][currentOffset].add(gb[i], gq[i], w);
            ++currentOffset;
            ++offSet;
            continue;
          }
          currentOffset = 1;
          counters[i - offSet].add(gb[i], gq[i], w);
        }
      }
      std::string ns;
      std::vector<uint8_t> nq;
      const double fortyPercent = 0.40 * cl.cnt;
      const uint32_t size = uint32_t(std::llround(cl.cnt));
      for (const auto& bg : beginningGap) {
        char bch;
        uint32_t q;
        bg.second.best(bch, q);
        const uint32_t tot = bg.second.total();
        if (bch == '-' || tot < fortyPercent) continue;
        ns.push_back(bch);
        nq.push_back(uint8_t(q / tot));
      }
      for (uint32_t pos = 0; pos < counters.size(); ++pos) {
        auto ins = insertions.find(pos);
        if (ins != insertions.end()) {
          for (const auto& ci : ins->second) {
            char bch;
            uint32_t q;
            ci.second.best(bch, q);
            const uint32_t tot = ci.second.total();
            const uint32_t bc = ci.second.cnt[uint8_t(bch) & 127];
            const uint32_t nonInserting = size > tot ? size - tot : 0;
            if (bch == ' ' || bc <= nonInserting) continue;
            ns.push_back(bch);
            nq.push_back(uint8_t(q / bc));
          }
        }
        char bch;
        uint32_t q;
        counters[pos].best(bch, q);
        const uint32_t tot = counters[pos].total();
        if (bch == '-' || bch == ' ' || tot < fortyPercent) continue;
        ns.push_back(bch);
        nq.push_back(uint8_t(q / tot));
      }
      if (ns != cl.seq || nq != cl.qual) {
        cl.seq.swap(ns);
        cl.qual.swap(nq);
        changed.push_back(j.clus);
      }
      cl.avgErr = averageErrorRate(cl.qual);
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
      for (size_t i = 0; i < changed.size(); ++i) cs[changed[i]].storeIdx = first + uint32_t(i);
      storeSize = first + uint32_t(changed.size());
    }
    return 0;
  }

  void dropRemovedAndSort(std::vector<Clus>& cs) {
    cs.erase(std::remove_if(cs.begin(), cs.end(), [](const Clus& c) { return c.remove; }), cs.end());
    std::sort(cs.begin(), cs.end(), clusLess);
  }

  int runFullClustering(std::vector<Clus>& cs, const sdq_iter_par* its, uint32_t nIts) {
    for (uint32_t i = 0; i < nIts; ++i) {
      dropRemovedAndSort(cs);
      if (cs.size() < 2) break;
      if (cache.size() > (1u << 22)) {  // forget pairs that can never be asked for again
        std::vector<uint8_t> live(storeSize, 0);
        for (const Clus& c : cs) live[c.storeIdx] = 1;
        cache.prune(live);
      }
      int rc = collapseWithParameters(cs, its[i]);
      if (rc) return rc;
    }
    dropRemovedAndSort(cs);
    return 0;
  }

  int buildIndex(const std::vector<Clus>& cs) {  // indexKmers over the current clusters
    std::vector<uint32_t> idx(cs.size());
    std::vector<double> cnts(cs.size());
    std::vector<uint32_t> order(cs.size());
    for (uint32_t i = 0; i < cs.size(); ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cs[x].storeIdx < cs[y].storeIdx; });
    for (uint32_t i = 0; i < cs.size(); ++i) {
      idx[i] = cs[order[i]].storeIdx;
      cnts[i] = cs[order[i]].cnt;
    }
    int rc = sdgpu_update_cnts(ctx_, idx.data(), cnts.data(), uint32_t(idx.size()));
    if (rc) return fail(rc);
    rc = sdgpu_build_kmer_index(ctx_, idx.data(), uint32_t(idx.size()), O.kLen, O.runCutOff, int(O.kmersByPosition),
                                int(O.expandKmerPos), O.expandKmerSize);
    if (rc) return fail(rc);
    cache.clear();  // comparisons depend on the index through lowKmerMismatches
    return 0;
  }
};

}  // namespace

struct sdq_result {
  std::vector<Clus> clusters;
  std::vector<uint32_t> membership;
  sdq_stats stats;
};

extern "C" {

const char* sdq_last_error(void) { return g_err.c_str(); }

int sdq_run(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit, const sdq_iter_par* pars,
            uint32_t nPars, const uint8_t* bases, const uint8_t* quals, const uint64_t* offsets, uint32_t n,
            sdq_result** out) {
  if (!ctx || !opts || !out || (n && (!bases || !quals || !offsets)) || (nInit && !initPars) || (nPars && !pars)) {
    g_err = "sdq_run: null argument";
    return SDGPU_E_ARG;
  }
  *out = nullptr;
  const double tStart = nowSec();
  Runner R(ctx, *opts);
  auto res = new sdq_result();
  auto bail = [&](int rc) {
    delete res;
    return rc;
  };
  res->membership.assign(n, UINT32_MAX);
  uint64_t launches0 = 0, cells0 = 0;
  if (sdgpu_counters(ctx, &launches0, &cells0)) return bail(R.fail(SDGPU_E_CUDA));
  // exact-sequence dedupe in first-seen order (clusterDown.cpp:272-288; hashing instead of the
  // reference's O(N*U) scan, same result)
  {
    std::unordered_map<std::string, uint32_t> seen;
    seen.reserve(n);
    for (uint32_t r = 0; r < n; ++r) {
      const uint32_t len = uint32_t(offsets[r + 1] - offsets[r]);
      if (len <= opts->smallReadSize) continue;
      std::string s(reinterpret_cast<const char*>(bases + offsets[r]), len);
      auto it = seen.find(s);
      if (it == seen.end()) {
        it = seen.emplace(s, uint32_t(R.uniq.size())).first;
        R.uniq.emplace_back();
        R.uniq.back().seq = std::move(s);
      }
      Uniq& u = R.uniq[it->second];
      u.cnt += 1.0;
      u.inputIdx.push_back(r);
    }
  }
  // per-base median quality over the members (clusterDown.cpp:293-359, qualRep "median")
  {
    std::vector<uint8_t> col;
    for (Uniq& u : R.uniq) {
      const size_t len = u.seq.size(), k = u.inputIdx.size();
      u.qual.resize(len);
      if (k == 1) {
        std::memcpy(u.qual.data(), quals + offsets[u.inputIdx[0]], len);
        continue;
      }
      col.resize(k);
      for (size_t i = 0; i < len; ++i) {
        for (size_t m = 0; m < k; ++m) col[m] = quals[offsets[u.inputIdx[m]] + i];
        std::sort(col.begin(), col.end());
        u.qual[i] = (k % 2) ? col[k / 2] : uint8_t((uint32_t(col[k / 2 - 1]) + col[k / 2]) / 2);
      }
    }
  }
  const uint32_t nU = uint32_t(R.uniq.size());
  res->stats.inputReads = n;
  res->stats.uniqueSeqs = nU;
  if (nU == 0) {
    res->stats.secondsTotal = nowSec() - tStart;
    *out = res;
    return SDGPU_OK;
  }
  // the unique reads become sequence-store entries 0..nU-1
  {
    std::vector<uint8_t> b, q;
    std::vector<uint64_t> off{0};
    std::vector<double> cn;
    for (const Uniq& u : R.uniq) {
      b.insert(b.end(), u.seq.begin(), u.seq.end());
      q.insert(q.end(), u.qual.begin(), u.qual.end());
      off.push_back(b.size());
      cn.push_back(u.cnt);
    }
    if (sdgpu_upload_seqs(ctx, b.data(), q.data(), off.data(), cn.data(), nU)) return bail(R.fail(SDGPU_E_CUDA));
    R.storeSize = nU;
  }
  std::vector<Clus> clusters(nU);
  for (uint32_t u = 0; u < nU; ++u) {
    Clus& c = clusters[u];
    c.seq = R.uniq[u].seq;
    c.qual = R.uniq[u].qual;
    c.cnt = R.uniq[u].cnt;
    c.firstId = u;
    c.storeIdx = u;
    c.members = {u};
    c.avgErr = averageErrorRate(c.qual);
  }
  std::sort(clusters.begin(), clusters.end(), clusLess);  // clusterDown.cpp:389
  int rc = R.buildIndex(clusters);                         // clusterDown.cpp:412-417
  if (rc) return bail(rc);
  std::vector<Clus> singletons;
  if (!opts->startWithSingles) {  // clusterDown.cpp:467-484
    std::vector<Clus> keep;
    for (Clus& c : clusters) (c.cnt <= 1.01 ? singletons : keep).push_back(std::move(c));
    clusters.swap(keep);
  }
  if ((rc = R.runFullClustering(clusters, initPars, nInit))) return bail(rc);  // :488-490
  if (!opts->startWithSingles && !opts->leaveOutSinglets) {                  // :492-499
    for (Clus& c : singletons) clusters.push_back(std::move(c));
    if ((rc = R.runFullClustering(clusters, pars, nPars))) return bail(rc);
  }
  if (!opts->dontRecalLowFreq) {  // :502-519
    if ((rc = R.buildIndex(clusters))) return bail(rc);
    if ((rc = R.runFullClustering(clusters, pars, nPars))) return bail(rc);
  }
  std::sort(clusters.begin(), clusters.end(), clusLess);  // :560
  for (uint32_t c = 0; c < clusters.size(); ++c)
    for (uint32_t u : clusters[c].members)
      for (uint32_t r : R.uniq[u].inputIdx) res->membership[r] = c;
  uint64_t launches1 = 0, cells1 = 0;
  if (sdgpu_counters(ctx, &launches1, &cells1)) return bail(R.fail(SDGPU_E_CUDA));
  res->clusters = std::move(clusters);
  res->stats = R.st;
  res->stats.inputReads = n;
  res->stats.uniqueSeqs = nU;
  res->stats.finalClusters = res->clusters.size();
  res->stats.cellUpdates = cells1 - cells0;
  res->stats.secondsTotal = nowSec() - tStart;
  *out = res;
  return SDGPU_OK;
}

void sdq_free(sdq_result* r) { delete r; }

int sdq_get_stats(const sdq_result* r, sdq_stats* out) {
  if (!r || !out) return SDGPU_E_ARG;
  *out = r->stats;
  return SDGPU_OK;
}

uint32_t sdq_num_clusters(const sdq_result* r) { return r ? uint32_t(r->clusters.size()) : 0; }
uint32_t sdq_cluster_len(const sdq_result* r, uint32_t c) { return uint32_t(r->clusters[c].seq.size()); }
double sdq_cluster_cnt(const sdq_result* r, uint32_t c) { return r->clusters[c].cnt; }

int sdq_cluster_seq(const sdq_result* r, uint32_t c, uint8_t* seq, uint8_t* qual) {
  if (!r || c >= r->clusters.size()) return SDGPU_E_ARG;
  const Clus& cl = r->clusters[c];
  if (seq) std::memcpy(seq, cl.seq.data(), cl.seq.size());
  if (qual) std::memcpy(qual, cl.qual.data(), cl.qual.size());
  return SDGPU_OK;
}

int sdq_membership(const sdq_result* r, uint32_t* clusterOfRead) {
  if (!r || !clusterOfRead) return SDGPU_E_ARG;
  std::memcpy(clusterOfRead, r->membership.data(), r->membership.size() * sizeof(uint32_t));
  return SDGPU_OK;
}

}  // extern "C"
