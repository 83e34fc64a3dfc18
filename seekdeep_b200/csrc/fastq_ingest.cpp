// fastq_ingest.cpp — FASTQ ingest in front of the collapse (the reference reads its input through njhseq's
// SeqInput::readNextRead, clusterDown.cpp:207-232, then trims / filters per read, :233-270).
//
// Reads a FASTQ file -- plain or gzip, decided by the magic bytes -- straight into the packed (bases, quals, offsets)
// form sdq_run takes: four-line records ('@name', bases, '+...', qualities), qualities converted from Phred+33 to
// numbers, lower-case bases upper-cased (readVec::handelLowerCaseBases with the default "upper"), '\r' dropped.
// Multi-line records and a bases/qualities length mismatch are refused loudly, not repaired.
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/sdgpu_qluster.h"

struct sdq_reads {
  std::vector<uint8_t> bases, quals;
  std::vector<uint64_t> offsets{0};
  std::vector<char> names;          // '\0'-separated read names (without '@')
  std::vector<uint64_t> nameOffsets{0};
};

namespace {
thread_local std::string g_fqErr;

class LineReader {  // gzread handles both gzip and plain files
 public:
  explicit LineReader(gzFile f) : f_(f), buf_(1 << 20) {}
  // next line without its terminator; false at end of file
  bool next(std::string& line) {
    line.clear();
    for (;;) {
      if (pos_ == end_) {
        const int n = gzread(f_, buf_.data(), unsigned(buf_.size()));
        if (n < 0) {
          bad_ = true;
          return false;
        }
        if (n == 0) {  // end of file: a last line without a newline still counts
          if (!line.empty() && line.back() == '\r') line.pop_back();
          return !line.empty();
        }
        pos_ = 0;
        end_ = size_t(n);
      }
      const char* p = buf_.data() + pos_;
      const char* nl = static_cast<const char*>(std::memchr(p, '\n', end_ - pos_));
      if (nl) {
        line.append(p, size_t(nl - p));
        pos_ += size_t(nl - p) + 1;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        return true;
      }
      line.append(p, end_ - pos_);
      pos_ = end_;
    }
  }
  bool bad() const { return bad_; }

 private:
  gzFile f_;
  std::vector<char> buf_;
  size_t pos_ = 0, end_ = 0;
  bool bad_ = false;
};
}  // namespace

extern "C" {

const char* sdq_fastq_last_error(void) { return g_fqErr.c_str(); }

int sdq_read_fastq(const char* path, sdq_reads** out) {
  if (!path || !out) return SDGPU_E_ARG;
  *out = nullptr;
  gzFile f = gzopen(path, "rb");
  if (!f) {
    g_fqErr = std::string("cannot open ") + path;
    return SDGPU_E_ARG;
  }
  gzbuffer(f, 1 << 20);
  auto* r = new sdq_reads();
  LineReader in(f);
  std::string name, seq, plus, qual;
  uint64_t rec = 0;
  auto fail = [&](const std::string& msg) {
    g_fqErr = std::string(path) + ": record " + std::to_string(rec + 1) + ": " + msg;
    delete r;
    gzclose(f);
    return SDGPU_E_UNSUPPORTED;
  };
  while (in.next(name)) {
    if (name.empty()) continue;  // blank lines between records
    if (name[0] != '@') return fail("header line does not start with '@'");
    if (!in.next(seq) || !in.next(plus) || !in.next(qual)) return fail("truncated record");
    if (plus.empty() || plus[0] != '+') return fail("third line does not start with '+' (multi-line FASTQ is not supported)");
    if (seq.size() != qual.size()) return fail("bases and qualities differ in length");
    for (char c : seq) {
      const unsigned char u = static_cast<unsigned char>(c);
      if (u >= 128) return fail("base outside 7-bit ASCII");
      r->bases.push_back(uint8_t((u >= 'a' && u <= 'z') ? u - 32 : u));
    }
    for (char c : qual) {
      const int q = static_cast<unsigned char>(c) - 33;
      if (q < 0 || q > 93) return fail("quality character outside Phred+33");
      r->quals.push_back(uint8_t(q));
    }
    r->offsets.push_back(r->bases.size());
    r->names.insert(r->names.end(), name.begin() + 1, name.end());
    r->names.push_back('\0');
    r->nameOffsets.push_back(r->names.size());
    ++rec;
  }
  const bool bad = in.bad();
  gzclose(f);
  if (bad) {
    g_fqErr = std::string(path) + ": read error (corrupt gzip stream?)";
    delete r;
    return SDGPU_E_UNSUPPORTED;
  }
  *out = r;
  return SDGPU_OK;
}

void sdq_reads_free(sdq_reads* r) { delete r; }
uint32_t sdq_reads_count(const sdq_reads* r) { return r ? uint32_t(r->offsets.size() - 1) : 0u; }
const uint8_t* sdq_reads_bases(const sdq_reads* r) { return r ? r->bases.data() : nullptr; }
const uint8_t* sdq_reads_quals(const sdq_reads* r) { return r ? r->quals.data() : nullptr; }
const uint64_t* sdq_reads_offsets(const sdq_reads* r) { return r ? r->offsets.data() : nullptr; }
const char* sdq_reads_name(const sdq_reads* r, uint32_t i) {
  return (r && i + 1 < r->nameOffsets.size()) ? r->names.data() + r->nameOffsets[i] : nullptr;
}

int sdq_run_fastq(sdgpu_ctx* ctx, const sdq_opts* opts, const sdq_iter_par* initPars, uint32_t nInit, const sdq_iter_par* pars,
                  uint32_t nPars, const char* fastqPath, sdq_result** out) {
  sdq_reads* reads = nullptr;
  int rc = sdq_read_fastq(fastqPath, &reads);
  if (rc) return rc;
  rc = sdq_run(ctx, opts, initPars, nInit, pars, nPars, reads->bases.data(), reads->quals.data(), reads->offsets.data(),
               uint32_t(reads->offsets.size() - 1), out);
  sdq_reads_free(reads);  // the result owns its bytes (sdq_run copies what it keeps)
  return rc;
}

}  // extern "C"
