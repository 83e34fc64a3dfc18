// host_translate.cpp — base characters -> store codes on the host, the inner loop of sdgpu_upload_seqs* (compiled by the
// host compiler alone: it uses x86 vector intrinsics behind a run-time check).
//
// The ctx alphabet is dynamic (any 7-bit character can get a code), but A, C, G, T, N always own codes 0..4 and are all a
// read normally holds.  Their low nibbles differ (1, 3, 7, 4, 14), so 32 bases at a time are classified and translated by
// two byte shuffles; a block holding anything else goes through the table byte by byte.
#include <cstdint>
#include <cstring>

#if defined(__x86_64__)
#include <immintrin.h>

__attribute__((target("avx2"))) static uint64_t translateAvx2(const uint8_t* src, uint8_t* dst, uint64_t len, int* maxCode) {
  // nibble -> the one character of {A, C, G, T, N} that has it, and that character's code
  const __m256i chars = _mm256_setr_epi8(0, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 'N', 0, 0, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0,
                                         0, 'N', 0);
  const __m256i codes = _mm256_setr_epi8(0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 4, 0, 0, 0, 0, 1, 3, 0, 0, 2, 0, 0, 0, 0, 0, 0, 4, 0);
  const __m256i low = _mm256_set1_epi8(0x0f);
  __m256i mx = _mm256_setzero_si256();
  uint64_t x = 0;
  for (; x + 32 <= len; x += 32) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + x));
    const __m256i nib = _mm256_and_si256(v, low);
    const __m256i want = _mm256_shuffle_epi8(chars, nib);
    if (_mm256_movemask_epi8(_mm256_cmpeq_epi8(want, v)) != -1) break;  // something else in this block: the caller's table takes over
    const __m256i c = _mm256_shuffle_epi8(codes, nib);
    mx = _mm256_max_epu8(mx, c);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(dst + x), c);
  }
  alignas(32) uint8_t m[32];
  _mm256_store_si256(reinterpret_cast<__m256i*>(m), mx);
  int best = -1;
  if (x > 0)
    for (int i = 0; i < 32; ++i) best = m[i] > best ? m[i] : best;
  if (best > *maxCode) *maxCode = best;
  return x;  // bases translated
}
#endif

// Translates len characters with codeOf (256 entries, -1 = no code yet; A, C, G, T, N = 0..4 by construction).  Bases
// without a code are written as 0xFF and reported; *maxCode is raised to the largest code met.
extern "C" void sd_translate_bases(const uint8_t* src, uint8_t* dst, uint64_t len, const int16_t* codeOf, int* maxCode, bool* unknown) {
  uint64_t x = 0;
#if defined(__x86_64__)
  static const bool haveAvx2 = __builtin_cpu_supports("avx2") && codeOf['A'] == 0 && codeOf['C'] == 1 && codeOf['G'] == 2 &&
                               codeOf['T'] == 3 && codeOf['N'] == 4;
#endif
  int mc = *maxCode;
  bool unk = false;
  while (x < len) {
#if defined(__x86_64__)
    if (haveAvx2 && len - x >= 32) {
      x += translateAvx2(src + x, dst + x, len - x, &mc);
      if (x >= len) break;
    }
#endif
    // up to 32 bases through the table (a block with another character in it, or the tail)
    const uint64_t end = len - x < 32 ? len : x + 32;
    for (; x < end; ++x) {
      const int code = codeOf[src[x]];
      unk |= code < 0;
      mc = code > mc ? code : mc;
      dst[x] = uint8_t(code);
    }
  }
  *maxCode = mc;
  *unknown = *unknown || unk;
}
