// SHA-256 block function on the x86 SHA extensions (runtime-detected), used by the host-side
// Fiat-Shamir transcript (sha2.h).  The transcript is the serial tail of every phase (4.3 MB at
// n = 4096 MODP, 69 MB at n = 65536), so the portable C rounds in sha2.h are replaced when the
// CPU offers sha256rnds2/sha256msg1/sha256msg2.  Compiled by g++ with -msha -msse4.1 for this
// file only; everything else stays baseline x86-64.
#include <cpuid.h>
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

namespace sha2 {

bool cpu_has_sha_ni() {
  static const bool has = [] {
    unsigned a, b, c, d;
    if (!__get_cpuid_count(7, 0, &a, &b, &c, &d)) return false;
    bool sha = (b >> 29) & 1;
    if (!__get_cpuid(1, &a, &b, &c, &d)) return false;
    bool sse41 = (c >> 19) & 1, ssse3 = (c >> 9) & 1;
    return sha && sse41 && ssse3;
  }();
  return has;
}

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

// state: the eight working variables a..h; data: nblocks * 64 bytes
void sha256_ni_blocks(uint32_t state[8], const uint8_t* data, size_t nblocks) {
  const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
  __m128i tmp = _mm_loadu_si128(reinterpret_cast<const __m128i*>(&state[0]));     // DCBA
  __m128i st1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(&state[4]));     // HGFE
  tmp = _mm_shuffle_epi32(tmp, 0xB1);                                             // CDAB
  st1 = _mm_shuffle_epi32(st1, 0x1B);                                             // EFGH
  __m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                                     // ABEF
  st1 = _mm_blend_epi16(st1, tmp, 0xF0);                                          // CDGH
  while (nblocks--) {
    const __m128i save0 = st0, save1 = st1;
    __m128i m[4];
#pragma GCC unroll 16
    for (int i = 0; i < 16; ++i) {
      if (i < 4) {
        m[i] = _mm_shuffle_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(data + 16 * i)), bswap);
      } else {
        __m128i t = _mm_sha256msg1_epu32(m[(i - 4) & 3], m[(i - 3) & 3]);
        t = _mm_add_epi32(t, _mm_alignr_epi8(m[(i - 1) & 3], m[(i - 2) & 3], 4));
        m[i & 3] = _mm_sha256msg2_epu32(t, m[(i - 1) & 3]);
      }
      __m128i msg = _mm_add_epi32(m[i & 3], _mm_loadu_si128(reinterpret_cast<const __m128i*>(&K[4 * i])));
      st1 = _mm_sha256rnds2_epu32(st1, st0, msg);
      msg = _mm_shuffle_epi32(msg, 0x0E);
      st0 = _mm_sha256rnds2_epu32(st0, st1, msg);
    }
    st0 = _mm_add_epi32(st0, save0);
    st1 = _mm_add_epi32(st1, save1);
    data += 64;
  }
  tmp = _mm_shuffle_epi32(st0, 0x1B);        // FEBA
  st1 = _mm_shuffle_epi32(st1, 0xB1);        // DCHG
  st0 = _mm_blend_epi16(tmp, st1, 0xF0);     // DCBA
  st1 = _mm_alignr_epi8(st1, tmp, 8);        // HGFE
  _mm_storeu_si128(reinterpret_cast<__m128i*>(&state[0]), st0);
  _mm_storeu_si128(reinterpret_cast<__m128i*>(&state[4]), st1);
}

}  // namespace sha2
