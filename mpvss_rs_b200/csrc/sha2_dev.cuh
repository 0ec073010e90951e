// Device-side SHA-256 / SHA-512 (FIPS 180-4) for the per-share Fiat-Shamir transcripts.
//
// extract_secret_share and verify_share hash one small transcript PER SHARE -- F(pk) F(Y) F(a1) F(a2) with
// F(e) = len_u64_be || bytes (dleq.rs:58-61, 87-99; participant.rs:330-347, 378-385) -- and turn the digest into
// the challenge with hash_to_scalar, which hashes once more (modp.rs:142-148 and secp256k1.rs:121-131: SHA-256,
// big-endian; ristretto255.rs:196-205: SHA-512, little-endian).  n shares are n independent hash chains, so they
// run one chain per thread on the device and a1 / a2 never travel to the host.  The whole-box transcript of
// distribute / verify_distribution is ONE sequential chain over all participants; that one stays on the host's
// SHA-NI unit (row_hash_body with a single row spanning the box exists as `box_hash_body` to measure exactly that
// trade: one GPU thread hashes at about 15 MB/s, the host at 1.9 GB/s).
//
// Input rows are the framed rows the frame kernels already write (modp::frame_body, ec::frame_body): four
// slots of `slot_stride` bytes, each `len_u64_be || len bytes`, left-aligned.
#pragma once
#include "simt.h"

namespace shadev {

MP_DEV uint32_t ror32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
MP_DEV uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }

// one SHA-256 block; w = the 16 big-endian message words (destroyed: it is the rolling schedule window)
MP_DEV void sha256_compress(uint32_t (&st)[8], uint32_t (&w)[16]) {
  const uint32_t K[64] = {
      0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
      0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
      0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
      0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
      0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
      0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
      0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
      0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 64; ++i) {
    if (i >= 16) {
      const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint32_t s0 = ror32(w15, 7) ^ ror32(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = ror32(w2, 17) ^ ror32(w2, 19) ^ (w2 >> 10);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = h + S1 + ch + K[i] + w[i & 15];
    const uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
    const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

MP_DEV void sha256_init(uint32_t (&st)[8]) {
  st[0] = 0x6a09e667u; st[1] = 0xbb67ae85u; st[2] = 0x3c6ef372u; st[3] = 0xa54ff53au;
  st[4] = 0x510e527fu; st[5] = 0x9b05688cu; st[6] = 0x1f83d9abu; st[7] = 0x5be0cd19u;
}

// Streaming SHA-256 of one thread.  Bytes are shifted into the current word, so a word needs no clearing
// between blocks (four shifts push the old content out).
struct Sha256 {
  uint32_t st[8];
  uint32_t w[16];
  uint32_t fill;   // bytes in the current block
  uint32_t total;  // message bytes so far (transcripts here are far below 2^29 bytes)
};
MP_DEV void init(Sha256& s) {
  sha256_init(s.st);
  s.fill = 0;
  s.total = 0;
}
MP_DEV void put(Sha256& s, uint32_t byte) {
  const uint32_t k = s.fill >> 2;
  s.w[k] = (s.w[k] << 8) | (byte & 0xffu);
  if (++s.fill == 64) {
    sha256_compress(s.st, s.w);
    s.fill = 0;
  }
}
MP_DEV void update(Sha256& s, const uint8_t* p, uint32_t n) {
  s.total += n;
#pragma unroll 1
  for (uint32_t i = 0; i < n; ++i) put(s, p[i]);
}
// digest as the eight big-endian state words
MP_DEV void finalize(Sha256& s, uint32_t (&digest)[8]) {
  const uint32_t bits_hi = s.total >> 29, bits_lo = s.total << 3;
  put(s, 0x80u);
#pragma unroll 1
  while (s.fill != 56) put(s, 0u);
  s.w[14] = bits_hi;
  s.w[15] = bits_lo;
  sha256_compress(s.st, s.w);
#pragma unroll
  for (int i = 0; i < 8; ++i) digest[i] = s.st[i];
}

// SHA-256 of a 32-byte message given as eight big-endian words (the digest of a digest: hash_to_scalar)
MP_DEV void sha256_of_digest(const uint32_t (&m)[8], uint32_t (&out)[8]) {
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = m[i];
  w[8] = 0x80000000u;
#pragma unroll
  for (int i = 9; i < 15; ++i) w[i] = 0;
  w[15] = 256;
  sha256_init(out);
  sha256_compress(out, w);
}

// SHA-512 of a 32-byte message given as eight big-endian 32-bit words; out = the eight 64-bit state words
MP_DEV void sha512_of_digest(const uint32_t (&m)[8], uint64_t (&out)[8]) {
  const uint64_t K[80] = {
      0x428a2f98d728ae22ull, 0x7137449123ef65cdull, 0xb5c0fbcfec4d3b2full, 0xe9b5dba58189dbbcull, 0x3956c25bf348b538ull,
      0x59f111f1b605d019ull, 0x923f82a4af194f9bull, 0xab1c5ed5da6d8118ull, 0xd807aa98a3030242ull, 0x12835b0145706fbeull,
      0x243185be4ee4b28cull, 0x550c7dc3d5ffb4e2ull, 0x72be5d74f27b896full, 0x80deb1fe3b1696b1ull, 0x9bdc06a725c71235ull,
      0xc19bf174cf692694ull, 0xe49b69c19ef14ad2ull, 0xefbe4786384f25e3ull, 0x0fc19dc68b8cd5b5ull, 0x240ca1cc77ac9c65ull,
      0x2de92c6f592b0275ull, 0x4a7484aa6ea6e483ull, 0x5cb0a9dcbd41fbd4ull, 0x76f988da831153b5ull, 0x983e5152ee66dfabull,
      0xa831c66d2db43210ull, 0xb00327c898fb213full, 0xbf597fc7beef0ee4ull, 0xc6e00bf33da88fc2ull, 0xd5a79147930aa725ull,
      0x06ca6351e003826full, 0x142929670a0e6e70ull, 0x27b70a8546d22ffcull, 0x2e1b21385c26c926ull, 0x4d2c6dfc5ac42aedull,
      0x53380d139d95b3dfull, 0x650a73548baf63deull, 0x766a0abb3c77b2a8ull, 0x81c2c92e47edaee6ull, 0x92722c851482353bull,
      0xa2bfe8a14cf10364ull, 0xa81a664bbc423001ull, 0xc24b8b70d0f89791ull, 0xc76c51a30654be30ull, 0xd192e819d6ef5218ull,
      0xd69906245565a910ull, 0xf40e35855771202aull, 0x106aa07032bbd1b8ull, 0x19a4c116b8d2d0c8ull, 0x1e376c085141ab53ull,
      0x2748774cdf8eeb99ull, 0x34b0bcb5e19b48a8ull, 0x391c0cb3c5c95a63ull, 0x4ed8aa4ae3418acbull, 0x5b9cca4f7763e373ull,
      0x682e6ff3d6b2b8a3ull, 0x748f82ee5defb2fcull, 0x78a5636f43172f60ull, 0x84c87814a1f0ab72ull, 0x8cc702081a6439ecull,
      0x90befffa23631e28ull, 0xa4506cebde82bde9ull, 0xbef9a3f7b2c67915ull, 0xc67178f2e372532bull, 0xca273eceea26619cull,
      0xd186b8c721c0c207ull, 0xeada7dd6cde0eb1eull, 0xf57d4f7fee6ed178ull, 0x06f067aa72176fbaull, 0x0a637dc5a2c898a6ull,
      0x113f9804bef90daeull, 0x1b710b35131c471bull, 0x28db77f523047d84ull, 0x32caab7b40c72493ull, 0x3c9ebe0a15c9bebcull,
      0x431d67c49c100d4cull, 0x4cc5d4becb3e42b6ull, 0x597f299cfc657e2aull, 0x5fcb6fab3ad6faecull, 0x6c44198c4a475817ull};
  uint64_t w[16];
#pragma unroll
  for (int i = 0; i < 4; ++i) w[i] = (uint64_t)m[2 * i] << 32 | m[2 * i + 1];
  w[4] = 0x8000000000000000ull;
#pragma unroll
  for (int i = 5; i < 15; ++i) w[i] = 0;
  w[15] = 256;
  uint64_t st[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                    0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
  uint64_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int i = 0; i < 80; ++i) {
    if (i >= 16) {
      const uint64_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
      const uint64_t s0 = ror64(w15, 1) ^ ror64(w15, 8) ^ (w15 >> 7);
      const uint64_t s1 = ror64(w2, 19) ^ ror64(w2, 61) ^ (w2 >> 6);
      w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
    }
    const uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41);
    const uint64_t ch = (e & f) ^ (~e & g);
    const uint64_t t1 = h + S1 + ch + K[i] + w[i & 15];
    const uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39);
    const uint64_t mj = (a & b) ^ (a & c) ^ (b & c);
    const uint64_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  out[0] = st[0] + a; out[1] = st[1] + b; out[2] = st[2] + c; out[3] = st[3] + d;
  out[4] = st[4] + e; out[5] = st[5] + f; out[6] = st[6] + g; out[7] = st[7] + h;
}

MP_DEV uint32_t bswap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }

// the four frames of one row -> running hash.  A frame whose length field exceeds its slot is clamped (rows
// come from the library's own frame kernels, so this only keeps a corrupted row from reading out of bounds).
MP_DEV void absorb_row(Sha256& s, const uint8_t* row, uint32_t slot_stride) {
#pragma unroll 1
  for (uint32_t e = 0; e < 4; ++e) {
    const uint8_t* p = row + (size_t)e * slot_stride;
    uint32_t len = (uint32_t)p[6] << 8 | p[7];
    if (len > slot_stride - 8) len = slot_stride - 8;
    update(s, p, 8 + len);
  }
}

// One thread per row: digest = SHA-256(F(h1) F(h2) F(a1) F(a2)); the challenge integer before its reduction
// is hash_to_scalar's inner hash of that digest, written as little-endian u32 limbs:
//   wide = 0: int_be(SHA-256(digest)), 8 limbs     (ModpGroup: the challenge itself, 256 < 2047 bits; secp256k1)
//   wide = 1: int_le(SHA-512(digest)), 16 limbs    (ristretto255)
// zero-filled up to out_stride limbs per row (ModpGroup scalars are 64 limbs).
struct RowHashArgs {
  const uint8_t* rows;
  uint32_t row_stride;   // bytes from one row to the next
  uint32_t slot_stride;  // bytes from one frame of a row to the next (8 + element bytes)
  uint32_t* out;         // n x out_stride limbs
  uint32_t out_stride;
  uint8_t* digest_out;   // optional: n x 32 bytes, the SHA-256 digests themselves
  uint32_t n, wide;
};
MP_DEV void row_hash_body(const RowHashArgs& A, uint32_t tid) {
  if (tid >= A.n) return;
  Sha256 s;
  init(s);
  absorb_row(s, A.rows + (size_t)tid * A.row_stride, A.slot_stride);
  uint32_t d[8];
  finalize(s, d);
  if (A.digest_out) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint8_t* o = A.digest_out + (size_t)tid * 32 + 4 * i;
      o[0] = (uint8_t)(d[i] >> 24); o[1] = (uint8_t)(d[i] >> 16); o[2] = (uint8_t)(d[i] >> 8); o[3] = (uint8_t)d[i];
    }
  }
  uint32_t* out = A.out + (size_t)tid * A.out_stride;
  uint32_t used;
  if (A.wide) {
    uint64_t h[8];
    sha512_of_digest(d, h);
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // output bytes are the words big-endian; read them as one little-endian integer
      out[2 * k] = bswap32((uint32_t)(h[k] >> 32));
      out[2 * k + 1] = bswap32((uint32_t)h[k]);
    }
    used = 16;
  } else {
    uint32_t h[8];
    sha256_of_digest(d, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = h[7 - i];
    used = 8;
  }
  for (uint32_t i = used; i < A.out_stride; ++i) out[i] = 0;
}

// The whole-box transcript as ONE chain on one device thread (rows in participant order): what moving the
// distribution transcript onto the device would cost.  Kept as a measured alternative (tunable "device_hash"),
// never the default: SHA-256 cannot be split across threads.
struct BoxHashArgs {
  const uint8_t* rows;
  uint32_t row_stride, slot_stride, n;
  uint8_t* digest_out;  // 32 bytes
};
MP_DEV void box_hash_body(const BoxHashArgs& A, uint32_t tid) {
  if (tid != 0) return;
  Sha256 s;
  init(s);
#pragma unroll 1
  for (uint32_t j = 0; j < A.n; ++j) absorb_row(s, A.rows + (size_t)j * A.row_stride, A.slot_stride);
  uint32_t d[8];
  finalize(s, d);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint8_t* o = A.digest_out + 4 * i;
    o[0] = (uint8_t)(d[i] >> 24); o[1] = (uint8_t)(d[i] >> 16); o[2] = (uint8_t)(d[i] >> 8); o[3] = (uint8_t)d[i];
  }
}

}  // namespace shadev
