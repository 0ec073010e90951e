// 256-bit prime-field arithmetic, one field element per thread (8 little-endian u32
// limbs in registers), Montgomery representation with R = 2^256.  One template serves
// the four 256-bit moduli of the elliptic-curve groups:
//   secp256k1 base field p and scalar field n   (replaces k256's FieldElement / Scalar)
//   curve25519 base field 2^255-19 and scalar field l (replaces curve25519-dalek's)
// The product uses the same even/odd 64-bit-column accumulators as modp_arith.cuh so
// that every 32x32->64 MAC is one IMAD.WIDE.U32(.X) with the carry in a predicate.
#pragma once
#include "simt.h"

namespace fp256 {

struct Modulus {
  uint32_t m[8];    // modulus
  uint32_t r2[8];   // R^2 mod m
  uint32_t one[8];  // R mod m
  uint32_t np;      // -m^-1 mod 2^32
};

struct Fe {
  uint32_t v[8];
};

MP_DEV Fe fe_zero() {
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
  return r;
}
MP_DEV bool is_zero(const Fe& a) {
  uint32_t x = a.v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) x |= a.v[i];
  return x == 0;
}
MP_DEV bool eq(const Fe& a, const Fe& b) {
  uint32_t x = a.v[0] ^ b.v[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) x |= a.v[i] ^ b.v[i];
  return x == 0;
}
MP_DEV Fe load(const uint32_t* p) {
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = p[i];
  return r;
}
MP_DEV void store(uint32_t* p, const Fe& a) {
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = a.v[i];
}

// r = a - m if a >= m (a < 2m, with an optional carry bit above the 8 limbs)
MP_DEV Fe cond_sub(const Fe& a, uint32_t carry, const Modulus& M) {
  Fe t;
  t.v[0] = simt::sub_cc(a.v[0], M.m[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) t.v[i] = simt::subc_cc(a.v[i], M.m[i]);
  uint32_t borrow = simt::subc(0, 0);  // 0xffffffff if a < m
  bool take = carry || borrow == 0;
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = take ? t.v[i] : a.v[i];
  return r;
}
MP_DEV Fe add(const Fe& a, const Fe& b, const Modulus& M) {
  Fe s;
  s.v[0] = simt::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) s.v[i] = simt::addc_cc(a.v[i], b.v[i]);
  uint32_t c = simt::addc(0, 0);
  return cond_sub(s, c, M);
}
MP_DEV Fe sub(const Fe& a, const Fe& b, const Modulus& M) {
  Fe d;
  d.v[0] = simt::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) d.v[i] = simt::subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = simt::subc(0, 0);  // 0xffffffff when a < b
  Fe r;
  r.v[0] = simt::add_cc(d.v[0], M.m[0] & borrow);
#pragma unroll
  for (int i = 1; i < 8; ++i) r.v[i] = simt::addc_cc(d.v[i], M.m[i] & borrow);
  return r;
}
MP_DEV Fe neg(const Fe& a, const Modulus& M) { return sub(fe_zero(), a, M); }
MP_DEV Fe dbl(const Fe& a, const Modulus& M) { return add(a, a, M); }

// One 32-bit digit of the Montgomery product (thread-local form of modp::mm_digit).
template <bool FIRST>
MP_DEV void mm_digit(uint32_t (&P)[10], uint32_t (&S)[10], const Fe& a, uint32_t b, const Modulus& M) {
  if (FIRST) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      P[i] = simt::mul_lo(a.v[i], b);
      P[i + 1] = simt::mul_hi(a.v[i], b);
      S[i] = simt::mul_lo(a.v[i + 1], b);
      S[i + 1] = simt::mul_hi(a.v[i + 1], b);
    }
    P[8] = P[9] = 0;
    S[8] = S[9] = 0;
  } else {
    P[0] = simt::add_cc(P[0], S[1]);
#pragma unroll
    for (int x = 0; x < 8; x += 2) {
      S[x] = simt::madc_lo_cc(a.v[x + 1], b, S[x + 2]);
      S[x + 1] = simt::madc_hi_cc(a.v[x + 1], b, S[x + 3]);
    }
    S[8] = simt::addc(0, 0);
    S[9] = 0;
    P[0] = simt::mad_lo_cc(a.v[0], b, P[0]);
    P[1] = simt::madc_hi_cc(a.v[0], b, P[1]);
#pragma unroll
    for (int i = 2; i < 8; i += 2) {
      P[i] = simt::madc_lo_cc(a.v[i], b, P[i]);
      P[i + 1] = simt::madc_hi_cc(a.v[i], b, P[i + 1]);
    }
    P[8] = simt::addc(P[8], 0);
  }
  uint32_t m = simt::mul_lo(P[0], M.np);
  S[0] = simt::mad_lo_cc(M.m[1], m, S[0]);
  S[1] = simt::madc_hi_cc(M.m[1], m, S[1]);
#pragma unroll
  for (int i = 3; i < 8; i += 2) {
    S[i - 1] = simt::madc_lo_cc(M.m[i], m, S[i - 1]);
    S[i] = simt::madc_hi_cc(M.m[i], m, S[i]);
  }
  S[8] = simt::addc(S[8], 0);
  P[0] = simt::mad_lo_cc(M.m[0], m, P[0]);
  P[1] = simt::madc_hi_cc(M.m[0], m, P[1]);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    P[i] = simt::madc_lo_cc(M.m[i], m, P[i]);
    P[i + 1] = simt::madc_hi_cc(M.m[i], m, P[i + 1]);
  }
  P[8] = simt::addc(P[8], 0);
}

// r = a * b / R mod m, fully reduced
MP_DEV Fe mul(const Fe& a, const Fe& b, const Modulus& M) {
  uint32_t A0[10], A1[10];
  mm_digit<true>(A0, A1, a, b.v[0], M);
  mm_digit<false>(A1, A0, a, b.v[1], M);
  mm_digit<false>(A0, A1, a, b.v[2], M);
  mm_digit<false>(A1, A0, a, b.v[3], M);
  mm_digit<false>(A0, A1, a, b.v[4], M);
  mm_digit<false>(A1, A0, a, b.v[5], M);
  mm_digit<false>(A0, A1, a, b.v[6], M);
  mm_digit<false>(A1, A0, a, b.v[7], M);
  // value = S + (P >> 32) with P = A1, S = A0
  Fe r;
  r.v[0] = simt::add_cc(A0[0], A1[1]);
#pragma unroll
  for (int i = 1; i < 8; ++i) r.v[i] = simt::addc_cc(A0[i], A1[i + 1]);
  uint32_t ov = simt::addc(A0[8], A1[9]);
  return cond_sub(r, ov, M);
}
MP_DEV Fe sqr(const Fe& a, const Modulus& M) { return mul(a, a, M); }

MP_DEV Fe to_mont(const Fe& a, const Modulus& M) {
  Fe r2 = load(M.r2);
  return mul(a, r2, M);
}
MP_DEV Fe from_mont(const Fe& a, const Modulus& M) {
  Fe one = fe_zero();
  one.v[0] = 1;
  return mul(a, one, M);
}
MP_DEV Fe mont_one(const Modulus& M) { return load(M.one); }

// a^e for a 256-bit exponent given as 8 limbs (plain binary ladder; used for inversion and
// square roots, a few hundred field multiplications per call)
MP_NOINLINE Fe pow(const Fe& a, const uint32_t (&e)[8], const Modulus& M) {
  Fe r = mont_one(M);
  bool started = false;
#pragma unroll 1
  for (int i = 255; i >= 0; --i) {
    if (started) r = sqr(r, M);
    if ((e[i >> 5] >> (i & 31)) & 1u) {
      r = started ? mul(r, a, M) : a;
      started = true;
    }
  }
  return r;
}
// a^-1 = a^(m-2)  (m prime); 0 -> 0
MP_NOINLINE Fe inv(const Fe& a, const Modulus& M) {
  uint32_t e[8];
  e[0] = simt::sub_cc(M.m[0], 2);
#pragma unroll
  for (int i = 1; i < 8; ++i) e[i] = simt::subc_cc(M.m[i], 0);
  return pow(a, e, M);
}

}  // namespace fp256
