// Special-form reduction for the two curve base fields, plain (non-Montgomery) representation:
//   secp256k1  p = 2^256 - 2^32 - 977      (2^256 = 2^32 + 977 mod p)
//   curve25519 p = 2^255 - 19              (2^256 = 38 mod p)
// A product is 64 IMAD.WIDE (36 for a square) plus an 8-MAC fold instead of the 128 MACs of the
// generic Montgomery product in fp256.cuh; results are fully reduced to [0, p) because the curve
// code compares field elements.  Same even/odd accumulator layout as fp256.cuh / modp_arith.cuh.
#pragma once
#include "fp256.cuh"

namespace fpsp {

using fp256::Fe;

// t[0..15] = a * b
MP_DEV void mul_wide(uint32_t (&t)[16], const Fe& a, const Fe& b) {
  uint32_t E[18], O[18];  // E[x]: column x ; O[x]: column x + 1
#pragma unroll
  for (int x = 0; x < 18; ++x) E[x] = O[x] = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ja = i & 1, jb = 1 - ja;
    {  // limbs j = ja, ja+2, ... : even columns
      const int p = i + ja;
      E[p] = simt::mad_lo_cc(a.v[ja], b.v[i], E[p]);
      E[p + 1] = simt::madc_hi_cc(a.v[ja], b.v[i], E[p + 1]);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        E[p + 2 * k] = simt::madc_lo_cc(a.v[ja + 2 * k], b.v[i], E[p + 2 * k]);
        E[p + 2 * k + 1] = simt::madc_hi_cc(a.v[ja + 2 * k], b.v[i], E[p + 2 * k + 1]);
      }
      // Column p+8 holds 0 here: the first row that reaches a new top pair adds one product to an
      // (almost) empty pair and cannot carry out, so whatever was stored before is 0 and the carry
      // is simply the new value (same argument for the odd accumulator below).
      E[p + 8] = simt::addc(0, 0);
    }
    {  // limbs j = jb, jb+2, ... : odd columns, stored at index column - 1
      const int p = i + jb - 1;
      O[p] = simt::mad_lo_cc(a.v[jb], b.v[i], O[p]);
      O[p + 1] = simt::madc_hi_cc(a.v[jb], b.v[i], O[p + 1]);
#pragma unroll
      for (int k = 1; k < 4; ++k) {
        O[p + 2 * k] = simt::madc_lo_cc(a.v[jb + 2 * k], b.v[i], O[p + 2 * k]);
        O[p + 2 * k + 1] = simt::madc_hi_cc(a.v[jb + 2 * k], b.v[i], O[p + 2 * k + 1]);
      }
      O[p + 8] = simt::addc(0, 0);
    }
  }
  t[0] = E[0];
  t[1] = simt::add_cc(E[1], O[0]);
#pragma unroll
  for (int x = 2; x < 16; ++x) t[x] = simt::addc_cc(E[x], O[x - 1]);
}

// t[0..15] = a * a : 28 off-diagonal products doubled + 8 squares
MP_DEV void sqr_wide(uint32_t (&t)[16], const Fe& a) {
  uint32_t E[18], O[18];
#pragma unroll
  for (int x = 0; x < 18; ++x) E[x] = O[x] = 0;
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    // products a_i * a_j for j > i, split by the parity of the column i + j
#pragma unroll
    for (int par = 0; par < 2; ++par) {
      const int j0 = i + 1 + ((i + 1 + i + par) & 1);  // first j > i with (i + j) % 2 == par
      if (j0 < 8) {
        const int cnt = (8 - j0 + 1) / 2;
        if (par == 0) {
          const int p = i + j0;
          E[p] = simt::mad_lo_cc(a.v[j0], a.v[i], E[p]);
          E[p + 1] = simt::madc_hi_cc(a.v[j0], a.v[i], E[p + 1]);
#pragma unroll
          for (int k = 1; k < cnt; ++k) {
            E[p + 2 * k] = simt::madc_lo_cc(a.v[j0 + 2 * k], a.v[i], E[p + 2 * k]);
            E[p + 2 * k + 1] = simt::madc_hi_cc(a.v[j0 + 2 * k], a.v[i], E[p + 2 * k + 1]);
          }
          E[p + 2 * cnt] = simt::addc_cc(E[p + 2 * cnt], 0);
          E[p + 2 * cnt + 1] = simt::addc(E[p + 2 * cnt + 1], 0);
        } else {
          const int p = i + j0 - 1;
          O[p] = simt::mad_lo_cc(a.v[j0], a.v[i], O[p]);
          O[p + 1] = simt::madc_hi_cc(a.v[j0], a.v[i], O[p + 1]);
#pragma unroll
          for (int k = 1; k < cnt; ++k) {
            O[p + 2 * k] = simt::madc_lo_cc(a.v[j0 + 2 * k], a.v[i], O[p + 2 * k]);
            O[p + 2 * k + 1] = simt::madc_hi_cc(a.v[j0 + 2 * k], a.v[i], O[p + 2 * k + 1]);
          }
          O[p + 2 * cnt] = simt::addc_cc(O[p + 2 * cnt], 0);
          O[p + 2 * cnt + 1] = simt::addc(O[p + 2 * cnt + 1], 0);
        }
      }
    }
  }
  // off = E + (O << 32); t = 2 * off + diag
  uint32_t s[16];
  s[0] = E[0];
  s[1] = simt::add_cc(E[1], O[0]);
#pragma unroll
  for (int x = 2; x < 16; ++x) s[x] = simt::addc_cc(E[x], O[x - 1]);
  t[0] = simt::add_cc(s[0], s[0]);
#pragma unroll
  for (int x = 1; x < 16; ++x) t[x] = simt::addc_cc(s[x], s[x]);
  t[0] = simt::mad_lo_cc(a.v[0], a.v[0], t[0]);
  t[1] = simt::madc_hi_cc(a.v[0], a.v[0], t[1]);
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    t[2 * i] = simt::madc_lo_cc(a.v[i], a.v[i], t[2 * i]);
    t[2 * i + 1] = simt::madc_hi_cc(a.v[i], a.v[i], t[2 * i + 1]);
  }
}

// r[0..8] + r9 * 2^288 = lo[0..7] + hi[0..7] * k (+ hi * 2^32 when SHIFTED), k < 2^32.  lo and the
// shifted copy of hi are the start values of the even / odd column accumulators, so the whole sum
// costs the 8 MACs and one combining pass.
template <bool SHIFTED>
MP_DEV void fold(uint32_t (&r)[9], uint32_t& r9, const uint32_t* lo, const uint32_t* hi, uint32_t k) {
  uint32_t E[9], O[9];  // E[x]: column x ; O[x]: column x + 1
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    E[x] = lo[x];
    O[x] = SHIFTED ? hi[x] : 0u;
  }
  E[0] = simt::mad_lo_cc(hi[0], k, E[0]);
  E[1] = simt::madc_hi_cc(hi[0], k, E[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    E[j] = simt::madc_lo_cc(hi[j], k, E[j]);
    E[j + 1] = simt::madc_hi_cc(hi[j], k, E[j + 1]);
  }
  E[8] = simt::addc(0, 0);
  O[0] = simt::mad_lo_cc(hi[1], k, O[0]);
  O[1] = simt::madc_hi_cc(hi[1], k, O[1]);
#pragma unroll
  for (int j = 3; j < 8; j += 2) {
    O[j - 1] = simt::madc_lo_cc(hi[j], k, O[j - 1]);
    O[j] = simt::madc_hi_cc(hi[j], k, O[j]);
  }
  O[8] = simt::addc(0, 0);
  r[0] = E[0];
  r[1] = simt::add_cc(E[1], O[0]);
#pragma unroll
  for (int x = 2; x < 9; ++x) r[x] = simt::addc_cc(E[x], O[x - 1]);
  r9 = simt::addc(O[8], 0);
}

// The two primes as compile-time limbs: every use below unrolls to immediates, no loads.
struct SecpP {
  MP_DEV static constexpr uint32_t limb(int i) { return i == 0 ? 0xFFFFFC2Fu : (i == 1 ? 0xFFFFFFFEu : 0xFFFFFFFFu); }
};
struct EdP {
  MP_DEV static constexpr uint32_t limb(int i) { return i == 0 ? 0xFFFFFFEDu : (i == 7 ? 0x7FFFFFFFu : 0xFFFFFFFFu); }
};

// r = a - p if a >= p (a < 2p, with an optional carry bit above the 8 limbs)
template <class Pr>
MP_DEV Fe cond_sub_p(const Fe& a, uint32_t carry) {
  Fe t;
  t.v[0] = simt::sub_cc(a.v[0], Pr::limb(0));
#pragma unroll
  for (int i = 1; i < 8; ++i) t.v[i] = simt::subc_cc(a.v[i], Pr::limb(i));
  uint32_t borrow = simt::subc(0, 0);
  bool take = carry || borrow == 0;
  Fe r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = take ? t.v[i] : a.v[i];
  return r;
}
template <class Pr>
MP_DEV Fe add_p(const Fe& a, const Fe& b) {
  Fe s;
  s.v[0] = simt::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) s.v[i] = simt::addc_cc(a.v[i], b.v[i]);
  uint32_t c = simt::addc(0, 0);
  return cond_sub_p<Pr>(s, c);
}
template <class Pr>
MP_DEV Fe sub_p(const Fe& a, const Fe& b) {
  Fe d;
  d.v[0] = simt::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) d.v[i] = simt::subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = simt::subc(0, 0);  // 0xffffffff when a < b
  Fe r;
  r.v[0] = simt::add_cc(d.v[0], Pr::limb(0) & borrow);
#pragma unroll
  for (int i = 1; i < 8; ++i) r.v[i] = simt::addc_cc(d.v[i], Pr::limb(i) & borrow);
  return r;
}

// ---- secp256k1: 2^256 = 2^32 + 977 (mod p) -----------------------------------------------------
MP_DEV Fe secp_reduce(const uint32_t (&t)[16]) {
  // r = lo + hi*977 + (hi << 32)   (< 2^290)
  uint32_t r[9], r9;
  fold<true>(r, r9, t, t + 8, 977u);
  // second fold of the part above 2^256: top = r[8] + r9 * 2^32 (< 2^34),
  // top * (2^32 + 977) = top*977 + (top << 32), three words w0..w2
  uint32_t w0 = simt::mul_lo(r[8], 977u), w1 = simt::mul_hi(r[8], 977u) + r9 * 977u;
  w1 = simt::add_cc(w1, r[8]);
  uint32_t w2 = simt::addc(r9, 0);
  Fe s;
  s.v[0] = simt::add_cc(r[0], w0);
  s.v[1] = simt::addc_cc(r[1], w1);
  s.v[2] = simt::addc_cc(r[2], w2);
#pragma unroll
  for (int x = 3; x < 8; ++x) s.v[x] = simt::addc_cc(r[x], 0);
  uint32_t c1 = simt::addc(0, 0);
  // a carry out of 2^256 leaves a tiny s; 2^256 + s - p is what the conditional subtraction returns
  return cond_sub_p<SecpP>(s, c1);
}

// ---- curve25519: 2^256 = 38 (mod p), p = 2^255 - 19 --------------------------------------------
MP_DEV Fe ed_reduce(const uint32_t (&t)[16]) {
  uint32_t r[9], r9;
  fold<false>(r, r9, t, t + 8, 38u);  // < 39 * 2^256, r9 = 0
  (void)r9;
  // fold r[8] (< 39) and bit 255: value = low255 + 19 * (2*r[8] + bit255)
  uint32_t top = (r[8] << 1) | (r[7] >> 31);
  Fe s;
  s.v[0] = simt::add_cc(r[0], top * 19u);
#pragma unroll
  for (int x = 1; x < 7; ++x) s.v[x] = simt::addc_cc(r[x], 0);
  s.v[7] = simt::addc(r[7] & 0x7fffffffu, 0);
  // s < 2^255 + 19*79: at most one more wrap of bit 255
  uint32_t b = s.v[7] >> 31;
  s.v[7] &= 0x7fffffffu;
  s.v[0] = simt::add_cc(s.v[0], 19u * b);
#pragma unroll
  for (int x = 1; x < 8; ++x) s.v[x] = simt::addc_cc(s.v[x], 0);
  return cond_sub_p<EdP>(s, 0);
}

}  // namespace fpsp
