// ristretto255 group operations, one point per thread (replaces curve25519-dalek's
// RistrettoPoint * Scalar / + / compress / decompress behind
// /root/reference/src/groups/ristretto255.rs:161-220).  Extended twisted-Edwards coordinates
// (a = -1) over the field 2^255-19 in plain representation (fpspecial.cuh); encoding and decoding follow
// RFC 9496 sections 4.3.1-4.3.2.
#pragma once
#include "fpspecial.cuh"
#ifdef __CUDACC__
#pragma nv_diag_suppress 550  // range checks keep only the borrow of a subtraction
#endif

namespace rist {

using fp256::Fe;
using fp256::Modulus;

// Base-field arithmetic mod 2^255-19: plain representation, special-form reduction.
namespace F {
using fp256::eq;
using fp256::fe_zero;
using fp256::is_zero;
using fp256::load;
using fp256::store;
// Field constants are immediates (fpsp::EdP); the Modulus argument is kept so the curve code
// reads the same for both fields.
MP_DEV Fe add(const Fe& a, const Fe& b, const Modulus&) { return fpsp::add_p<fpsp::EdP>(a, b); }
MP_DEV Fe sub(const Fe& a, const Fe& b, const Modulus&) { return fpsp::sub_p<fpsp::EdP>(a, b); }
MP_DEV Fe neg(const Fe& a, const Modulus&) { return fpsp::sub_p<fpsp::EdP>(fe_zero(), a); }
MP_DEV Fe dbl(const Fe& a, const Modulus&) { return fpsp::add_p<fpsp::EdP>(a, a); }
// Operands by value: a non-inlined callee that reads its operands through references keeps ptxas
// from pairing mad.lo.cc / madc.hi.cc into IMAD.WIDE (twice the instructions); by value they arrive
// in registers.
MP_NOINLINE Fe mul(Fe a, Fe b) {
  uint32_t t[16];
  fpsp::mul_wide(t, a, b);
  return fpsp::ed_reduce(t);
}
MP_NOINLINE Fe sqr(Fe a) {
  uint32_t t[16];
  fpsp::sqr_wide(t, a);
  return fpsp::ed_reduce(t);
}
MP_DEV Fe mul(const Fe& a, const Fe& b, const Modulus&) { return mul(a, b); }
MP_DEV Fe sqr(const Fe& a, const Modulus&) { return sqr(a); }
MP_DEV Fe to_mont(const Fe& a, const Modulus&) { return a; }
MP_DEV Fe from_mont(const Fe& a, const Modulus&) { return a; }
MP_DEV Fe mont_one(const Modulus&) {
  Fe r = fe_zero();
  r.v[0] = 1;
  return r;
}
MP_NOINLINE Fe sqn(Fe a, int n) {
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = F::sqr(a);
  return a;
}
// a^((p-5)/8) = a^(2^252 - 3): the classic curve25519 chain, 251 squarings + 11 multiplications instead
// of the 252 + ~250 of bit-by-bit square-and-multiply
MP_NOINLINE Fe pow_p58(const Fe& a) {
  Fe t0 = F::sqr(a);                      // 2
  Fe t1 = F::mul(sqn(t0, 2), a);          // 9
  t0 = F::mul(t0, t1);                    // 11
  t0 = F::mul(F::sqr(t0), t1);            // 31 = 2^5 - 1
  t1 = F::mul(sqn(t0, 5), t0);            // 2^10 - 1
  Fe t2 = F::mul(sqn(t1, 10), t1);        // 2^20 - 1
  t2 = F::mul(sqn(t2, 20), t2);           // 2^40 - 1
  t1 = F::mul(sqn(t2, 10), t1);           // 2^50 - 1
  t2 = F::mul(sqn(t1, 50), t1);           // 2^100 - 1
  Fe t3 = F::mul(sqn(t2, 100), t2);       // 2^200 - 1
  t1 = F::mul(sqn(t3, 50), t1);           // 2^250 - 1
  return F::mul(sqn(t1, 2), a);           // 2^252 - 3
}
// a^(p-2) = (a^((p-5)/8))^8 * a^3
MP_NOINLINE Fe inv(const Fe& a) { return F::mul(sqn(pow_p58(a), 3), F::mul(F::sqr(a), a)); }
}  // namespace F

struct Consts {
  Modulus P;                   // 2^255 - 19
  Modulus N;                   // l = 2^252 + 27742317777372353535851937790883648493
  uint32_t d[8], d2[8];        // d, 2d
  uint32_t sqrt_m1[8];         // sqrt(-1)
  uint32_t invsqrt_a_minus_d[8];
  uint32_t bx[8], by[8];       // basepoint, affine
  uint32_t pm5d8[8];           // (p - 5) / 8
};

struct Ext {
  Fe X, Y, Z, T;
};
struct Aff {  // affine Edwards coordinates of a representative
  Fe x, y;
  uint32_t inf;  // unused (the identity (0, 1) is an ordinary point); kept for the policy interface
};

MP_DEV Ext ext_identity(const Modulus& P) {
  Ext r;
  r.X = F::fe_zero();
  r.Y = F::mont_one(P);
  r.Z = F::mont_one(P);
  r.T = F::fe_zero();
  return r;
}
MP_DEV Ext ext_from_aff(const Aff& a, const Modulus& P) {
  Ext r;
  r.X = a.x;
  r.Y = a.y;
  r.Z = F::mont_one(P);
  r.T = F::mul(a.x, a.y, P);
  return r;
}

// add-2008-hwcd-3 (unified, complete for a = -1 and non-square d): 9M
MP_NOINLINE Ext ext_add(Ext p, Ext q, const Consts& C) {
  using namespace F;
  const Modulus& P = C.P;
  Fe A = F::mul(F::sub(p.Y, p.X, P), F::sub(q.Y, q.X, P), P);
  Fe B = F::mul(F::add(p.Y, p.X, P), F::add(q.Y, q.X, P), P);
  Fe Cc = F::mul(F::mul(p.T, load(C.d2), P), q.T, P);
  Fe D = F::dbl(F::mul(p.Z, q.Z, P), P);
  Fe E = F::sub(B, A, P), F = F::sub(D, Cc, P), G = F::add(D, Cc, P), H = F::add(B, A, P);
  Ext r;
  r.X = F::mul(E, F, P);
  r.Y = F::mul(G, H, P);
  r.T = F::mul(E, H, P);
  r.Z = F::mul(F, G, P);
  return r;
}
// dbl-2008-hwcd with a = -1: 4M + 4S
MP_NOINLINE Ext ext_dbl(Ext p, const Consts& C, bool need_t = true) {
  using namespace F;
  const Modulus& P = C.P;
  Fe A = F::sqr(p.X, P), B = F::sqr(p.Y, P), Cc = F::dbl(F::sqr(p.Z, P), P);
  Fe D = F::neg(A, P);
  Fe E = F::sub(F::sub(F::sqr(F::add(p.X, p.Y, P), P), A, P), B, P);
  Fe G = F::add(D, B, P), F = F::sub(G, Cc, P), H = F::sub(D, B, P);
  Ext r;
  r.X = F::mul(E, F, P);
  r.Y = F::mul(G, H, P);
  // T is only read by additions: a doubling that feeds another doubling skips it (3M + 4S)
  r.T = need_t ? F::mul(E, H, P) : F::fe_zero();
  r.Z = F::mul(F, G, P);
  return r;
}

// [k]p for a small scalar k < 4^nd, fixed 2-bit windows (the generic ladder of ec_kernels.cuh with
// the T coordinate skipped in the first doubling of every window)
template <class DigitFn>
MP_DEV Ext small_mul_ext(const Ext& p, DigitFn digit, uint32_t nd, const Consts& C) {
  Ext t2 = ext_dbl(p, C), t3 = ext_add(t2, p, C);
  Ext acc = ext_identity(C.P);
#pragma unroll 1
  for (int s = (int)nd - 1; s >= 0; --s) {
    if (s != (int)nd - 1) {
      acc = ext_dbl(acc, C, false);
      acc = ext_dbl(acc, C);
    }
    uint32_t d = digit(s);
    Ext q = (d == 3) ? t3 : ((d == 2) ? t2 : p);
    if (d) acc = ext_add(acc, q, C);
  }
  return acc;
}

// canonical-form helpers (RFC 9496 section 4.1)
MP_DEV bool is_negative(const Fe& a_mont, const Modulus& P) { return F::from_mont(a_mont, P).v[0] & 1u; }
MP_DEV Fe ct_abs(const Fe& a, const Modulus& P) { return is_negative(a, P) ? F::neg(a, P) : a; }

// SQRT_RATIO_M1 (RFC 9496 section 4.2); returns was_square, r in *out
MP_NOINLINE bool sqrt_ratio_m1(Fe* out, const Fe& u, const Fe& v, const Consts& C) {
  using namespace F;
  const Modulus& P = C.P;
  Fe v3 = F::mul(F::sqr(v, P), v, P);
  Fe v7 = F::mul(F::sqr(v3, P), v, P);
  Fe r = F::mul(F::mul(u, v3, P), F::pow_p58(F::mul(u, v7, P)), P);
  Fe check = F::mul(v, F::sqr(r, P), P);
  Fe sm1 = load(C.sqrt_m1);
  Fe nu = F::neg(u, P);
  bool correct = eq(check, u);
  bool flipped = eq(check, nu);
  bool flipped_i = eq(check, F::mul(nu, sm1, P));
  if (flipped || flipped_i) r = F::mul(r, sm1, P);
  *out = ct_abs(r, P);
  return correct || flipped;
}

// RFC 9496 section 4.3.2
MP_NOINLINE void encode(uint8_t* out, const Ext& p, const Consts& C) {
  using namespace F;
  const Modulus& P = C.P;
  Fe u1 = F::mul(F::add(p.Z, p.Y, P), F::sub(p.Z, p.Y, P), P);
  Fe u2 = F::mul(p.X, p.Y, P);
  Fe invsqrt;
  sqrt_ratio_m1(&invsqrt, F::mont_one(P), F::mul(u1, F::sqr(u2, P), P), C);
  Fe den1 = F::mul(invsqrt, u1, P), den2 = F::mul(invsqrt, u2, P);
  Fe z_inv = F::mul(F::mul(den1, den2, P), p.T, P);
  Fe sm1 = load(C.sqrt_m1);
  Fe ix0 = F::mul(p.X, sm1, P), iy0 = F::mul(p.Y, sm1, P);
  Fe enchanted = F::mul(den1, load(C.invsqrt_a_minus_d), P);
  bool rotate = is_negative(F::mul(p.T, z_inv, P), P);
  Fe x = rotate ? iy0 : p.X;
  Fe y = rotate ? ix0 : p.Y;
  Fe den_inv = rotate ? enchanted : den2;
  if (is_negative(F::mul(x, z_inv, P), P)) y = F::neg(y, P);
  Fe s = F::from_mont(ct_abs(F::mul(den_inv, F::sub(p.Z, y, P), P), P), P);
  for (int i = 0; i < 8; ++i) {
    out[4 * i] = (uint8_t)s.v[i];
    out[4 * i + 1] = (uint8_t)(s.v[i] >> 8);
    out[4 * i + 2] = (uint8_t)(s.v[i] >> 16);
    out[4 * i + 3] = (uint8_t)(s.v[i] >> 24);
  }
}

// RFC 9496 section 4.3.1; false for non-canonical or invalid encodings
MP_NOINLINE bool decode(Aff& a, const uint8_t* in, const Consts& C) {
  using namespace F;
  const Modulus& P = C.P;
  a.inf = 0;
  a.x = fe_zero();
  a.y = F::mont_one(P);
  Fe s;
  for (int i = 0; i < 8; ++i)
    s.v[i] = (uint32_t)in[4 * i] | (uint32_t)in[4 * i + 1] << 8 | (uint32_t)in[4 * i + 2] << 16 |
             (uint32_t)in[4 * i + 3] << 24;
  // canonical: s < p and s non-negative (even)
  Fe t;
  t.v[0] = simt::sub_cc(s.v[0], P.m[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) t.v[i] = simt::subc_cc(s.v[i], P.m[i]);
  if (simt::subc(0, 0) == 0) return false;
  if (s.v[0] & 1u) return false;
  Fe sm = F::to_mont(s, P), one = F::mont_one(P);
  Fe ss = F::sqr(sm, P);
  Fe u1 = F::sub(one, ss, P), u2 = F::add(one, ss, P);
  Fe u2_sqr = F::sqr(u2, P);
  Fe v = F::sub(F::neg(F::mul(load(C.d), F::sqr(u1, P), P), P), u2_sqr, P);
  Fe invsqrt;
  bool was_square = sqrt_ratio_m1(&invsqrt, one, F::mul(v, u2_sqr, P), C);
  Fe den_x = F::mul(invsqrt, u2, P);
  Fe den_y = F::mul(F::mul(invsqrt, den_x, P), v, P);
  Fe x = ct_abs(F::mul(F::dbl(sm, P), den_x, P), P);
  Fe y = F::mul(u1, den_y, P);
  Fe tt = F::mul(x, y, P);
  if (!was_square || is_negative(tt, P) || is_zero(y)) return false;
  a.x = x;
  a.y = y;
  return true;
}

// ---- curve policy for ec_kernels.cuh ---------------------------------------------------------
struct RistCurve {
  using Consts = rist::Consts;
  using Point = Ext;
  using Affine = Aff;
  static constexpr int EB = 32;
  MP_DEV static Point small_mul(const Point& p, uint32_t k, uint32_t nd, const Consts& C) {
    return small_mul_ext(p, [k](int s) { return (k >> (2 * s)) & 3u; }, nd, C);
  }
  // [e]p for a full-width scalar (8 little-endian limbs): the same 2-bit windows, table in registers
  MP_DEV static Point scalar_mul_wide(const Point& p, const uint32_t* e, const Consts& C) {
    return small_mul_ext(p, [e](int s) { return (e[s >> 4] >> ((s & 15) * 2)) & 3u; }, 128, C);
  }
  MP_DEV static Point infinity(const Consts& C) { return ext_identity(C.P); }
  MP_DEV static Point from_aff(const Affine& a, const Consts& C) { return ext_from_aff(a, C.P); }
  MP_DEV static Point dbl(const Point& p, const Consts& C) { return ext_dbl(p, C); }
  MP_DEV static Point add(const Point& p, const Point& q, const Consts& C) { return ext_add(p, q, C); }
  MP_DEV static Point madd(const Point& p, const Affine& q, const Consts& C) {
    return ext_add(p, ext_from_aff(q, C.P), C);
  }
  MP_DEV static Affine to_affine(const Point& p, const Consts&) {
    Affine a;
    const Fe zi = F::inv(p.Z);
    a.x = F::mul(p.X, zi);
    a.y = F::mul(p.Y, zi);
    a.inf = 0;
    return a;
  }
  MP_DEV static bool decode(Affine& a, const uint8_t* in, const Consts& C) { return rist::decode(a, in, C); }
  MP_DEV static void encode(uint8_t* out, const Point& p, const Consts& C) { rist::encode(out, p, C); }
};

}  // namespace rist
