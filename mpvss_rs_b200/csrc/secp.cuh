// secp256k1 group operations, one point per thread (replaces k256's
// ProjectivePoint * Scalar / + / to_bytes / from_bytes behind
// /root/reference/src/groups/secp256k1.rs:91-152).  Jacobian coordinates over the base
// field in plain representation (special-form reduction, fpspecial.cuh); Z = 0 is the identity.  Kernel bodies are
// written against simt.h, so tests/emu runs them on the CPU as well.
#pragma once
#include "fpspecial.cuh"
#ifdef __CUDACC__
#pragma nv_diag_suppress 550  // range checks keep only the borrow of a subtraction
#endif

namespace secp {

using fp256::Fe;
using fp256::Modulus;

// Base-field arithmetic: plain representation with the special-form reduction of fpspecial.cuh
// (the scalar field keeps the generic Montgomery code of fp256.cuh).
namespace F {
using fp256::eq;
using fp256::fe_zero;
using fp256::is_zero;
using fp256::load;
using fp256::store;
// Field constants are immediates (fpsp::SecpP); the Modulus argument is kept so the curve code
// reads the same for both fields.
MP_DEV Fe add(const Fe& a, const Fe& b, const Modulus&) { return fpsp::add_p<fpsp::SecpP>(a, b); }
MP_DEV Fe sub(const Fe& a, const Fe& b, const Modulus&) { return fpsp::sub_p<fpsp::SecpP>(a, b); }
MP_DEV Fe neg(const Fe& a, const Modulus&) { return fpsp::sub_p<fpsp::SecpP>(fe_zero(), a); }
MP_DEV Fe dbl(const Fe& a, const Modulus&) { return fpsp::add_p<fpsp::SecpP>(a, a); }
// Operands by value: a non-inlined callee that reads its operands through references keeps ptxas
// from pairing mad.lo.cc / madc.hi.cc into IMAD.WIDE (twice the instructions); by value they arrive
// in registers.
MP_DEV Fe mul_inl(const Fe& a, const Fe& b) {
  uint32_t t[16];
  fpsp::mul_wide(t, a, b);
  return fpsp::secp_reduce(t);
}
MP_DEV Fe sqr_inl(const Fe& a) {
  uint32_t t[16];
  fpsp::sqr_wide(t, a);
  return fpsp::secp_reduce(t);
}
MP_NOINLINE Fe mul(Fe a, Fe b) { return mul_inl(a, b); }
MP_NOINLINE Fe sqr(Fe a) { return sqr_inl(a); }
MP_DEV Fe mul(const Fe& a, const Fe& b, const Modulus&) { return mul(a, b); }
MP_DEV Fe sqr(const Fe& a, const Modulus&) { return sqr(a); }
MP_DEV Fe to_mont(const Fe& a, const Modulus&) { return a; }
MP_DEV Fe from_mont(const Fe& a, const Modulus&) { return a; }
MP_DEV Fe mont_one(const Modulus&) {
  Fe r = fe_zero();
  r.v[0] = 1;
  return r;
}
MP_NOINLINE Fe pow(const Fe& a, const uint32_t (&e)[8], const Modulus& P) {
  Fe r = F::mont_one(P);
  bool started = false;
#pragma unroll 1
  for (int i = 255; i >= 0; --i) {
    if (started) r = F::sqr(r, P);
    if ((e[i >> 5] >> (i & 31)) & 1u) {
      r = started ? F::mul(r, a, P) : a;
      started = true;
    }
  }
  return r;
}
// a^(2^n) by n squarings
MP_NOINLINE Fe sqn(Fe a, int n) {
#pragma unroll 1
  for (int i = 0; i < n; ++i) a = F::sqr(a);
  return a;
}
// Shared prefix of the inversion and square-root chains: x2 = a^(2^2 - 1), x22 = a^(2^22 - 1),
// x223 = a^(2^223 - 1) (the exponents p - 2 and (p + 1) / 4 are runs of ones around the 2^32 + 977 gap;
// 238 squarings + 10 multiplications).
struct Runs {
  Fe x2, x22, x223;
};
MP_NOINLINE Runs runs_of_ones(const Fe& a) {
  Runs r;
  r.x2 = F::mul(sqn(a, 1), a);
  Fe x3 = F::mul(sqn(r.x2, 1), a);
  Fe x6 = F::mul(sqn(x3, 3), x3);
  Fe x9 = F::mul(sqn(x6, 3), x3);
  Fe x11 = F::mul(sqn(x9, 2), r.x2);
  r.x22 = F::mul(sqn(x11, 11), x11);
  Fe x44 = F::mul(sqn(r.x22, 22), r.x22);
  Fe x88 = F::mul(sqn(x44, 44), x44);
  Fe x176 = F::mul(sqn(x88, 88), x88);
  Fe x220 = F::mul(sqn(x176, 44), x44);
  r.x223 = F::mul(sqn(x220, 3), x3);
  return r;
}
// a^(p-2): 255 squarings + 15 multiplications instead of the 256 + ~250 of bit-by-bit square-and-multiply
// (p - 2 = 2^256 - 2^32 - 979 = [223 ones] 0 [22 ones] 0000 1 0 11 01 in binary)
MP_NOINLINE Fe inv(const Fe& a, const Modulus&) {
  Runs r = runs_of_ones(a);
  Fe t = F::mul(sqn(r.x223, 23), r.x22);
  t = F::mul(sqn(t, 5), a);
  t = F::mul(sqn(t, 3), r.x2);
  return F::mul(sqn(t, 2), a);
}
// a^((p+1)/4), a square root of a when a is a square (p = 3 mod 4): 253 squarings + 13 multiplications
// ((p + 1) / 4 = 2^254 - 2^30 - 244 = [223 ones] 0 [22 ones] 0000 11 00)
MP_NOINLINE Fe sqrt_candidate(const Fe& a, const Modulus&) {
  Runs r = runs_of_ones(a);
  Fe t = F::mul(sqn(r.x223, 23), r.x22);
  t = F::mul(sqn(t, 6), r.x2);
  return sqn(t, 2);
}
}  // namespace F

// Field products inside the point doubling / mixed addition: called (default) or inlined (EC_INLINE_FIELD:
// no argument moves and CALL/RET, the scheduler sees across products, at 3x the code size).
namespace FX {
#ifdef EC_INLINE_FIELD
MP_DEV Fe mul(const Fe& a, const Fe& b, const Modulus&) { return F::mul_inl(a, b); }
MP_DEV Fe sqr(const Fe& a, const Modulus&) { return F::sqr_inl(a); }
#else
MP_DEV Fe mul(const Fe& a, const Fe& b, const Modulus& P) { return F::mul(a, b, P); }
MP_DEV Fe sqr(const Fe& a, const Modulus& P) { return F::sqr(a, P); }
#endif
}  // namespace FX

// Device-resident constants (built on the host in secp_api.cu).
struct Consts {
  Modulus P;             // base field  2^256 - 2^32 - 977
  Modulus N;             // scalar field (group order)
  uint32_t b7[8];        // 7
  uint32_t gx[8], gy[8]; // generator, affine
  uint32_t sqrt_e[8];    // (p + 1) / 4
};

struct Jac {
  Fe X, Y, Z;
};
struct Aff {  // affine; inf = identity
  Fe x, y;
  uint32_t inf;
};

MP_DEV Jac jac_infinity(const Modulus& P) {
  Jac r;
  r.X = F::mont_one(P);
  r.Y = F::mont_one(P);
  r.Z = F::fe_zero();
  return r;
}
MP_DEV bool jac_is_inf(const Jac& p) { return F::is_zero(p.Z); }
MP_DEV Jac jac_from_aff(const Aff& a, const Modulus& P) {
  if (a.inf) return jac_infinity(P);
  Jac r;
  r.X = a.x;
  r.Y = a.y;
  r.Z = F::mont_one(P);
  return r;
}
MP_DEV Jac jac_neg(const Jac& p, const Modulus& P) {
  Jac r = p;
  r.Y = F::neg(p.Y, P);
  return r;
}

// dbl-2009-l (a = 0): 2M + 5S
MP_NOINLINE Jac jac_dbl(Jac p, const Modulus& P) {
  using namespace F;
  if (jac_is_inf(p)) return p;
  Fe A = FX::sqr(p.X, P), B = FX::sqr(p.Y, P), C = FX::sqr(B, P);
  Fe t = F::add(p.X, B, P);
  Fe D = F::dbl(F::sub(F::sub(FX::sqr(t, P), A, P), C, P), P);
  Fe E = F::add(F::dbl(A, P), A, P);
  Fe E2 = FX::sqr(E, P);
  Jac r;
  r.X = F::sub(E2, F::dbl(D, P), P);
  Fe C8 = F::dbl(F::dbl(F::dbl(C, P), P), P);
  r.Y = F::sub(FX::mul(E, F::sub(D, r.X, P), P), C8, P);
  r.Z = F::dbl(FX::mul(p.Y, p.Z, P), P);
  return r;
}

// add-2007-bl with the exceptional cases handled: 11M + 5S
MP_NOINLINE Jac jac_add(Jac p, Jac q, const Modulus& P) {
  using namespace F;
  if (jac_is_inf(p)) return q;
  if (jac_is_inf(q)) return p;
  Fe Z1Z1 = F::sqr(p.Z, P), Z2Z2 = F::sqr(q.Z, P);
  Fe U1 = F::mul(p.X, Z2Z2, P), U2 = F::mul(q.X, Z1Z1, P);
  Fe S1 = F::mul(F::mul(p.Y, q.Z, P), Z2Z2, P), S2 = F::mul(F::mul(q.Y, p.Z, P), Z1Z1, P);
  Fe H = F::sub(U2, U1, P), rr = F::sub(S2, S1, P);
  if (is_zero(H)) {
    if (is_zero(rr)) return jac_dbl(p, P);
    return jac_infinity(P);
  }
  Fe I = F::sqr(F::dbl(H, P), P), J = F::mul(H, I, P), r2 = F::dbl(rr, P), V = F::mul(U1, I, P);
  Jac r;
  r.X = F::sub(F::sub(F::sqr(r2, P), J, P), F::dbl(V, P), P);
  r.Y = F::sub(F::mul(r2, F::sub(V, r.X, P), P), F::dbl(F::mul(S1, J, P), P), P);
  Fe zz = F::add(p.Z, q.Z, P);
  r.Z = F::mul(F::sub(F::sub(F::sqr(zz, P), Z1Z1, P), Z2Z2, P), H, P);
  return r;
}

// mixed addition (q affine): madd-2007-bl, 7M + 4S
MP_NOINLINE Jac jac_madd(Jac p, Aff q, const Modulus& P) {
  using namespace F;
  if (q.inf) return p;
  if (jac_is_inf(p)) return jac_from_aff(q, P);
  Fe Z1Z1 = FX::sqr(p.Z, P);
  Fe U2 = FX::mul(q.x, Z1Z1, P), S2 = FX::mul(FX::mul(q.y, p.Z, P), Z1Z1, P);
  Fe H = F::sub(U2, p.X, P), rr = F::sub(S2, p.Y, P);
  if (is_zero(H)) {
    if (is_zero(rr)) return jac_dbl(p, P);
    return jac_infinity(P);
  }
  Fe HH = FX::sqr(H, P), I = F::dbl(F::dbl(HH, P), P), J = FX::mul(H, I, P), r2 = F::dbl(rr, P), V = FX::mul(p.X, I, P);
  Jac r;
  r.X = F::sub(F::sub(FX::sqr(r2, P), J, P), F::dbl(V, P), P);
  r.Y = F::sub(FX::mul(r2, F::sub(V, r.X, P), P), F::dbl(FX::mul(p.Y, J, P), P), P);
  r.Z = F::sub(F::sub(FX::sqr(F::add(p.Z, H, P), P), Z1Z1, P), HH, P);
  return r;
}

// [k]p for a small scalar k < 4^nd (top base-4 digit non-zero), fixed 2-bit windows, with every
// table entry affine.  Neither the doubling nor the addition formulas of a = 0 short Weierstrass
// curves use b, so a Jacobian point (X : Y : Z) of y^2 = x^3 + b can be read as the *affine* point
// (X, Y) of the isomorphic curve y^2 = x^3 + b Z^6; results computed there map back by multiplying
// their Z by the scale.  Rescaling twice (by the Z of 2P, then by the Z of 3P) leaves P, 2P, 3P all
// affine on one curve, so the window additions are mixed additions (7M + 4S instead of 11M + 5S).
// The group element is the same as with the generic ladder: only the representative changes.
// `digit(s)` returns base-4 digit s of the scalar (s = 0 least significant), nd digits in all.
template <class DigitFn>
MP_DEV Jac small_mul_iso(const Jac& p, DigitFn digit, uint32_t nd, const Modulus& P) {
  if (jac_is_inf(p)) return p;
  Fe x1 = p.X, y1 = p.Y;
  // 2P on the first curve: mdbl-2007-bl (Z1 = 1, a = 0), 1M + 5S
  Fe XX = F::sqr(x1, P), YY = F::sqr(y1, P), YYYY = F::sqr(YY, P);
  Fe S = F::dbl(F::sub(F::sub(F::sqr(F::add(x1, YY, P), P), XX, P), YYYY, P), P);
  Fe M = F::add(F::dbl(XX, P), XX, P);
  Fe X2 = F::sub(F::sqr(M, P), F::dbl(S, P), P);
  Fe Y2 = F::sub(F::mul(M, F::sub(S, X2, P), P), F::dbl(F::dbl(F::dbl(YYYY, P), P), P), P);
  Fe Z2 = F::dbl(y1, P);
  // second curve: scale by Z2, 2P becomes affine
  Fe Z2s = F::sqr(Z2, P), Z2c = F::mul(Z2s, Z2, P);
  x1 = F::mul(x1, Z2s, P);
  y1 = F::mul(y1, Z2c, P);
  // 3P = P + 2P, both affine: mmadd-2007-bl, 4M + 2S (P != +-2P in a group of prime order)
  Fe H = F::sub(X2, x1, P), HH = F::sqr(H, P), I = F::dbl(F::dbl(HH, P), P), J = F::mul(H, I, P);
  Fe r = F::dbl(F::sub(Y2, y1, P), P), V = F::mul(x1, I, P);
  Fe X3 = F::sub(F::sub(F::sqr(r, P), J, P), F::dbl(V, P), P);
  Fe Y3 = F::sub(F::mul(r, F::sub(V, X3, P), P), F::dbl(F::mul(y1, J, P), P), P);
  Fe Z3 = F::dbl(H, P);
  // third curve: scale by Z3, all three affine
  Fe Z3s = F::sqr(Z3, P), Z3c = F::mul(Z3s, Z3, P);
  Aff t1, t2, t3;
  t1.x = F::mul(x1, Z3s, P);
  t1.y = F::mul(y1, Z3c, P);
  t2.x = F::mul(X2, Z3s, P);
  t2.y = F::mul(Y2, Z3c, P);
  t3.x = X3;
  t3.y = Y3;
  t1.inf = t2.inf = t3.inf = 0;
  const Fe scale = F::mul(F::mul(p.Z, Z2, P), Z3, P);
  Jac acc = jac_infinity(P);
#pragma unroll 1
  for (int s = (int)nd - 1; s >= 0; --s) {
    if (s != (int)nd - 1) {
      acc = jac_dbl(acc, P);
      acc = jac_dbl(acc, P);
    }
    uint32_t d = digit(s);
    Aff q = (d == 3) ? t3 : ((d == 2) ? t2 : t1);
    if (d) acc = jac_madd(acc, q, P);
  }
  acc.Z = F::mul(acc.Z, scale, P);
  return acc;
}

MP_NOINLINE Aff jac_to_aff(const Jac& p, const Modulus& P) {
  using namespace F;
  Aff a;
  if (jac_is_inf(p)) {
    a.x = fe_zero();
    a.y = fe_zero();
    a.inf = 1;
    return a;
  }
  Fe zi = F::inv(p.Z, P), zi2 = F::sqr(zi, P);
  a.x = F::mul(p.X, zi2, P);
  a.y = F::mul(p.Y, F::mul(zi2, zi, P), P);
  a.inf = 0;
  return a;
}

// ---- SEC1 compressed encoding (secp256k1.rs:133-152) ---------------------------------
// 33 bytes 02/03 || x (big-endian).  The identity is 33 zero bytes (k256's encoding).
MP_NOINLINE void encode(uint8_t* out, const Aff& a, const Modulus& P) {
  if (a.inf) {
    for (int i = 0; i < 33; ++i) out[i] = 0;
    return;
  }
  Fe x = F::from_mont(a.x, P), y = F::from_mont(a.y, P);
  out[0] = 2 + (y.v[0] & 1u);
  for (int i = 0; i < 8; ++i) {
    uint32_t w = x.v[7 - i];
    out[1 + 4 * i] = (uint8_t)(w >> 24);
    out[2 + 4 * i] = (uint8_t)(w >> 16);
    out[3 + 4 * i] = (uint8_t)(w >> 8);
    out[4 + 4 * i] = (uint8_t)w;
  }
}
// returns false for an invalid encoding (reference: bytes_to_element -> None)
MP_NOINLINE bool decode(Aff& a, const uint8_t* in, const Consts& C) {
  using namespace F;
  const Modulus& P = C.P;
  bool all_zero = true;
  for (int i = 0; i < 33; ++i) all_zero = all_zero && in[i] == 0;
  a.inf = 0;
  a.x = fe_zero();
  a.y = fe_zero();
  if (all_zero) {
    a.inf = 1;
    return true;
  }
  if (in[0] != 2 && in[0] != 3) return false;
  Fe x;
  for (int i = 0; i < 8; ++i)
    x.v[7 - i] = (uint32_t)in[1 + 4 * i] << 24 | (uint32_t)in[2 + 4 * i] << 16 | (uint32_t)in[3 + 4 * i] << 8 |
                 in[4 + 4 * i];
  // x < p
  Fe t;
  t.v[0] = simt::sub_cc(x.v[0], P.m[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) t.v[i] = simt::subc_cc(x.v[i], P.m[i]);
  if (simt::subc(0, 0) == 0) return false;
  Fe xm = F::to_mont(x, P);
  Fe y2 = F::add(F::mul(F::sqr(xm, P), xm, P), load(C.b7), P);
  Fe y = F::sqrt_candidate(y2, P);
  if (!eq(F::sqr(y, P), y2)) return false;
  Fe yn = F::from_mont(y, P);
  if ((yn.v[0] & 1u) != (uint32_t)(in[0] & 1)) y = F::neg(y, P);
  a.x = xm;
  a.y = y;
  return true;
}

// ---- curve policy for ec_kernels.cuh ---------------------------------------------------------
struct SecpCurve {
  using Consts = secp::Consts;
  using Point = Jac;
  using Affine = Aff;
  static constexpr int EB = 33;
  MP_DEV static Point small_mul(const Point& p, uint32_t k, uint32_t nd, const Consts& C) {
    return small_mul_iso(p, [k](int s) { return (k >> (2 * s)) & 3u; }, nd, C.P);
  }
  // [e]p for a full-width scalar (8 little-endian limbs) through the same 2-bit windows with the affine
  // table P, 2P, 3P: 256 doublings + ~96 mixed additions, the table in registers (no per-thread point
  // array in local memory as a 4-bit window table would need)
  MP_DEV static Point scalar_mul_wide(const Point& p, const uint32_t* e, const Consts& C) {
    return small_mul_iso(p, [e](int s) { return (e[s >> 4] >> ((s & 15) * 2)) & 3u; }, 128, C.P);
  }
  MP_DEV static Point infinity(const Consts& C) { return jac_infinity(C.P); }
  MP_DEV static Point from_aff(const Affine& a, const Consts& C) { return jac_from_aff(a, C.P); }
  MP_DEV static Point dbl(const Point& p, const Consts& C) { return jac_dbl(p, C.P); }
  MP_DEV static Point add(const Point& p, const Point& q, const Consts& C) { return jac_add(p, q, C.P); }
  MP_DEV static Point madd(const Point& p, const Affine& q, const Consts& C) { return jac_madd(p, q, C.P); }
  MP_DEV static Affine to_affine(const Point& p, const Consts& C) { return jac_to_aff(p, C.P); }
  MP_DEV static bool decode(Affine& a, const uint8_t* in, const Consts& C) { return secp::decode(a, in, C); }
  MP_DEV static void encode(uint8_t* out, const Point& p, const Consts& C) {
    secp::encode(out, jac_to_aff(p, C.P), C.P);
  }
};

}  // namespace secp
