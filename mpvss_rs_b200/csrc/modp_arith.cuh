// Cooperative 2048-bit Montgomery arithmetic for ModpGroup (replaces
// num-bigint's BigInt::modpow / `*` / `%` behind groups/modp.rs:122-132 of the
// reference).
//
// Layout: one 2048-bit value = 64 little-endian u32 limbs, spread over a *group*
// of TPI lanes of one warp (TPI in {4, 8, 16}); lane k of the group keeps limbs
// [k*L, k*L+L), L = 64/TPI, in registers.  A warp holds 32/TPI independent values.
//
// mont_mul is a fused operand-scanning Montgomery product, one 32-bit digit of b
// per iteration (64 iterations):
//     W <- (W + a_k * b_j + q_k * m_j) / 2^32        (per lane, on its own window)
// Each lane's window W is kept as two interleaved accumulators (`even`/`odd`
// 64-bit aligned column pairs) so that every 32x32->64 multiply-accumulate is one
// IMAD.WIDE.U32 with the carry chained through a predicate (mad.lo.cc/madc.hi.cc).
// Per iteration the group exchanges exactly two words by warp shuffle: the
// Montgomery digit m_j (from lane 0 of the group) and the limb each lane shifts
// out at the bottom of its window (to lane k-1).  b_j is read from shared memory
// (broadcast LDS.128 every 4 iterations).  Windows carry at most ~2 bits above
// their L limbs (bounds in DESIGN.md §modp), so no carry crosses lanes until the
// single carry-resolve at the end (ballot + add trick).
//
// Values are kept in [0, 2^2048) ("almost Montgomery"): the final conditional
// subtraction only fires on overflow of 2^2048; canonical reduction below q
// happens once, when results leave the kernel.
//
// Squarings run through the same fused loop by default (8192 MACs issued for the 6240 a symmetric
// product needs).  Two dedicated squarings exist and are bit-exact -- mont_sqr below (symmetric rows,
// 13 MACs per row) and mont_sqr_split in modp_sqr.cuh (block products + reduction-only loop) -- but at
// the benchmark's 1.74 warps per scheduler the launch is bound by dependency latency per warp, so the
// instruction count of a product decides and both measure slower.
#pragma once
#include "simt.h"

namespace modp {

template <int TPI>
struct Cfg {
  static_assert(TPI == 4 || TPI == 8 || TPI == 16, "TPI must be 4, 8 or 16");
  static constexpr int L = 64 / TPI;
};

// Per-lane slice of the modulus constants.
template <int L>
struct Mod {
  uint32_t q[L];   // limbs [k*L, k*L+L) of q
  uint32_t nq[L];  // same slice of 2^2048 - q
  uint32_t np;     // -q^{-1} mod 2^32
  const uint32_t* qh;  // 64 limbs of (q + 1) / 2 in global memory (mont_sqr only)
};

// Identity of a lane inside its group.
struct Lane {
  int k;          // lane index inside the group, 0..TPI-1
  int lane0;      // warp lane of the group's lane 0
  int shift;      // == lane0 (bit position of the group in a ballot)
  uint32_t gmask; // (1 << TPI) - 1
};

template <int TPI>
MP_DEV Lane make_lane() {
  Lane ln;
  int wl = (int)simt::lane_id();
  ln.k = wl & (TPI - 1);
  ln.lane0 = wl - ln.k;
  ln.shift = ln.lane0;
  ln.gmask = (TPI == 32) ? 0xffffffffu : ((1u << TPI) - 1u);
  return ln;
}

// Resolve a carry (or borrow) chain across the lanes of a group.
//   gen  : this lane produced a carry out of its L limbs
//   prop : this lane's limbs are all ones (all zeros for borrows)
// returns carry-in for this lane; *cout = carry out of the whole group.
template <int TPI>
MP_DEV uint32_t resolve(const Lane& ln, bool gen, bool prop, uint32_t* cout) {
  uint32_t G = (simt::ballot(gen) >> ln.shift) & ln.gmask;
  uint32_t P = (simt::ballot(prop) >> ln.shift) & ln.gmask;
  uint32_t C = ((G | P) + G) ^ P;  // bit i = carry into lane i, bit TPI = carry out
  *cout = (C >> TPI) & 1u;
  return (C >> ln.k) & 1u;
}

template <int L>
MP_DEV bool all_ones(const uint32_t (&r)[L]) {
  uint32_t x = r[0];
#pragma unroll
  for (int i = 1; i < L; ++i) x &= r[i];
  return x == 0xffffffffu;
}

// r += addend (per-lane slices of a 2048-bit value), carries resolved across the
// group; returns the carry out of bit 2048.
template <int TPI>
MP_DEV uint32_t add_resolve(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&x)[Cfg<TPI>::L], uint32_t mask,
                            const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  r[0] = simt::add_cc(r[0], x[0] & mask);
#pragma unroll
  for (int i = 1; i < L; ++i) r[i] = simt::addc_cc(r[i], x[i] & mask);
  uint32_t c = simt::addc(0, 0);
  uint32_t cout;
  uint32_t cin = resolve<TPI>(ln, c != 0, all_ones<L>(r), &cout);
  r[0] = simt::add_cc(r[0], cin);
#pragma unroll
  for (int i = 1; i < L; ++i) r[i] = simt::addc_cc(r[i], 0);
  return cout;
}

// One digit of the fused Montgomery product.  P is the accumulator whose column
// pairs are aligned with the current lowest limb ("even" role); S holds the other
// accumulator ("odd" role; for FIRST == false it arrives as the previous
// iteration's even accumulator and is shifted down one column pair in place).
template <int TPI, bool FIRST>
MP_DEV void mm_digit(uint32_t (&P)[Cfg<TPI>::L + 2], uint32_t (&S)[Cfg<TPI>::L + 2],
                     const uint32_t (&a)[Cfg<TPI>::L], uint32_t b, const Mod<Cfg<TPI>::L>& M, const Lane& ln,
                     uint32_t& in) {
  constexpr int L = Cfg<TPI>::L;
  if (FIRST) {
#pragma unroll
    for (int i = 0; i < L; i += 2) {
      P[i] = simt::mul_lo(a[i], b);
      P[i + 1] = simt::mul_hi(a[i], b);
      S[i] = simt::mul_lo(a[i + 1], b);
      S[i + 1] = simt::mul_hi(a[i + 1], b);
    }
    P[L] = P[L + 1] = 0;
    S[L] = S[L + 1] = 0;
  } else {
    // limb shifted in from lane k+1 lands on (new) column L-1 = S[L] before the shift
    S[L] = simt::add_cc(S[L], in);
    S[L + 1] = simt::addc(S[L + 1], 0);
    // stray high half of the dropped pair -> column 0; its carry feeds the odd chain
    P[0] = simt::add_cc(P[0], S[1]);
#pragma unroll
    for (int x = 0; x < L; x += 2) {
      S[x] = simt::madc_lo_cc(a[x + 1], b, S[x + 2]);
      S[x + 1] = simt::madc_hi_cc(a[x + 1], b, S[x + 3]);
    }
    S[L] = simt::addc(0, 0);
    S[L + 1] = 0;
    P[0] = simt::mad_lo_cc(a[0], b, P[0]);
    P[1] = simt::madc_hi_cc(a[0], b, P[1]);
#pragma unroll
    for (int i = 2; i < L; i += 2) {
      P[i] = simt::madc_lo_cc(a[i], b, P[i]);
      P[i + 1] = simt::madc_hi_cc(a[i], b, P[i + 1]);
    }
    P[L] = simt::addc(P[L], 0);
  }
  // Montgomery digit from the lowest limb of the whole value (group lane 0)
  uint32_t m = simt::shfl(simt::mul_lo(P[0], M.np), ln.lane0);
  S[0] = simt::mad_lo_cc(M.q[1], m, S[0]);
  S[1] = simt::madc_hi_cc(M.q[1], m, S[1]);
#pragma unroll
  for (int i = 3; i < L; i += 2) {
    S[i - 1] = simt::madc_lo_cc(M.q[i], m, S[i - 1]);
    S[i] = simt::madc_hi_cc(M.q[i], m, S[i]);
  }
  S[L] = simt::addc(S[L], 0);
  P[0] = simt::mad_lo_cc(M.q[0], m, P[0]);
  P[1] = simt::madc_hi_cc(M.q[0], m, P[1]);
#pragma unroll
  for (int i = 2; i < L; i += 2) {
    P[i] = simt::madc_lo_cc(M.q[i], m, P[i]);
    P[i + 1] = simt::madc_hi_cc(M.q[i], m, P[i + 1]);
  }
  P[L] = simt::addc(P[L], 0);
  // P[0] is now the limb that leaves this lane's window (zero on group lane 0)
  // The top lane of a group reads lane 0 of the next group (lane 31 wraps to lane 0), whose P[0] is
  // exactly zero here (P[0] + q[0]*m = 0 mod 2^32 by the choice of m): no select needed.
  in = simt::shfl(P[0], ((int)simt::lane_id() + 1) & 31);
}

// Fold the two accumulators of a finished digit loop (last call had P = A1, S = A0), hand the
// window overflows to the next lane, resolve carries and apply the conditional subtraction.
template <int TPI, bool DOUBLE = false>
MP_DEV void mm_finish(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&A0)[Cfg<TPI>::L + 2],
                      const uint32_t (&A1)[Cfg<TPI>::L + 2], uint32_t in, const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  // value = S + (P >> 32) + in * 2^(32(L-1))
  r[0] = simt::add_cc(A0[0], A1[1]);
#pragma unroll
  for (int i = 1; i < L; ++i) r[i] = simt::addc_cc(A0[i], A1[i + 1]);
  uint32_t ov = simt::addc(A0[L], A1[L + 1]);
  r[L - 1] = simt::add_cc(r[L - 1], in);
  ov = simt::addc(ov, 0);
  // hand the (<= 2 bit) overflow of each window to the next lane, resolve carries
  uint32_t ovin = simt::shfl(ov, (int)simt::lane_id() - 1);
  uint32_t ovtop = simt::shfl(ov, ln.lane0 + TPI - 1);
  if (ln.k == 0) ovin = 0;
  r[0] = simt::add_cc(r[0], ovin);
#pragma unroll
  for (int i = 1; i < L; ++i) r[i] = simt::addc_cc(r[i], 0);
  uint32_t c = simt::addc(0, 0);
  uint32_t cout;
  uint32_t cin = resolve<TPI>(ln, c != 0, all_ones<L>(r), &cout);
  r[0] = simt::add_cc(r[0], cin);
#pragma unroll
  for (int i = 1; i < L; ++i) r[i] = simt::addc_cc(r[i], 0);
  // value >= 2^2048  ->  subtract q once (add 2^2048 - q, drop the carry)
  uint32_t over = ovtop + cout;
  if (!DOUBLE) {
    (void)add_resolve<TPI>(r, M.nq, 0u - over, ln);
  } else {
    // mont_sqr: the loop produced V with 2V = a^2 / 2^2048 (mod q), V <= 2^2047 + q.  Double across
    // the lanes (the bit leaving a lane enters the next one), then take q off at most twice:
    // 2V <= 2^2048 + 2q, and after the first subtraction a value in [2^2048, 2^2048 + nq) is still
    // possible, which the second round catches through the carry of the first.
    uint32_t top = r[L - 1] >> 31;
    uint32_t bit_in = simt::shfl(top, ((int)simt::lane_id() - 1) & 31);
    uint32_t top_all = simt::shfl(top, ln.lane0 + TPI - 1);
    if (ln.k == 0) bit_in = 0;
#pragma unroll
    for (int i = L - 1; i >= 1; --i) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
    r[0] = (r[0] << 1) | bit_in;
    uint32_t over2 = 2u * over + top_all;                 // 0, 1 or 2
    uint32_t m1 = over2 ? 0xffffffffu : 0u;
    uint32_t c1 = add_resolve<TPI>(r, M.nq, m1, ln);
    over2 = over2 - (m1 & 1u) + c1;
    (void)add_resolve<TPI>(r, M.nq, over2 ? 0xffffffffu : 0u, ln);
  }
}

// r = a * b * 2^-2048 mod q, result in [0, 2^2048).  `bs` points at the 64 limbs of
// b (shared memory, visible to the whole group).  r may alias a.
template <int TPI>
MP_DEV void mont_mul_fused(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&a)[Cfg<TPI>::L], const uint32_t* bs,
                           const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  uint32_t A0[L + 2], A1[L + 2];
  uint32_t in = 0;
  // Each accumulator array moves down one 64-bit pair every second digit, so the register
  // assignment repeats after PER = L + 2 digits: a loop body of exactly PER digits closes on
  // itself without the register moves a 4-digit body needs at its back edge (13 % of the
  // instructions).  64 mod PER digits are peeled in front (4, 4, 10 for L = 4, 8, 16).
  constexpr int PER = L + 2, PEEL = 64 % PER;
  static_assert(PEEL >= 2 && PEEL % 2 == 0 && PER % 2 == 0, "digit loop layout");
  const uint2* b2 = reinterpret_cast<const uint2*>(bs);
  {
    uint2 bw = b2[0];
    mm_digit<TPI, true>(A0, A1, a, bw.x, M, ln, in);
    mm_digit<TPI, false>(A1, A0, a, bw.y, M, ln, in);
  }
#pragma unroll
  for (int d = 2; d < PEEL; d += 2) {
    uint2 bw = b2[d / 2];
    mm_digit<TPI, false>(A0, A1, a, bw.x, M, ln, in);
    mm_digit<TPI, false>(A1, A0, a, bw.y, M, ln, in);
  }
  // Two periods per trip: ptxas still leaves a few register moves at the back edge of the period-matched
  // body (10 per trip in the chain kernel); unrolling by two halves their share (259 -> 245 instructions per
  // 10 digits at TPI = 8).
#ifndef MPVSS_MODP_UNROLL
#define MPVSS_MODP_UNROLL 2
#endif
  constexpr int UNROLL = MPVSS_MODP_UNROLL;
#pragma unroll UNROLL
  for (int it = 0; it < (64 - PEEL) / PER; ++it) {
    const uint2* p = b2 + (PEEL + it * PER) / 2;
#pragma unroll
    for (int d = 0; d < PER; d += 2) {
      uint2 bw = p[d / 2];
      mm_digit<TPI, false>(A0, A1, a, bw.x, M, ln, in);
      mm_digit<TPI, false>(A1, A0, a, bw.y, M, ln, in);
    }
  }
  mm_finish<TPI>(r, A0, A1, in, M, ln);
}

// ---- split accumulators ----------------------------------------------------------------------
// Same product with the a*b rows and the q*m rows kept in separate windows (AP/AS and QP/QS), so
// that the four carry chains of a row are independent of each other and the a*b chains of row j+1
// do not wait for the Montgomery digit of row j: the only coupling is the low word, where
//     m_j = (AP[0] + QP[0] + stray + c) * np,   exported word = AP[0] + QP[0] + c  (carry c kept
// for the next row's column).  Costs 6 more carry instructions per row than the fused form; meant
// for low occupancy, where dependency latency and not issue bandwidth is the limit.
template <int TPI>
MP_DEV void mm_digit_il(uint32_t (&AP)[Cfg<TPI>::L + 2], uint32_t (&AS)[Cfg<TPI>::L + 2],
                        uint32_t (&QP)[Cfg<TPI>::L + 2], uint32_t (&QS)[Cfg<TPI>::L + 2],
                        const uint32_t (&a)[Cfg<TPI>::L], uint32_t b, const Mod<Cfg<TPI>::L>& M, const Lane& ln,
                        uint32_t& in, uint32_t& cprev) {
  constexpr int L = Cfg<TPI>::L;
  // a*b half: identical to mm_digit on (AP, AS)
  AS[L] = simt::add_cc(AS[L], in);
  AS[L + 1] = simt::addc(AS[L + 1], 0);
  AP[0] = simt::add_cc(AP[0], AS[1]);
#pragma unroll
  for (int x = 0; x < L; x += 2) {
    AS[x] = simt::madc_lo_cc(a[x + 1], b, AS[x + 2]);
    AS[x + 1] = simt::madc_hi_cc(a[x + 1], b, AS[x + 3]);
  }
  AS[L] = simt::addc(0, 0);
  AS[L + 1] = 0;
  AP[0] = simt::mad_lo_cc(a[0], b, AP[0]);
  AP[1] = simt::madc_hi_cc(a[0], b, AP[1]);
#pragma unroll
  for (int i = 2; i < L; i += 2) {
    AP[i] = simt::madc_lo_cc(a[i], b, AP[i]);
    AP[i + 1] = simt::madc_hi_cc(a[i], b, AP[i + 1]);
  }
  AP[L] = simt::addc(AP[L], 0);
  // Montgomery digit from the low word of the whole value (plain adds: no carry flag across the shuffle)
  uint32_t m = simt::shfl(simt::mul_lo(AP[0] + QP[0] + QS[1] + cprev, M.np), ln.lane0);
  // q*m half on (QP, QS); the odd window is shifted down inside its chain
  QP[0] = simt::add_cc(QP[0], QS[1]);
  QS[0] = simt::madc_lo_cc(M.q[1], m, QS[2]);
  QS[1] = simt::madc_hi_cc(M.q[1], m, QS[3]);
#pragma unroll
  for (int i = 3; i < L; i += 2) {
    QS[i - 1] = simt::madc_lo_cc(M.q[i], m, QS[i + 1]);
    QS[i] = simt::madc_hi_cc(M.q[i], m, QS[i + 2]);
  }
  QS[L] = simt::addc(0, 0);
  QS[L + 1] = 0;
  QP[0] = simt::mad_lo_cc(M.q[0], m, QP[0]);
  QP[1] = simt::madc_hi_cc(M.q[0], m, QP[1]);
#pragma unroll
  for (int i = 2; i < L; i += 2) {
    QP[i] = simt::madc_lo_cc(M.q[i], m, QP[i]);
    QP[i + 1] = simt::madc_hi_cc(M.q[i], m, QP[i + 1]);
  }
  QP[L] = simt::addc(QP[L], 0);
  // word leaving the window (zero on group lane 0) and the carry of its column
  uint32_t e = simt::add_cc(AP[0], QP[0]);
  uint32_t c = simt::addc(0, 0);
  e = simt::add_cc(e, cprev);
  cprev = simt::addc(c, 0);
  in = simt::shfl(e, ((int)simt::lane_id() + 1) & 31);
}

template <int TPI>
MP_DEV void mont_mul_il(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&a)[Cfg<TPI>::L], const uint32_t* bs,
                        const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  uint32_t A0[L + 2], A1[L + 2], Q0[L + 2], Q1[L + 2];
#pragma unroll
  for (int i = 0; i < L + 2; ++i) A0[i] = A1[i] = Q0[i] = Q1[i] = 0;
  uint32_t in = 0, cprev = 0;
  constexpr int PER = L + 2, PEEL = 64 % PER;
  const uint2* b2 = reinterpret_cast<const uint2*>(bs);
#pragma unroll
  for (int d = 0; d < PEEL; d += 2) {
    uint2 bw = b2[d / 2];
    mm_digit_il<TPI>(A0, A1, Q0, Q1, a, bw.x, M, ln, in, cprev);
    mm_digit_il<TPI>(A1, A0, Q1, Q0, a, bw.y, M, ln, in, cprev);
  }
#pragma unroll 1
  for (int it = 0; it < (64 - PEEL) / PER; ++it) {
    const uint2* p = b2 + (PEEL + it * PER) / 2;
#pragma unroll
    for (int d = 0; d < PER; d += 2) {
      uint2 bw = p[d / 2];
      mm_digit_il<TPI>(A0, A1, Q0, Q1, a, bw.x, M, ln, in, cprev);
      mm_digit_il<TPI>(A1, A0, Q1, Q0, a, bw.y, M, ln, in, cprev);
    }
  }
  // merge the two windows (plus the pending carry of the lowest column) and finish as usual;
  // after an even number of rows A0/Q0 hold the odd role: their word 0 is the lowest column
  A0[0] = simt::add_cc(A0[0], cprev);
#pragma unroll
  for (int i = 1; i < L + 1; ++i) A0[i] = simt::addc_cc(A0[i], 0);
  A0[L + 1] = simt::addc(A0[L + 1], 0);
  A0[0] = simt::add_cc(A0[0], Q0[0]);
#pragma unroll
  for (int i = 1; i < L + 1; ++i) A0[i] = simt::addc_cc(A0[i], Q0[i]);
  A0[L + 1] = simt::addc(A0[L + 1], Q0[L + 1]);
  // word 0 of the even-role arrays is the exported (dropped) word: only words 1.. count
  A1[1] = simt::add_cc(A1[1], Q1[1]);
#pragma unroll
  for (int i = 2; i < L + 1; ++i) A1[i] = simt::addc_cc(A1[i], Q1[i]);
  A1[L + 1] = simt::addc(A1[L + 1], Q1[L + 1]);
  mm_finish<TPI>(r, A0, A1, in, M, ln);
}

template <int TPI>
MP_DEV void mont_mul(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&a)[Cfg<TPI>::L], const uint32_t* bs,
                     const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
#ifdef MPVSS_MODP_SPLIT_ACC
  mont_mul_il<TPI>(r, a, bs, M, ln);
#else
  mont_mul_fused<TPI>(r, a, bs, M, ln);
#endif
}

// ---- dedicated squaring ----------------------------------------------------------------------
// r = a * a * 2^-2048 mod q through the same fused digit loop, issuing L/2 + 1 instead of L
// multiply-accumulates of a per row (5 of 8 at TPI = 8; 13 instead of 16 wide MACs per row with
// the reduction).  Every unordered pair {i, j}, i != j, of limbs is multiplied exactly once, in the
// row of the limb the other one follows within half a period: row j (digit a_j) takes limb i when
//     d = (i - j) mod L  is in [1, L/2 - 1]          (every lane, every row)
//     d = L/2 or d = 0 (i != j)  and  i > j          (one MAC slot each, digit masked per lane)
// so each lane runs the same L/2 + 1 MAC slots in every row and no lane idles.  Off-diagonal
// products would have to be doubled; instead the loop accumulates HALF the square,
//     T'' = (a^2 + p*q) / 2,   p = a mod 2   (a^2 + p*q is even and congruent to a^2),
// i.e. off-diagonal products once, the squares a_i^2 halved, and p*(q+1)/2 as the start value.
// The halved squares of a lane's own limbs form one 2L-word number h = (sum a_i^2 B^(2i)) >> 1
// that is added to the window in two pieces (rows L*k and L*k + L - 2 of the lane's own block k);
// the bit shifted out at the bottom belongs to column 2Lk - 1, which lane k-1 holds at offset L-1.
// The fused reduction then yields V with 2V = a^2 * 2^-2048 (mod q); mm_finish<DOUBLE> doubles it.
struct SqMasks {
  uint32_t gt, ge, eq, prev;  // all-ones if this lane's index is >, >=, == the block's / == block - 1
};

template <int TPI, int RP>
MP_DEV void sq_digit(uint32_t (&P)[Cfg<TPI>::L + 2], uint32_t (&S)[Cfg<TPI>::L + 2],
                     const uint32_t (&a)[Cfg<TPI>::L], const uint32_t (&h)[2 * Cfg<TPI>::L], uint32_t pbn31,
                     uint32_t b, const SqMasks& mk, const Mod<Cfg<TPI>::L>& M, const Lane& ln, uint32_t& in) {
  constexpr int L = Cfg<TPI>::L;
  constexpr int H = L / 2;
  // halved squares of the lane's own limbs (only the lane that owns this block adds anything)
  if (RP == 0) {
    P[0] = simt::add_cc(P[0], h[0] & mk.eq);
#pragma unroll
    for (int i = 1; i <= L + 1; ++i) {
      uint32_t v = h[i] & mk.eq;
      if (i == L - 1) v |= pbn31 & mk.prev;
      P[i] = simt::addc_cc(P[i], v);
    }
  } else if (RP == L - 2) {
    P[4] = simt::add_cc(P[4], h[L + 2] & mk.eq);
#pragma unroll
    for (int i = 5; i <= L + 1; ++i) P[i] = simt::addc_cc(P[i], h[L - 2 + i] & mk.eq);
  }
  const uint32_t b_tie = b & (RP < H ? mk.ge : mk.gt);
  const uint32_t b_gt = b & mk.gt;
  // limb shifted in from lane k+1 lands on (new) column L-1 = S[L] before the shift
  S[L] = simt::add_cc(S[L], in);
  S[L + 1] = simt::addc(S[L + 1], 0);
  // stray high half of the dropped pair -> column 0; its carry feeds the odd chain
  P[0] = simt::add_cc(P[0], S[1]);
#pragma unroll
  for (int x = 0; x < L; x += 2) {
    const int sl = (x + 1 - RP + L) % L;  // slot of limb x + 1 in this row
    if (sl <= H) {
      const uint32_t d = sl == 0 ? b_gt : (sl == H ? b_tie : b);
      S[x] = simt::madc_lo_cc(a[x + 1], d, S[x + 2]);
      S[x + 1] = simt::madc_hi_cc(a[x + 1], d, S[x + 3]);
    } else {
      S[x] = simt::addc_cc(S[x + 2], 0);
      S[x + 1] = simt::addc_cc(S[x + 3], 0);
    }
  }
  S[L] = simt::addc(0, 0);
  S[L + 1] = 0;
  bool started = false;
#pragma unroll
  for (int i = 0; i < L; i += 2) {
    const int sl = (i - RP + L) % L;
    if (sl <= H) {
      const uint32_t d = sl == 0 ? b_gt : (sl == H ? b_tie : b);
      P[i] = started ? simt::madc_lo_cc(a[i], d, P[i]) : simt::mad_lo_cc(a[i], d, P[i]);
      P[i + 1] = simt::madc_hi_cc(a[i], d, P[i + 1]);
      started = true;
    } else if (started) {
      P[i] = simt::addc_cc(P[i], 0);
      P[i + 1] = simt::addc_cc(P[i + 1], 0);
    }
  }
  P[L] = simt::addc_cc(P[L], 0);
  P[L + 1] = simt::addc(P[L + 1], 0);
  // reduction half of the row: identical to mm_digit, except that the top pair of P can hold a
  // full word right after a block addition, so its carry goes on into P[L + 1]
  uint32_t m = simt::shfl(simt::mul_lo(P[0], M.np), ln.lane0);
  S[0] = simt::mad_lo_cc(M.q[1], m, S[0]);
  S[1] = simt::madc_hi_cc(M.q[1], m, S[1]);
#pragma unroll
  for (int i = 3; i < L; i += 2) {
    S[i - 1] = simt::madc_lo_cc(M.q[i], m, S[i - 1]);
    S[i] = simt::madc_hi_cc(M.q[i], m, S[i]);
  }
  S[L] = simt::addc(S[L], 0);
  P[0] = simt::mad_lo_cc(M.q[0], m, P[0]);
  P[1] = simt::madc_hi_cc(M.q[0], m, P[1]);
#pragma unroll
  for (int i = 2; i < L; i += 2) {
    P[i] = simt::madc_lo_cc(M.q[i], m, P[i]);
    P[i + 1] = simt::madc_hi_cc(M.q[i], m, P[i + 1]);
  }
  P[L] = simt::addc_cc(P[L], 0);
  P[L + 1] = simt::addc(P[L + 1], 0);
  in = simt::shfl(P[0], ((int)simt::lane_id() + 1) & 31);
}

template <int TPI, int RP, bool END = (RP >= Cfg<TPI>::L)>
struct SqRows {
  static MP_DEV void run(uint32_t (&A0)[Cfg<TPI>::L + 2], uint32_t (&A1)[Cfg<TPI>::L + 2],
                         const uint32_t (&a)[Cfg<TPI>::L], const uint32_t (&h)[2 * Cfg<TPI>::L], uint32_t pbn31,
                         const uint2* bp, const SqMasks& mk, const Mod<Cfg<TPI>::L>& M, const Lane& ln,
                         uint32_t& in) {
    uint2 bw = bp[RP / 2];
    sq_digit<TPI, RP>(A0, A1, a, h, pbn31, bw.x, mk, M, ln, in);
    sq_digit<TPI, RP + 1>(A1, A0, a, h, pbn31, bw.y, mk, M, ln, in);
    SqRows<TPI, RP + 2>::run(A0, A1, a, h, pbn31, bp, mk, M, ln, in);
  }
};
template <int TPI, int RP>
struct SqRows<TPI, RP, true> {
  static MP_DEV void run(uint32_t (&)[Cfg<TPI>::L + 2], uint32_t (&)[Cfg<TPI>::L + 2], const uint32_t (&)[Cfg<TPI>::L],
                         const uint32_t (&)[2 * Cfg<TPI>::L], uint32_t, const uint2*, const SqMasks&,
                         const Mod<Cfg<TPI>::L>&, const Lane&, uint32_t&) {}
};

// r = a^2 * 2^-2048 mod q, result in [0, 2^2048).  `as` points at the 64 limbs of a in shared
// memory (the digits of the rows); r may alias a.
template <int TPI>
MP_DEV void mont_sqr(uint32_t (&r)[Cfg<TPI>::L], const uint32_t (&a)[Cfg<TPI>::L], const uint32_t* as,
                     const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  uint32_t h[2 * L];
  {
    uint32_t w[2 * L];
#pragma unroll
    for (int i = 0; i < L; ++i) {
      w[2 * i] = simt::mul_lo(a[i], a[i]);
      w[2 * i + 1] = simt::mul_hi(a[i], a[i]);
    }
#pragma unroll
    for (int i = 0; i < 2 * L - 1; ++i) h[i] = (w[i] >> 1) | (w[i + 1] << 31);
    h[2 * L - 1] = w[2 * L - 1] >> 1;
    h[0] |= 0;
    // parity bit of the lane's block: to lane k-1 (column 2Lk - 1), lane 0's selects (q+1)/2
    const uint32_t pb = w[0] & 1u;
    const uint32_t pbn = simt::shfl(pb, ((int)simt::lane_id() + 1) & 31);
    const uint32_t pb0 = simt::shfl(pb, ln.lane0);
    uint32_t A0[L + 2], A1[L + 2];
    const uint32_t m0 = 0u - pb0;
#pragma unroll
    for (int i = 0; i < L; ++i) {
      A0[i] = M.qh[ln.k * L + i] & m0;
      A1[i] = 0;
    }
    A0[L] = A0[L + 1] = A1[L] = A1[L + 1] = 0;
    uint32_t in = 0;
    const uint32_t pbn31 = pbn << 31;
    const uint2* b2 = reinterpret_cast<const uint2*>(as);
#pragma unroll 1
    for (int kp = 0; kp < TPI; ++kp) {
      SqMasks mk;
      mk.gt = ln.k > kp ? 0xffffffffu : 0u;
      mk.ge = ln.k >= kp ? 0xffffffffu : 0u;
      mk.eq = ln.k == kp ? 0xffffffffu : 0u;
      mk.prev = ln.k + 1 == kp ? 0xffffffffu : 0u;
      SqRows<TPI, 0>::run(A0, A1, a, h, pbn31, b2 + kp * (L / 2), mk, M, ln, in);
    }
    mm_finish<TPI, true>(r, A0, A1, in, M, ln);
  }
}

// Two independent Montgomery products on the same lane group, digit loops interleaved so that
// each lane always has two dependency chains in flight (the carry chains of one product hide
// the fixed latencies and shuffle round trips of the other).
template <int TPI>
MP_DEV void mont_mul2(uint32_t (&r0)[Cfg<TPI>::L], const uint32_t (&a0)[Cfg<TPI>::L], const uint32_t* b0s,
                      uint32_t (&r1)[Cfg<TPI>::L], const uint32_t (&a1)[Cfg<TPI>::L], const uint32_t* b1s,
                      const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  uint32_t A0[L + 2], A1[L + 2], B0[L + 2], B1[L + 2];
  uint32_t in0 = 0, in1 = 0;
  const uint4* p0 = reinterpret_cast<const uint4*>(b0s);
  const uint4* p1 = reinterpret_cast<const uint4*>(b1s);
  {
    uint4 u = p0[0], v = p1[0];
    mm_digit<TPI, true>(A0, A1, a0, u.x, M, ln, in0);
    mm_digit<TPI, true>(B0, B1, a1, v.x, M, ln, in1);
    mm_digit<TPI, false>(A1, A0, a0, u.y, M, ln, in0);
    mm_digit<TPI, false>(B1, B0, a1, v.y, M, ln, in1);
    mm_digit<TPI, false>(A0, A1, a0, u.z, M, ln, in0);
    mm_digit<TPI, false>(B0, B1, a1, v.z, M, ln, in1);
    mm_digit<TPI, false>(A1, A0, a0, u.w, M, ln, in0);
    mm_digit<TPI, false>(B1, B0, a1, v.w, M, ln, in1);
  }
#pragma unroll 1
  for (int j = 1; j < 16; ++j) {
    uint4 u = p0[j], v = p1[j];
    mm_digit<TPI, false>(A0, A1, a0, u.x, M, ln, in0);
    mm_digit<TPI, false>(B0, B1, a1, v.x, M, ln, in1);
    mm_digit<TPI, false>(A1, A0, a0, u.y, M, ln, in0);
    mm_digit<TPI, false>(B1, B0, a1, v.y, M, ln, in1);
    mm_digit<TPI, false>(A0, A1, a0, u.z, M, ln, in0);
    mm_digit<TPI, false>(B0, B1, a1, v.z, M, ln, in1);
    mm_digit<TPI, false>(A1, A0, a0, u.w, M, ln, in0);
    mm_digit<TPI, false>(B1, B0, a1, v.w, M, ln, in1);
  }
  mm_finish<TPI>(r0, A0, A1, in0, M, ln);
  mm_finish<TPI>(r1, B0, B1, in1, M, ln);
}

// Bring a value of [0, 2^2048) into [0, q): canonical representative.
template <int TPI>
MP_DEV void canonical(uint32_t (&r)[Cfg<TPI>::L], const Mod<Cfg<TPI>::L>& M, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  // Two conditional subtractions: one suffices when 2^2048 < 2q (the group modulus), but the subgroup
  // order g = (q-1)/2 is just below 2^2047, so a value of [2g, 2^2048) needs the second one.
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    uint32_t t[L];
#pragma unroll
    for (int i = 0; i < L; ++i) t[i] = r[i];
    uint32_t ge = add_resolve<TPI>(t, M.nq, 0xffffffffu, ln);  // carry <=> r >= q
    if (ge) {
#pragma unroll
      for (int i = 0; i < L; ++i) r[i] = t[i];
    }
  }
}

// Write a group's value (64 limbs, each lane its slice) to memory: used for global stores, possibly
// under a per-group predicate, so it contains no barrier.
template <int TPI>
MP_DEV void stage(uint32_t* dst64, const uint32_t (&v)[Cfg<TPI>::L], const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  uint4* d = reinterpret_cast<uint4*>(dst64 + ln.k * L);
#pragma unroll
  for (int i = 0; i < L; i += 4) d[i / 4] = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

// Shared-memory staging so that every lane can read all 64 limbs; must be called by the whole warp.
// The leading warp barrier orders the write after the other lanes' earlier reads of the buffer (the
// shuffles of the preceding product already serialise them in practice, but only __syncwarp is a
// memory-ordering guarantee; compute-sanitizer racecheck is clean with it).  The caller issues the
// trailing barrier (often once for several staged values) before anybody reads.
template <int TPI>
MP_DEV void stage_shared(uint32_t* dst64, const uint32_t (&v)[Cfg<TPI>::L], const Lane& ln) {
  simt::syncwarp();
  stage<TPI>(dst64, v, ln);
}

template <int TPI>
MP_DEV void load_slice(uint32_t (&v)[Cfg<TPI>::L], const uint32_t* src64, const Lane& ln) {
  constexpr int L = Cfg<TPI>::L;
  const uint4* s = reinterpret_cast<const uint4*>(src64 + ln.k * L);
#pragma unroll
  for (int i = 0; i < L; i += 4) {
    uint4 w = s[i / 4];
    v[i] = w.x;
    v[i + 1] = w.y;
    v[i + 2] = w.z;
    v[i + 3] = w.w;
  }
}

}  // namespace modp
