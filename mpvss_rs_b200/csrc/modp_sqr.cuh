// Split Montgomery squaring for TPI = 8 (lane k of a group holds the 256-bit block A_k of a):
//   1. block products, thread-local: lane k computes A_k^2 and A_k * A_l for l = k+1, k+2, k+3
//      (mod 8) and, on lanes 0..3, l = k+4 -- every unordered pair once, 36 + 4*64 MAC slots per
//      lane instead of the 512 a fused product spends on a*b -- and scatters the 512-bit results
//      into the group's shared-memory scratch, grouped by block column c = k + l;
//   2. column sums: lane j adds up the pieces of two 256-bit blocks of T = a^2 (blocks paired so
//      that every lane handles 7 cross pieces), doubles the cross sum, adds the square piece and
//      writes the block plus its carry word back;
//   3. reduction-only digit loop: the fused loop of mont_mul without the a*b half (8 MACs per row);
//      lane k starts from T[8k..8k+7], the upper half of T enters at the top lane one word per row.
// Same result as mont_mul(a, a): r = a^2 * 2^-2048 mod q in [0, 2^2048).
#pragma once
#include "fpspecial.cuh"
#include "modp_arith.cuh"

namespace modp {

// scratch layout per group, in words (all 16-byte aligned)
enum {
  SQS_CR = 0,        // 28 cross products, 16 words each, ordered by block column then by the smaller index
  SQS_SQ = 448,      // 8 squares, 16 words each
  SQS_ZERO = 576,    // 8 zero words (padding target of unused slots)
  SQS_DUMMY = 584,   // 4 x 16 words: sinks of the unused product slot on lanes 4..7
  SQS_T = 648,       // T = a^2 as 16 blocks of 8 words, carries not yet propagated between blocks
  SQS_C = 776,       // carry word of each block (belongs to word 0 of the next block)
  SQS_WORDS = 792
};

// Per-lane addresses into the scratch (word offsets); depends on the lane index only.
struct SqPlan {
  uint32_t cr[4];       // where the lane's four cross products go
  uint32_t partner[4];  // word offset of the partner block inside the staged copy of a
  uint32_t px[3], py[7];  // pieces (8 words each) summed into the lane's two blocks
  uint32_t sqx, sqy;    // square pieces of the two blocks
  uint32_t tx, ty, cx, cy;  // where the two blocks and their carry words are stored
  uint32_t cin0;        // carry word added to word 0 of the lane's start window (block k - 1)
  uint32_t c8;          // carry of block 7: top lane only
  uint32_t tin, tin_stride;  // upper half of T: top lane reads T[64 + j] after row j
  uint32_t cst, cst_stride;  // carries of blocks 8..14: top lane, one every 8 rows
};

MP_DEV uint32_t sq_cw(int c) {  // number of pairs k < l <= 7 with k + l = c
  if (c < 1 || c > 13) return 0u;
  return c <= 7 ? (uint32_t)((c + 1) / 2) : (uint32_t)((c - 1) / 2 - c + 8);
}
MP_DEV uint32_t sq_prefix(int c) {
  uint32_t p = 0;
  for (int x = 1; x < c; ++x) p += sq_cw(x);
  return p;
}
MP_DEV void sq_pieces(int b, uint32_t* out, int cap) {
  int n = 0;
  for (int sel = 0; sel < 2; ++sel) {
    const int c = b - sel;  // low halves of column b, high halves of column b - 1
    for (uint32_t s = 0; s < sq_cw(c); ++s) out[n++] = SQS_CR + (sq_prefix(c) + s) * 16 + (sel ? 8u : 0u);
  }
  for (; n < cap; ++n) out[n] = SQS_ZERO;
}

MP_DEV SqPlan make_sq_plan(int k) {
  SqPlan p;
#pragma unroll
  for (int d = 1; d <= 4; ++d) {
    const int l = (k + d) & 7;
    const int lo = k < l ? k : l, hi = k < l ? l : k, c = lo + hi;
    const int s = lo - (c > 7 ? c - 7 : 0);
    const bool valid = d < 4 || k < 4;
    p.cr[d - 1] = valid ? SQS_CR + (sq_prefix(c) + (uint32_t)s) * 16 : (uint32_t)SQS_DUMMY + 16u * (uint32_t)(k - 4);
    p.partner[d - 1] = 8u * (uint32_t)l;
  }
  const int bx = k < 4 ? k : 19 - k, by = k < 4 ? 7 - k : k + 4;
  uint32_t tmp[8];
  sq_pieces(bx, tmp, 3);
#pragma unroll
  for (int i = 0; i < 3; ++i) p.px[i] = tmp[i];
  sq_pieces(by, tmp, 7);
#pragma unroll
  for (int i = 0; i < 7; ++i) p.py[i] = tmp[i];
  p.sqx = SQS_SQ + (uint32_t)(bx >> 1) * 16 + (uint32_t)(bx & 1) * 8;
  p.sqy = SQS_SQ + (uint32_t)(by >> 1) * 16 + (uint32_t)(by & 1) * 8;
  p.tx = SQS_T + 8u * (uint32_t)bx;
  p.ty = SQS_T + 8u * (uint32_t)by;
  p.cx = SQS_C + (uint32_t)bx;
  p.cy = SQS_C + (uint32_t)by;
  p.cin0 = k ? SQS_C + (uint32_t)k - 1 : (uint32_t)SQS_ZERO;
  const bool top = k == 7;
  p.c8 = top ? SQS_C + 7u : (uint32_t)SQS_ZERO;
  p.tin = top ? SQS_T + 64u : (uint32_t)SQS_ZERO;
  p.tin_stride = top ? 1u : 0u;
  p.cst = top ? SQS_C + 8u : (uint32_t)SQS_ZERO;
  p.cst_stride = top ? 1u : 0u;
  return p;
}

MP_DEV void sq_store16(uint32_t* dst, const uint32_t (&t)[16]) {
  uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 4; ++i) d4[i] = make_uint4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
}
MP_DEV void sq_load8(uint32_t (&v)[8], const uint32_t* src) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4 u = s4[0], w = s4[1];
  v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  v[4] = w.x; v[5] = w.y; v[6] = w.z; v[7] = w.w;
}
// X (8 words + carry word) += 8 words at src
MP_DEV void sq_add8(uint32_t (&X)[9], const uint32_t* src) {
  uint32_t v[8];
  sq_load8(v, src);
  X[0] = simt::add_cc(X[0], v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) X[i] = simt::addc_cc(X[i], v[i]);
  X[8] = simt::addc(X[8], 0);
}
MP_DEV void sq_double9(uint32_t (&X)[9]) {
  X[0] = simt::add_cc(X[0], X[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) X[i] = simt::addc_cc(X[i], X[i]);
  X[8] = simt::addc(X[8], X[8]);
}
MP_DEV void sq_store_block(uint32_t* scratch, uint32_t t_off, uint32_t c_off, const uint32_t (&X)[9]) {
  uint4* d4 = reinterpret_cast<uint4*>(scratch + t_off);
  d4[0] = make_uint4(X[0], X[1], X[2], X[3]);
  d4[1] = make_uint4(X[4], X[5], X[6], X[7]);
  scratch[c_off] = X[8];
}

// One row of the reduction-only loop: W <- (W + q * m) / 2^32 with the word `in` entering at the top
// (mm_digit without its a*b half; the odd accumulator is shifted down inside the q*m chain).
template <bool CARRY>
MP_DEV void red_digit(uint32_t (&P)[10], uint32_t (&S)[10], const Mod<8>& M, const Lane& ln, uint32_t& in,
                      uint32_t cin, uint32_t tin_next) {
  constexpr int L = 8;
  S[L] = simt::add_cc(S[L], in);
  S[L + 1] = simt::addc(S[L + 1], 0);
  if (CARRY) {
    S[L] = simt::add_cc(S[L], cin);
    S[L + 1] = simt::addc(S[L + 1], 0);
  }
  // Montgomery digit first (plain add: the carry flag must not live across the shuffle) ...
  uint32_t m = simt::shfl(simt::mul_lo(P[0] + S[1], M.np), ln.lane0);
  // ... then the same add with its carry feeding the odd chain directly
  P[0] = simt::add_cc(P[0], S[1]);
  S[0] = simt::madc_lo_cc(M.q[1], m, S[2]);
  S[1] = simt::madc_hi_cc(M.q[1], m, S[3]);
#pragma unroll
  for (int i = 3; i < L; i += 2) {
    S[i - 1] = simt::madc_lo_cc(M.q[i], m, S[i + 1]);
    S[i] = simt::madc_hi_cc(M.q[i], m, S[i + 2]);
  }
  S[L] = simt::addc(0, 0);
  S[L + 1] = 0;
  P[0] = simt::mad_lo_cc(M.q[0], m, P[0]);
  P[1] = simt::madc_hi_cc(M.q[0], m, P[1]);
#pragma unroll
  for (int i = 2; i < L; i += 2) {
    P[i] = simt::madc_lo_cc(M.q[i], m, P[i]);
    P[i + 1] = simt::madc_hi_cc(M.q[i], m, P[i + 1]);
  }
  P[L] = simt::addc(P[L], 0);
  // limb leaving the window goes to lane k-1; on the top lane (which receives zero) the next word of
  // the upper half of T takes its place
  in = simt::shfl(P[0], ((int)simt::lane_id() + 1) & 31) | tin_next;
}

// r = a^2 * 2^-2048 mod q.  `as`: the 64 limbs of a staged in shared memory; `scratch`: SQS_WORDS
// words of shared memory private to the group, whose SQS_ZERO words are zero.  r may alias a.
MP_DEV void mont_sqr_split(uint32_t (&r)[8], const uint32_t (&a)[8], const uint32_t* as, uint32_t* scratch,
                           const SqPlan& pl, const Mod<8>& M, const Lane& ln) {
  // ---- 1. block products ----
  {
    fp256::Fe A, B;
#pragma unroll
    for (int i = 0; i < 8; ++i) A.v[i] = a[i];
    uint32_t t[16];
    fpsp::sqr_wide(t, A);
    sq_store16(scratch + SQS_SQ + ln.k * 16, t);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      sq_load8(B.v, as + pl.partner[d]);
      fpsp::mul_wide(t, A, B);
      sq_store16(scratch + pl.cr[d], t);
    }
  }
  simt::syncwarp();
  // ---- 2. column sums: T blocks (carries between blocks stay separate) ----
  {
    uint32_t X[9], Y[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) X[i] = Y[i] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) sq_add8(X, scratch + pl.px[i]);
#pragma unroll
    for (int i = 0; i < 7; ++i) sq_add8(Y, scratch + pl.py[i]);
    sq_double9(X);
    sq_double9(Y);
    sq_add8(X, scratch + pl.sqx);
    sq_add8(Y, scratch + pl.sqy);
    sq_store_block(scratch, pl.tx, pl.cx, X);
    sq_store_block(scratch, pl.ty, pl.cy, Y);
  }
  simt::syncwarp();
  // ---- 3. reduction ----
  uint32_t A0[10], A1[10];
  {
    uint32_t v[8];
    sq_load8(v, scratch + SQS_T + 8 * ln.k);
    A0[0] = simt::add_cc(v[0], scratch[pl.cin0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) A0[i] = simt::addc_cc(v[i], 0);
    A0[8] = simt::addc(scratch[pl.c8], 0);
    A0[9] = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) A1[i] = 0;
  }
  uint32_t in = 0;
#pragma unroll
  for (int j = 0; j < 64; j += 2) {
    {
      const uint32_t tn = scratch[pl.tin + (uint32_t)j * pl.tin_stride];
      if ((j & 7) == 1 && j > 1)
        red_digit<true>(A0, A1, M, ln, in, scratch[pl.cst + (uint32_t)((j - 1) / 8 - 1) * pl.cst_stride], tn);
      else
        red_digit<false>(A0, A1, M, ln, in, 0u, tn);
    }
    {
      const int j1 = j + 1;
      const uint32_t tn = scratch[pl.tin + (uint32_t)j1 * pl.tin_stride];
      if ((j1 & 7) == 1 && j1 > 1)
        red_digit<true>(A1, A0, M, ln, in, scratch[pl.cst + (uint32_t)((j1 - 1) / 8 - 1) * pl.cst_stride], tn);
      else
        red_digit<false>(A1, A0, M, ln, in, 0u, tn);
    }
  }
  simt::syncwarp();  // scratch is free again for the next squaring
  mm_finish<8>(r, A0, A1, in, M, ln);
}

}  // namespace modp
