// CPU execution of the elliptic-curve kernel bodies (tests only): thread-per-instance bodies are
// simply called in a loop; simt.h's emulation back-end supplies the carry-chain primitives.
#include "../../mpvss_rs_b200/csrc/secp.cuh"
#include "../../mpvss_rs_b200/csrc/rist.cuh"
#include "../../mpvss_rs_b200/csrc/ec_kernels.cuh"
#include "../../mpvss_rs_b200/csrc/sha2_dev.cuh"
#include <vector>

namespace {
template <class Cv>
int t_exp2(const void* consts, const uint8_t* b1, uint32_t b1s, const uint32_t* e1, uint32_t e1s, const uint8_t* b2,
           uint32_t b2s, const uint32_t* e2, uint32_t e2s, uint32_t n, uint8_t* out, uint32_t* status) {
  ec::Exp2Args<Cv> A{(const typename Cv::Consts*)consts, b1, e1, b2, e2, out, nullptr, status, n, b1s, e1s, b2s, e2s, nullptr};
  for (uint32_t t = 0; t < n; ++t) ec::exp2_body<Cv>(A, t);
  return 0;
}
// fixed-base table of the generator, then e1 * G [+ e2 * B2] through it and e * G alone
template <class Cv>
int t_comb(const void* consts, const uint8_t* gen, uint32_t* tbl, const uint32_t* e1, const uint8_t* b2,
           const uint32_t* e2, uint32_t n, uint8_t* out_fixed, uint8_t* out_exp2, uint32_t* status) {
  const typename Cv::Consts* C = (const typename Cv::Consts*)consts;
  ec::CombArgs<Cv> B{C, gen, tbl};
  for (uint32_t t = 0; t < (uint32_t)ec::COMB_ENTRIES; ++t) ec::comb_build_body<Cv>(B, t);
  ec::FixedArgs<Cv> F{C, tbl, e1, out_fixed, n};
  for (uint32_t t = 0; t < n; ++t) ec::fixed_body<Cv>(F, t, tbl);
  ec::Exp2Args<Cv> A{C, gen, e1, b2, e2, out_exp2, nullptr, status, n, 0, 8, (uint32_t)Cv::EB, 8, tbl};
  for (uint32_t t = 0; t < n; ++t) ec::exp2_body<Cv>(A, t, tbl);
  return 0;
}
template <class Cv>
int t_add(const void* consts, const uint8_t* a, const uint8_t* b, uint32_t n, uint8_t* out, uint32_t* status) {
  ec::AddArgs<Cv> A{(const typename Cv::Consts*)consts, a, b, out, status, n};
  for (uint32_t t = 0; t < n; ++t) ec::add_body<Cv>(A, t);
  return 0;
}
// chunked Horner end to end: decode commitments, K*n partials, per-position sum, encode
template <class Cv>
int t_poly_eval_exp(const void* consts, const uint8_t* commitments, uint32_t t, const uint32_t* pos, uint32_t n,
                    uint32_t K, uint8_t* out, uint32_t* status) {
  const typename Cv::Consts* C = (const typename Cv::Consts*)consts;
  std::vector<uint32_t> cxy((size_t)t * 16), cst(t);
  ec::DecodeArgs<Cv> D{C, commitments, cxy.data(), cst.data(), t};
  for (uint32_t i = 0; i < t; ++i) ec::decode_body<Cv>(D, i);
  for (uint32_t i = 0; i < t; ++i) status[i] = cst[i];
  uint32_t B = (t + K - 1) / K;
  K = (t + B - 1) / B;
  std::vector<typename Cv::Point> part((size_t)K * n);
  ec::HornerArgs<Cv> H{C, cxy.data(), cst.data(), pos, part.data(), t, n, K, B};
  for (uint32_t i = 0; i < K * n; ++i) ec::horner_body<Cv>(H, i);
  ec::SumArgs<Cv> S{C, part.data(), nullptr, out, n, K, 1, n, K * n};
  for (uint32_t i = 0; i < n; ++i) ec::sum_body<Cv>(S, i);
  return 0;
}
}  // namespace

#define EXPORT_CURVE(prefix, Cv)                                                                                      \
  extern "C" int emu_##prefix##_sizeof_consts() { return (int)sizeof(Cv::Consts); }                                   \
  extern "C" int emu_##prefix##_exp2(const void* c, const uint8_t* b1, uint32_t b1s, const uint32_t* e1, uint32_t e1s, \
                                     const uint8_t* b2, uint32_t b2s, const uint32_t* e2, uint32_t e2s, uint32_t n,   \
                                     uint8_t* out, uint32_t* st) {                                                    \
    return t_exp2<Cv>(c, b1, b1s, e1, e1s, b2, b2s, e2, e2s, n, out, st);                                             \
  }                                                                                                                   \
  extern "C" int emu_##prefix##_comb(const void* c, const uint8_t* gen, uint32_t* tbl, const uint32_t* e1,             \
                                     const uint8_t* b2, const uint32_t* e2, uint32_t n, uint8_t* of, uint8_t* oe,     \
                                     uint32_t* st) {                                                                  \
    return t_comb<Cv>(c, gen, tbl, e1, b2, e2, n, of, oe, st);                                                        \
  }                                                                                                                   \
  extern "C" int emu_##prefix##_add(const void* c, const uint8_t* a, const uint8_t* b, uint32_t n, uint8_t* out,      \
                                    uint32_t* st) {                                                                   \
    return t_add<Cv>(c, a, b, n, out, st);                                                                            \
  }                                                                                                                   \
  extern "C" int emu_##prefix##_poly_eval_exp(const void* c, const uint8_t* cm, uint32_t t, const uint32_t* pos,      \
                                              uint32_t n, uint32_t K, uint8_t* out, uint32_t* st) {                   \
    return t_poly_eval_exp<Cv>(c, cm, t, pos, n, K, out, st);                                                         \
  }

EXPORT_CURVE(secp, secp::SecpCurve)
EXPORT_CURVE(rist, rist::RistCurve)

extern "C" {
int emu_ec_poly(const void* modN, const uint32_t* coeffs, uint32_t t, const uint32_t* pos, uint32_t n, uint32_t* out) {
  ec::PolyArgs A{(const fp256::Modulus*)modN, coeffs, pos, out, t, n};
  for (uint32_t i = 0; i < n; ++i) ec::poly_body(A, i);
  return 0;
}
int emu_ec_lagrange(const void* modN, const uint32_t* pos, uint32_t k, uint32_t* out) {
  ec::LagrangeArgs A{(const fp256::Modulus*)modN, pos, out, k};
  for (uint32_t i = 0; i < k; ++i) ec::lagrange_body(A, i);
  return 0;
}
int emu_ec_inv(const void* modN, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* status) {
  ec::InvArgs A{(const fp256::Modulus*)modN, in, out, status, n};
  for (uint32_t i = 0; i < n; ++i) ec::inv_body(A, i);
  return 0;
}
// per-share transcripts on the device: rows -> hash_to_scalar's integer (sha2_dev.cuh), then the scalar-field half
int emu_row_hash(const uint8_t* rows, uint32_t row_stride, uint32_t slot_stride, uint32_t* out, uint32_t out_stride,
                 uint8_t* digest_out, uint32_t n, uint32_t wide) {
  shadev::RowHashArgs A{rows, row_stride, slot_stride, out, out_stride, digest_out, n, wide};
  for (uint32_t i = 0; i < n; ++i) shadev::row_hash_body(A, i);
  return 0;
}
int emu_box_hash(const uint8_t* rows, uint32_t row_stride, uint32_t slot_stride, uint32_t n, uint8_t* digest_out) {
  shadev::BoxHashArgs A{rows, row_stride, slot_stride, n, digest_out};
  for (uint32_t i = 0; i < 2; ++i) shadev::box_hash_body(A, i);
  return 0;
}
int emu_ec_proof(const void* modN, const uint32_t* h, const uint32_t* sk, const uint32_t* w, const uint8_t* c_in,
                 uint8_t* c_out, uint8_t* r_out, uint32_t* ok, uint32_t n, uint32_t wide, uint32_t big_endian) {
  ec::ProofArgs A{(const fp256::Modulus*)modN, h, sk, w, c_in, c_out, r_out, ok, n, wide, big_endian};
  for (uint32_t i = 0; i < n; ++i) ec::proof_body(A, i);
  return 0;
}
// special-form fields: out_mul = a*b mod p, out_sqr = a*a mod p (which: 0 secp256k1, 1 curve25519)
int emu_fpsp_ops(int which, const void* mod, const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* mul,
                 uint32_t* sqr) {
  const fp256::Modulus& M = *(const fp256::Modulus*)mod;
  for (uint32_t i = 0; i < n; ++i) {
    fp256::Fe x = fp256::load(a + 8 * i), y = fp256::load(b + 8 * i);
    if (which == 0) {
      fp256::store(mul + 8 * i, secp::F::mul(x, y, M));
      fp256::store(sqr + 8 * i, secp::F::sqr(x, M));
    } else {
      fp256::store(mul + 8 * i, rist::F::mul(x, y, M));
      fp256::store(sqr + 8 * i, rist::F::sqr(x, M));
    }
  }
  return 0;
}
// field-level checks: out = a*b/R, a+b, a-b, a^-1 (mod the modulus in `mod`)
int emu_fp_ops(const void* mod, const uint32_t* a, const uint32_t* b, uint32_t n, uint32_t* mul, uint32_t* add,
               uint32_t* sub, uint32_t* inv) {
  const fp256::Modulus& M = *(const fp256::Modulus*)mod;
  for (uint32_t i = 0; i < n; ++i) {
    fp256::Fe x = fp256::load(a + 8 * i), y = fp256::load(b + 8 * i);
    fp256::store(mul + 8 * i, fp256::mul(x, y, M));
    fp256::store(add + 8 * i, fp256::add(x, y, M));
    fp256::store(sub + 8 * i, fp256::sub(x, y, M));
    fp256::store(inv + 8 * i, fp256::from_mont(fp256::inv(fp256::to_mont(x, M), M), M));
  }
  return 0;
}
}
