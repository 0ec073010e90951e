// Lane-per-thread emulator harness for the ModpGroup kernel bodies (tests only).
// Builds mpvss_rs_b200/csrc/modp_kernels.cuh with g++ (-DMPVSS_SIMT_EMU) so the
// GPU-less container can execute the exact kernel source against the oracle.
#include <thread>
#include <vector>
#include <cstdlib>
#include "../../mpvss_rs_b200/csrc/modp_kernels.cuh"

namespace {
template <typename F>
void run_warps(uint32_t nwarps, size_t smem_words, F body) {
  for (uint32_t w = 0; w < nwarps; ++w) {
    simt::EmuWarp warp;
    std::vector<uint32_t> smem_raw(smem_words + 4, 0);
    uint32_t* smem = smem_raw.data();
    while (reinterpret_cast<uintptr_t>(smem) & 15) ++smem;
    std::vector<std::thread> lanes;
    for (uint32_t l = 0; l < 32; ++l)
      lanes.emplace_back([&, l] {
        simt::g_lane.warp = &warp;
        simt::g_lane.lane = l;
        simt::g_cf = 0;
        body(w, smem);
      });
    for (auto& t : lanes) t.join();
  }
}
template <int TPI> uint32_t warps_for(uint32_t n) { return (n + 32 / TPI - 1) / (32 / TPI); }
}  // namespace

#define DISPATCH(tpi, CALL)        \
  switch (tpi) {                   \
    case 4: { constexpr int T = 4; CALL; break; }   \
    case 8: { constexpr int T = 8; CALL; break; }   \
    case 16: { constexpr int T = 16; CALL; break; } \
    default: return -1;            \
  }

// `first` / `steps`: the chunk of the polynomial to evaluate (top coefficient index, Horner steps); the whole
// polynomial is (t - 1, t - 1)
extern "C" int emu_modp_horner(int tpi, const uint32_t* consts, const uint32_t* cm, uint32_t t, const uint16_t* ops,
                               uint32_t n, uint32_t nops, uint32_t* out, uint32_t first, uint32_t steps) {
  const bool np1 = consts[modp::C_NP] == 1u;
  // the whole polynomial runs the unchunked instantiation, anything else the chunked one (one chunk for all CTAs)
  const bool chunked = !(first == t - 1 && steps == t - 1);
  modp::HornerArgs A{consts, cm, ops, nullptr, nullptr, out, t, n, nops, 1, chunked ? &first : nullptr,
                     chunked ? &steps : nullptr};
  DISPATCH(tpi, run_warps(warps_for<T>(n), modp::horner_smem_words<T>, [&](uint32_t w, uint32_t* s) {
             if (chunked) {
               if (np1) modp::horner_body<T, true, true>(A, w, s, nops, 0);
               else modp::horner_body<T, false, true>(A, w, s, nops, 0);
             } else {
               if (np1) modp::horner_body<T, true, false>(A, w, s, nops, 0);
               else modp::horner_body<T, false, false>(A, w, s, nops, 0);
             }
           }));
  return 0;
}

extern "C" int emu_modp_frames(const uint32_t* x, const uint32_t* y, const uint32_t* a1, const uint32_t* a2, uint32_t n,
                               uint8_t* out) {
  modp::FrameArgs A{x, y, a1, a2, out, n};
  for (uint32_t i = 0; i < 4 * n; ++i) modp::frame_body(A, i);
  return 0;
}

extern "C" int emu_modp_resp(const uint32_t* order, const uint32_t* alpha, uint32_t alpha_stride, const uint32_t* w,
                             const uint32_t* c, uint32_t c_stride, uint32_t n, uint32_t* out) {
  modp::RespArgs A{order, alpha, w, c, out, n, alpha_stride, c_stride};
  for (uint32_t i = 0; i < n; ++i) modp::resp_body(A, i);
  return 0;
}

// bucket multi-exponentiation end to end (8 lanes per value): buckets, window products, fold
extern "C" int emu_modp_msm(const uint32_t* consts, const uint32_t* bases_mont, const uint32_t* idx,
                            const uint32_t* start, uint32_t windows, uint32_t k, uint32_t* buckets, uint32_t* wprod,
                            uint32_t* out) {
  constexpr int T = 8;
  modp::MsmBucketArgs B{consts, bases_mont, idx, start, buckets, windows, k};
  run_warps(warps_for<T>(windows * 255), modp::msm_smem_words<T>,
            [&](uint32_t w, uint32_t* s) { modp::msm_bucket_body<T>(B, w, s); });
  modp::MsmWindowArgs W{consts, buckets, wprod, windows};
  run_warps(windows, modp::msm_window_smem_words<T>, [&](uint32_t w, uint32_t* s) { modp::msm_window_body<T>(W, w, s); });
  modp::MsmFoldArgs F{consts, wprod, out, windows};
  run_warps(1, modp::msm_smem_words<T>, [&](uint32_t, uint32_t* s) { modp::msm_fold_body<T>(F, s); });
  return 0;
}

extern "C" int emu_modp_exp2(int tpi, const uint32_t* consts, const uint32_t* b1, uint32_t b1s, const uint32_t* e1,
                             uint32_t e1s, uint32_t e1w, const uint32_t* b2, uint32_t b2s, const uint32_t* e2,
                             uint32_t e2s, uint32_t e2w, uint32_t n, uint32_t* out, const uint32_t* comb) {
  modp::Exp2Args A{consts, b1, e1, b2, e2, out, n, b1s, e1s, e1w, b2s, e2s, e2w, comb};
  DISPATCH(tpi, run_warps(warps_for<T>(n), modp::exp2_smem_words<T>,
                          [&](uint32_t w, uint32_t* s) { modp::exp2_body<T>(A, w, s); }));
  return 0;
}

extern "C" int emu_modp_mul(int tpi, const uint32_t* consts, const uint32_t* a, uint32_t as, const uint32_t* b,
                            uint32_t bs, uint32_t n, uint32_t mode, uint32_t* out) {
  modp::MulArgs A{consts, a, b, out, n, mode, as, bs};
  DISPATCH(tpi, run_warps(warps_for<T>(n), modp::mul_smem_words<T>,
                          [&](uint32_t w, uint32_t* s) { modp::mul_body<T>(A, w, s); }));
  return 0;
}

// split squaring (TPI = 8): out = a^(2^reps) / R^(2^reps - 1) mod q
extern "C" int emu_modp_sqr_split(const uint32_t* consts, const uint32_t* a, uint32_t n, uint32_t reps, uint32_t* out) {
  modp::MulArgs A{consts, a, nullptr, out, n, reps, 64, 0};
  run_warps((n + 3) / 4, modp::sqrtest_smem_words, [&](uint32_t w, uint32_t* s) { modp::sqr_split_test_body(A, w, s); });
  return 0;
}

// fixed-base table: only the first `rows` byte positions are filled (step 1 is cut short by the test)
extern "C" int emu_modp_comb_build(int tpi, const uint32_t* consts, const uint32_t* base, uint32_t* tbl,
                                   uint32_t rows) {
  modp::CombArgs A{consts, base, tbl, rows};
  DISPATCH(tpi, {
    run_warps(1, modp::comb_smem_words<T>, [&](uint32_t w, uint32_t* s) { modp::comb1_body<T>(A, w, s); });
    run_warps(warps_for<T>(rows), modp::comb_smem_words<T>, [&](uint32_t w, uint32_t* s) { modp::comb2_body<T>(A, w, s); });
  });
  return 0;
}

extern "C" int emu_modp_poly(const uint32_t* coeffs, uint32_t t, const uint32_t* order, const uint32_t* pos, uint32_t n,
                             uint32_t* out) {
  modp::PolyArgs A{coeffs, order, pos, out, t, n};
  for (uint32_t i = 0; i < n; ++i) modp::poly_body(A, i);
  return 0;
}

// num / den / negative: parts x k rows (part p of position i in row p * k + i)
extern "C" int emu_modp_lagrange(const uint32_t* order, const uint32_t* pos, uint32_t k, uint32_t* num, uint32_t* den,
                                 uint32_t* negative, uint32_t parts) {
  modp::LagrangeArgs A{order, pos, num, den, negative, k, parts};
  for (uint32_t i = 0; i < 2 * (parts ? parts : 1) * k + 3; ++i) modp::lagrange_body(A, i);
  return 0;
}
