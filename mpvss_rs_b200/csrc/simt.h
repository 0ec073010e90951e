// SIMT primitives used by the kernels: carry-chain arithmetic (PTX add.cc / madc.*),
// warp shuffles, ballots and warp barriers.
//
// Two back-ends behind one set of names:
//   * nvcc (device code): inline PTX for sm_100a.  ptxas fuses every
//     mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32(.X) with the carry in a
//     predicate register, which is what makes the even/odd accumulator layout in
//     modp_arith.cuh run on the integer multiply-add pipe without extra adds.
//   * g++ (tests only, -DMPVSS_SIMT_EMU): every lane is a host thread, the carry
//     flag is a thread_local, shuffles/ballots go through a barrier.  This lets
//     the container without a GPU execute the *same* kernel bodies bit for bit
//     (tests/emu).  It is test infrastructure, never part of the shipped library.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__) && !defined(MPVSS_SIMT_EMU)
// ------------------------------------------------------------------ device ----
#define MP_DEV __device__ __forceinline__
#define MP_HOSTDEV __host__ __device__ __forceinline__
#define MP_NOINLINE static __device__ __noinline__

namespace simt {

MP_DEV uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MP_DEV uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MP_DEV uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MP_DEV uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MP_DEV uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MP_DEV uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// d = lo(a*b) + c, sets CF
MP_DEV uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
MP_DEV uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
MP_DEV uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
MP_DEV uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
MP_DEV uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
MP_DEV uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
MP_DEV uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }

MP_DEV uint32_t lane_id() { return threadIdx.x & 31u; }
MP_DEV uint32_t shfl(uint32_t v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
MP_DEV uint32_t ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
MP_DEV void syncwarp() { __syncwarp(); }

}  // namespace simt

#else
// --------------------------------------------------------------- emulation ----
#include <barrier>
#include <cstring>
#define MP_DEV inline
#define MP_HOSTDEV inline
#define MP_NOINLINE inline

struct alignas(8) uint2 {
  uint32_t x, y;
};
struct alignas(16) uint4 {
  uint32_t x, y, z, w;
};
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace simt {

struct EmuWarp {
  std::barrier<> bar{32};
  uint32_t slot[32];
};
struct EmuLane {
  EmuWarp* warp;
  uint32_t lane;
};
inline thread_local EmuLane g_lane;
inline thread_local uint32_t g_cf;  // the PTX CC.CF flag of this lane

inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b + g_cf;
  g_cf = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + g_cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b;
  g_cf = (uint32_t)((s >> 32) & 1);  // borrow
  return (uint32_t)s;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a - b - g_cf;
  g_cf = (uint32_t)((s >> 32) & 1);
  return (uint32_t)s;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - g_cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

inline uint32_t lane_id() { return g_lane.lane; }
inline uint32_t shfl(uint32_t v, int src_lane) {
  EmuWarp* w = g_lane.warp;
  w->slot[g_lane.lane] = v;
  w->bar.arrive_and_wait();
  uint32_t r = w->slot[src_lane & 31];
  w->bar.arrive_and_wait();
  return r;
}
inline uint32_t ballot(bool p) {
  EmuWarp* w = g_lane.warp;
  w->slot[g_lane.lane] = p ? 1u : 0u;
  w->bar.arrive_and_wait();
  uint32_t r = 0;
  for (int i = 0; i < 32; ++i) r |= w->slot[i] << i;
  w->bar.arrive_and_wait();
  return r;
}
inline void syncwarp() { g_lane.warp->bar.arrive_and_wait(); }

}  // namespace simt
#endif
