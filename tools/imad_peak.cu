// Integer multiply-add roofline probe for B200 (sm_100a).
// Measures, per SM and clock, the issue rate of the instruction shapes the MODP
// kernels are made of: IMAD.WIDE.U32 (independent), IMAD.WIDE.U32.X carry chains
// (mad.lo.cc/madc.hi.cc as ptxas fuses them), 32-bit IMAD lo / hi, a 1:1 mix of
// IMAD.WIDE with IADD3, and SHFL.  Prints one JSON object; bench.py and DESIGN.md
// use "wide_mac_per_clk_sm" x SMs x clock as the denominator of the IMAD roofline.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t a = seed * (threadIdx.x + 1) | 1u, b = seed ^ (blockIdx.x * 2654435761u) | 1u;
  uint32_t r0 = a, r1 = b, r2 = a ^ b, r3 = a + b, r4 = a * 3, r5 = b * 5, r6 = a * 7, r7 = b * 9;
  uint32_t r8 = a + 1, r9 = b + 2, r10 = a + 3, r11 = b + 4, r12 = a + 5, r13 = b + 6, r14 = a + 7, r15 = b + 8;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {  // 8 independent 64-bit accumulators: IMAD.WIDE.U32
      asm volatile(
          "{ .reg .u64 t;\n"
          "mov.b64 t,{%0,%1};  mad.wide.u32 t,%2,%17,t; mov.b64 {%0,%1},t;\n"
          "mov.b64 t,{%2,%3};  mad.wide.u32 t,%4,%17,t; mov.b64 {%2,%3},t;\n"
          "mov.b64 t,{%4,%5};  mad.wide.u32 t,%6,%17,t; mov.b64 {%4,%5},t;\n"
          "mov.b64 t,{%6,%7};  mad.wide.u32 t,%8,%17,t; mov.b64 {%6,%7},t;\n"
          "mov.b64 t,{%8,%9};  mad.wide.u32 t,%10,%17,t; mov.b64 {%8,%9},t;\n"
          "mov.b64 t,{%10,%11};  mad.wide.u32 t,%12,%17,t; mov.b64 {%10,%11},t;\n"
          "mov.b64 t,{%12,%13};  mad.wide.u32 t,%14,%17,t; mov.b64 {%12,%13},t;\n"
          "mov.b64 t,{%14,%15};  mad.wide.u32 t,%0,%17,t; mov.b64 {%14,%15},t;\n"
          "}\n"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9),
            "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15)
          : "r"(a), "r"(b));
    } else if (MODE == 1) {  // two carry chains of 4 wide MACs each (the kernel's shape)
      asm volatile(
          "mad.lo.cc.u32 %0,%8,%17,%0;  madc.hi.cc.u32 %1,%8,%17,%1;\n"
          "madc.lo.cc.u32 %2,%10,%17,%2; madc.hi.cc.u32 %3,%10,%17,%3;\n"
          "madc.lo.cc.u32 %4,%12,%17,%4; madc.hi.cc.u32 %5,%12,%17,%5;\n"
          "madc.lo.cc.u32 %6,%14,%17,%6; madc.hi.u32 %7,%14,%17,%7;\n"
          "mad.lo.cc.u32 %8,%0,%17,%8;  madc.hi.cc.u32 %9,%0,%17,%9;\n"
          "madc.lo.cc.u32 %10,%2,%17,%10; madc.hi.cc.u32 %11,%2,%17,%11;\n"
          "madc.lo.cc.u32 %12,%4,%17,%12; madc.hi.cc.u32 %13,%4,%17,%13;\n"
          "madc.lo.cc.u32 %14,%6,%17,%14; madc.hi.u32 %15,%6,%17,%15;\n"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9),
            "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15)
          : "r"(a), "r"(b));
    } else if (MODE == 2) {  // 16 independent 32-bit IMAD (lo)
      asm volatile(
          "mad.lo.u32 %0,%5,%17,%0; mad.lo.u32 %1,%6,%17,%1; mad.lo.u32 %2,%7,%17,%2; mad.lo.u32 %3,%8,%17,%3;\n"
          "mad.lo.u32 %4,%9,%17,%4; mad.lo.u32 %5,%10,%17,%5; mad.lo.u32 %6,%11,%17,%6; mad.lo.u32 %7,%12,%17,%7;\n"
          "mad.lo.u32 %8,%13,%17,%8; mad.lo.u32 %9,%14,%17,%9; mad.lo.u32 %10,%15,%17,%10; mad.lo.u32 %11,%0,%17,%11;\n"
          "mad.lo.u32 %12,%1,%17,%12; mad.lo.u32 %13,%2,%17,%13; mad.lo.u32 %14,%3,%17,%14; mad.lo.u32 %15,%4,%17,%15;\n"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9),
            "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15)
          : "r"(a), "r"(b));
    } else if (MODE == 3) {  // 16 independent IMAD.HI
      asm volatile(
          "mad.hi.u32 %0,%5,%17,%0; mad.hi.u32 %1,%6,%17,%1; mad.hi.u32 %2,%7,%17,%2; mad.hi.u32 %3,%8,%17,%3;\n"
          "mad.hi.u32 %4,%9,%17,%4; mad.hi.u32 %5,%10,%17,%5; mad.hi.u32 %6,%11,%17,%6; mad.hi.u32 %7,%12,%17,%7;\n"
          "mad.hi.u32 %8,%13,%17,%8; mad.hi.u32 %9,%14,%17,%9; mad.hi.u32 %10,%15,%17,%10; mad.hi.u32 %11,%0,%17,%11;\n"
          "mad.hi.u32 %12,%1,%17,%12; mad.hi.u32 %13,%2,%17,%13; mad.hi.u32 %14,%3,%17,%14; mad.hi.u32 %15,%4,%17,%15;\n"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9),
            "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15)
          : "r"(a), "r"(b));
    } else if (MODE == 4) {  // 4 wide MACs + 8 independent ALU ops (does the ALU pipe co-issue?)
      asm volatile(
          "{ .reg .u64 t;\n"
          "mov.b64 t,{%0,%1};  mad.wide.u32 t,%2,%17,t; mov.b64 {%0,%1},t;\n"
          "add.u32 %8,%8,%9;  xor.b32 %9,%9,%10;\n"
          "mov.b64 t,{%2,%3};  mad.wide.u32 t,%4,%17,t; mov.b64 {%2,%3},t;\n"
          "add.u32 %10,%10,%11; xor.b32 %11,%11,%12;\n"
          "mov.b64 t,{%4,%5};  mad.wide.u32 t,%6,%17,t; mov.b64 {%4,%5},t;\n"
          "add.u32 %12,%12,%13; xor.b32 %13,%13,%14;\n"
          "mov.b64 t,{%6,%7};  mad.wide.u32 t,%0,%17,t; mov.b64 {%6,%7},t;\n"
          "add.u32 %14,%14,%15; xor.b32 %15,%15,%8; }\n"
          : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8), "+r"(r9),
            "+r"(r10), "+r"(r11), "+r"(r12), "+r"(r13), "+r"(r14), "+r"(r15)
          : "r"(a), "r"(b));
    } else if (MODE == 5) {  // 8 SHFL
      r0 = __shfl_sync(0xffffffffu, r0, (threadIdx.x + 1) & 31);
      r1 = __shfl_sync(0xffffffffu, r1, (threadIdx.x + 2) & 31);
      r2 = __shfl_sync(0xffffffffu, r2, (threadIdx.x + 3) & 31);
      r3 = __shfl_sync(0xffffffffu, r3, (threadIdx.x + 4) & 31);
      r4 = __shfl_sync(0xffffffffu, r4, (threadIdx.x + 5) & 31);
      r5 = __shfl_sync(0xffffffffu, r5, (threadIdx.x + 6) & 31);
      r6 = __shfl_sync(0xffffffffu, r6, (threadIdx.x + 7) & 31);
      r7 = __shfl_sync(0xffffffffu, r7, (threadIdx.x + 8) & 31);
    }
  }
  long long t1 = clock64();
  uint32_t x = r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7 ^ r8 ^ r9 ^ r10 ^ r11 ^ r12 ^ r13 ^ r14 ^ r15;
  if (x == 0x12345678u) out[0] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Res { double per_clk_sm; double per_sec; };

template <int MODE>
Res run(int sms, int ctas_per_sm, int ops_per_iter, uint32_t* dout, long long* dcyc) {
  int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<grid, 256>>>(dout, 12345u, dcyc);  // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, 256>>>(dout, 12345u + rep, dcyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(cyc.begin(), cyc.end());
  double med = (double)cyc[grid / 2];
  double ops_per_sm = (double)ops_per_iter * ITERS * 256.0 * ctas_per_sm;
  Res r;
  r.per_clk_sm = ops_per_sm / med;           // per-CTA cycle count with ctas_per_sm co-resident CTAs
  r.per_sec = ops_per_sm * sms / (best * 1e-3);
  return r;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  uint32_t* dout; long long* dcyc;
  cudaMalloc(&dout, 64); cudaMalloc(&dcyc, sizeof(long long) * sms * 8);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);
  const char* names[6] = {"imad_wide", "imad_wide_carry_chain", "imad_lo", "imad_hi", "imad_wide_plus_alu", "shfl"};
  int ops[6] = {8, 8, 16, 16, 4, 8};
  for (int cps : {1, 2, 4, 8}) {
    Res r[6];
    r[0] = run<0>(sms, cps, ops[0], dout, dcyc);
    r[1] = run<1>(sms, cps, ops[1], dout, dcyc);
    r[2] = run<2>(sms, cps, ops[2], dout, dcyc);
    r[3] = run<3>(sms, cps, ops[3], dout, dcyc);
    r[4] = run<4>(sms, cps, ops[4], dout, dcyc);
    r[5] = run<5>(sms, cps, ops[5], dout, dcyc);
    for (int m = 0; m < 6; ++m)
      printf(", \"%s_w%d\": {\"per_clk_sm\": %.2f, \"tera_per_s\": %.3f}", names[m], cps * 8, r[m].per_clk_sm,
             r[m].per_sec / 1e12);
  }
  printf("}\n");
  return 0;
}
