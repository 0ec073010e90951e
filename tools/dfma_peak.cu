// FP64 multiply-add probe for B200 (sm_100a): can a 52-bit-limb DFMA formulation of the 2048-bit Montgomery
// product beat IMAD.WIDE?  Measures per SM and clock: independent DFMA, the exact "limb product" shape
// (DFMA.RZ hi, DADD, DFMA.RZ lo, IADD3 + IADD3.X accumulate: 2704 bits^2 of product per 5 instructions), the
// same without the integer accumulation, and a DFMA : IADD3 1:1 mix (does the ALU pipe issue in the shadow of
// the FP64 pipe?).  Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define ITERS 2048

template <int MODE>
__global__ void __launch_bounds__(256) probe(unsigned long long* out, double seed, long long* cycles, const double* bsrc) {
  __shared__ double bs[64];
  if (threadIdx.x < 64) bs[threadIdx.x] = bsrc[threadIdx.x];
  __syncthreads();
  double a[5], m[5];
  unsigned long long T[6];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    a[i] = seed + (double)(threadIdx.x * 5 + i);
    m[i] = seed * 3.0 + (double)(threadIdx.x * 7 + i);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) T[i] = threadIdx.x + i;
  double f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = seed + i;
  const double c1 = 0x1p104, c2 = 0x1p104 + 0x1p52;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {  // 16 independent DFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], 1.0000001, 0.5);
    } else if (MODE == 1 || MODE == 2) {  // one digit of a 5-limbs-per-lane Montgomery step: 10 limb products
      double b = bs[it & 63];
      double q = bs[(it + 17) & 63];
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        double hi = __fma_rz(a[i], b, c1), lo = __fma_rz(a[i], b, c2 - hi);
        double hi2 = __fma_rz(m[i], q, c1), lo2 = __fma_rz(m[i], q, c2 - hi2);
        if (MODE == 1) {
          T[i] += (unsigned long long)__double_as_longlong(lo) + (unsigned long long)__double_as_longlong(lo2);
          T[i + 1] += (unsigned long long)__double_as_longlong(hi) + (unsigned long long)__double_as_longlong(hi2);
        } else {
          f[i] += lo + lo2;  // keeps the values alive with FP64 instructions only (2 more DADD per pair)
          f[i + 8] += hi + hi2;
        }
      }
      if (MODE == 1) {  // shift the window down one column, as the real loop does
        unsigned long long c = T[0] >> 52;
#pragma unroll
        for (int i = 0; i < 5; ++i) T[i] = T[i + 1];
        T[0] += c; T[5] = 0;
      }
    } else if (MODE == 3) {  // 8 DFMA + 8 IADD3
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = fma(f[i], 1.0000001, 0.5);
        T[i % 6] += (unsigned long long)(it + i) ;
      }
    } else if (MODE == 4) {  // 8 DFMA + 16 32-bit integer adds
      unsigned* t32 = reinterpret_cast<unsigned*>(T);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = fma(f[i], 1.0000001, 0.5);
        t32[i] += it ^ i; t32[(i + 3) % 12] ^= t32[i];
      }
    }
  }
  long long t1 = clock64();
  unsigned long long x = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) x ^= T[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= (unsigned long long)__double_as_longlong(f[i]);
  if (x == 0x12345678u) out[0] = x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Res { double per_clk_sm; double per_sec; };

template <int MODE>
Res run(int sms, int ctas_per_sm, int threads, int ops_per_iter, unsigned long long* dout, long long* dcyc, const double* db) {
  int grid = sms * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<grid, threads>>>(dout, 12345.0, dcyc, db);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, threads>>>(dout, 12345.0 + rep, dcyc, db);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  std::vector<long long> cyc(grid);
  cudaMemcpy(cyc.data(), dcyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  std::sort(cyc.begin(), cyc.end());
  double med = (double)cyc[grid / 2];
  double ops_per_sm = (double)ops_per_iter * ITERS * threads * ctas_per_sm;
  Res r;
  r.per_clk_sm = ops_per_sm / med;
  r.per_sec = ops_per_sm * sms / (best * 1e-3);
  return r;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  unsigned long long* dout; long long* dcyc; double* db;
  cudaMalloc(&dout, 64); cudaMalloc(&dcyc, sizeof(long long) * sms * 8); cudaMalloc(&db, 64 * 8);
  std::vector<double> hb(64);
  for (int i = 0; i < 64; ++i) hb[i] = 4503599627370495.0 - i * 1234567.0;
  cudaMemcpy(db, hb.data(), 64 * 8, cudaMemcpyHostToDevice);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d", p.name, sms, p.clockRate);
  // ops counted: mode 0: 16 DFMA; modes 1, 2: 10 limb products; mode 3: 8 DFMA; mode 4: 8 DFMA
  const char* names[5] = {"dfma", "limb_product_full", "limb_product_fp_only", "dfma_with_iadd64", "dfma_with_2alu"};
  int ops[5] = {16, 10, 10, 8, 8};
  for (int warps : {4, 8, 16, 32}) {   // warps per SM (1 CTA per SM)
    int threads = warps * 32 > 256 ? 256 : warps * 32, cps = warps * 32 / threads;
    Res r[5];
    r[0] = run<0>(sms, cps, threads, ops[0], dout, dcyc, db);
    r[1] = run<1>(sms, cps, threads, ops[1], dout, dcyc, db);
    r[2] = run<2>(sms, cps, threads, ops[2], dout, dcyc, db);
    r[3] = run<3>(sms, cps, threads, ops[3], dout, dcyc, db);
    r[4] = run<4>(sms, cps, threads, ops[4], dout, dcyc, db);
    for (int k = 0; k < 5; ++k)
      printf(", \"%s_w%d\": {\"per_clk_sm\": %.2f, \"tera_per_s\": %.3f}", names[k], warps, r[k].per_clk_sm, r[k].per_sec / 1e12);
  }
  printf("}\n");
  return 0;
}
