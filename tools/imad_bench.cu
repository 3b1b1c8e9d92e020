// Micro-benchmark: issue rate of the integer pipes that bound MSM / big-field NTT on sm_100a.
// Prints G instr/s per kernel for IMAD (32-bit), IMAD.WIDE.U32 (+carry chain) and IADD3 chains, plus a full
// 8-limb Montgomery multiply loop (ff.cuh) -- the denominators for the "fraction of IMAD peak" figures in DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>
#include "../icicle_b200/csrc/ff.cuh"
using namespace b200;

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(uint32_t* out, uint32_t seed, int iters)
{
  uint32_t a0 = threadIdx.x + seed, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3;
  uint32_t b0 = a3 ^ 0x1234567, b1 = a2 ^ 0x89abcde, b2 = a1 ^ 0x3141592, b3 = a0 ^ 0x2718281;
  uint32_t x = seed | 1;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      if (MODE == 0) { // 4 independent 32-bit IMAD chains
        a0 = a0 * x + b0; a1 = a1 * x + b1; a2 = a2 * x + b2; a3 = a3 * x + b3;
      } else if (MODE == 1) { // 4 independent IMAD.WIDE.U32 chains (64-bit accumulate, no carry flag)
        uint64_t t0 = (uint64_t)a0 * x + (((uint64_t)b0 << 32) | a0);
        uint64_t t1 = (uint64_t)a1 * x + (((uint64_t)b1 << 32) | a1);
        uint64_t t2 = (uint64_t)a2 * x + (((uint64_t)b2 << 32) | a2);
        uint64_t t3 = (uint64_t)a3 * x + (((uint64_t)b3 << 32) | a3);
        a0 = (uint32_t)t0; b0 = (uint32_t)(t0 >> 32); a1 = (uint32_t)t1; b1 = (uint32_t)(t1 >> 32);
        a2 = (uint32_t)t2; b2 = (uint32_t)(t2 >> 32); a3 = (uint32_t)t3; b3 = (uint32_t)(t3 >> 32);
      } else if (MODE == 2) { // carry-chained IMAD.WIDE.U32.X (what the Montgomery multiply issues)
        mad_wide_cc(a0, b0, a1, x);
        madc_wide_cc(a2, b2, a3, x);
        madc_wide_cc(a1, b1, a0, x);
        madc_wide_cc(a3, b3, a2, x);
      } else if (MODE == 4) { // 4 independent DFMA chains (FP64 pipe)
        double d0 = __longlong_as_double(((long long)b0 << 32) | a0), d1 = __longlong_as_double(((long long)b1 << 32) | a1);
        double d2 = __longlong_as_double(((long long)b2 << 32) | a2), d3 = __longlong_as_double(((long long)b3 << 32) | a3);
        const double m = 1.0000001, c = 1e-9;
        d0 = fma(d0, m, c); d1 = fma(d1, m, c); d2 = fma(d2, m, c); d3 = fma(d3, m, c);
        long long l0 = __double_as_longlong(d0), l1 = __double_as_longlong(d1), l2 = __double_as_longlong(d2), l3 = __double_as_longlong(d3);
        a0 = (uint32_t)l0; b0 = (uint32_t)(l0 >> 32); a1 = (uint32_t)l1; b1 = (uint32_t)(l1 >> 32);
        a2 = (uint32_t)l2; b2 = (uint32_t)(l2 >> 32); a3 = (uint32_t)l3; b3 = (uint32_t)(l3 >> 32);
      } else if (MODE == 5) { // 2 DFMA chains + 2 carry-chained IMAD.WIDE.X: do the two pipes run concurrently?
        double d0 = __longlong_as_double(((long long)b0 << 32) | a0), d1 = __longlong_as_double(((long long)b1 << 32) | a1);
        const double m = 1.0000001, c = 1e-9;
        d0 = fma(d0, m, c); d1 = fma(d1, m, c);
        long long l0 = __double_as_longlong(d0), l1 = __double_as_longlong(d1);
        a0 = (uint32_t)l0; b0 = (uint32_t)(l0 >> 32); a1 = (uint32_t)l1; b1 = (uint32_t)(l1 >> 32);
        mad_wide_cc(a2, b2, a3, x);
        madc_wide_cc(a3, b3, a2, x);
      } else { // IADD3 chains
        a0 = a0 + b0 + x; a1 = a1 + b1 + x; a2 = a2 + b2 + x; a3 = a3 + b3 + x;
        b0 ^= a1; b1 ^= a2; b2 ^= a3; b3 ^= a0;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
}

template <class F>
__global__ void __launch_bounds__(256) k_montmul(uint32_t* out, int iters)
{
  F a, b;
  for (int i = 0; i < F::N; i++) { a.v[i] = threadIdx.x * 77 + i; b.v[i] = blockIdx.x * 13 + i * 5 + 1; }
  a.v[F::N - 1] &= 0x0fffffff; b.v[F::N - 1] &= 0x0fffffff;
  for (int i = 0; i < iters; i++) {
    a = a * b;
    b = b * a;
  }
  uint32_t r = 0;
  for (int i = 0; i < F::N; i++) r ^= a.v[i] ^ b.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <class K>
float time_ms(K launch)
{
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); launch();
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < 5; i++) launch();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main()
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, threads = 256, iters = 2000;
  uint32_t* out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  const double n_thread_instr = (double)blocks * threads * iters * 16 * 4;
  const char* names[4] = {"IMAD (32-bit)", "IMAD.WIDE.U32 (64-bit acc)", "IMAD.WIDE.U32.X (carry chain)", "IADD3+LOP3"};
  float ms;
  ms = time_ms([&] { k_pipe<0><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms)\n", names[0], n_thread_instr / ms / 1e6, ms);
  ms = time_ms([&] { k_pipe<1><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms)\n", names[1], n_thread_instr / ms / 1e6, ms);
  ms = time_ms([&] { k_pipe<2><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms)\n", names[2], n_thread_instr / ms / 1e6, ms);
  ms = time_ms([&] { k_pipe<4><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms)\n", "DFMA (FP64)", n_thread_instr / ms / 1e6, ms);
  ms = time_ms([&] { k_pipe<5><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms; 2 DFMA + 2 IMAD.WIDE.X per counted group of 4)\n", "DFMA + IMAD.WIDE.X mixed", n_thread_instr / ms / 1e6, ms);
  ms = time_ms([&] { k_pipe<3><<<blocks, threads>>>(out, 1, iters); });
  printf("%-34s %8.1f G thread-instr/s  (%.3f ms, 2 instr per counted op)\n", names[3], 2 * n_thread_instr / ms / 1e6, ms);
  {
    typedef Fp<params::bn254_fq> F;
    const int it = 500;
    ms = time_ms([&] { k_montmul<F><<<blocks, threads>>>(out, it); });
    double muls = (double)blocks * threads * it * 2;
    printf("%-34s %8.2f G mont-mul/s  (%.3f ms)  [bn254_fq, 8 limbs]\n", "mont_mul bn254", muls / ms / 1e6, ms);
  }
  {
    typedef Fp<params::bls12_381_fq> F;
    const int it = 300;
    ms = time_ms([&] { k_montmul<F><<<blocks, threads>>>(out, it); });
    double muls = (double)blocks * threads * it * 2;
    printf("%-34s %8.2f G mont-mul/s  (%.3f ms)  [bls12_381_fq, 12 limbs]\n", "mont_mul bls12-381", muls / ms / 1e6, ms);
  }
  {
    typedef Fp<params::babybear> F;
    const int it = 4000;
    ms = time_ms([&] { k_montmul<F><<<blocks, threads>>>(out, it); });
    double muls = (double)blocks * threads * it * 2;
    printf("%-34s %8.2f G mont-mul/s  (%.3f ms)  [babybear, 1 limb]\n", "mont_mul babybear", muls / ms / 1e6, ms);
  }
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("SMs %d, max clock %d MHz\n", sms, clk / 1000);
  return 0;
}
