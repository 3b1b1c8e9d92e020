// Developer probe (GPU box): a 254-bit Montgomery product on the FP64 pipe (5 x 52-bit limbs, R = 2^260) for BN254 Fq, next to the
// shipped IMAD.WIDE one (ff.cuh, 8 x 32-bit limbs, R = 2^256) -- DESIGN.md (f) item 1.
//   hi = fma_rz(a_i, b_j, 2^104)            -> mantissa = floor(a_i b_j / 2^52)   (exact: the sum lies in [2^104, 2^105), ulp 2^52)
//   lo = fma_rz(a_i, b_j, 2^104 + 2^52 - hi) -> mantissa = a_i b_j mod 2^52        (exact: the sum lies in [2^52, 2^53), ulp 1)
// The two bit patterns are added into 64-bit integer column accumulators that were pre-loaded with minus the sum of the
// exponent fields they will receive; Montgomery reduction limb by limb with q = (column * -p^-1) mod 2^52 on the integer pipe.
// Inputs may be anywhere in [0, 8p) and the result is < 2p (R has 6 spare bits), so chains need no conditional subtraction.
// Measures: products/s of each flavour alone, and of both flavours resident on the SMs at the same time (warps of one
// flavour fill the issue slots the other leaves idle).  Writes (a, b, result limbs) samples to gpurun_out/dfma_check.bin;
// tools/dfma_check.py verifies them with Python integers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/dfma_modmul tools/dfma_modmul.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../icicle_b200/csrc/ff.cuh"
using namespace b200;

namespace f52 {
constexpr int L = 5;
constexpr uint64_t M52 = (1ull << 52) - 1;
constexpr uint64_t C_LO = 0x433ull << 52; // bits of 2^52
constexpr uint64_t C_HI = 0x467ull << 52; // bits of 2^104
// BN254 Fq in 52-bit limbs, -p^-1 mod 2^52
__device__ __constant__ double P52[L] = {(double)0x8c16d87cfd47ull, (double)0x916871ca8d3c2ull, (double)0x181585d97816aull, (double)0xa029b85045b68ull,
                                         (double)0x30644e72e131ull};
constexpr uint64_t NP52 = 0x20782e4866389ull;
// minus the exponent fields each column receives: column k gets 2 * #{i+j=k} "lo" patterns and 2 * #{i+j=k-1} "hi" patterns
__host__ __device__ constexpr uint64_t col_init(int k)
{
  int nlo = 0, nhi = 0;
  for (int i = 0; i < L; i++)
    for (int j = 0; j < L; j++) {
      if (i + j == k) nlo += 2;
      if (i + j + 1 == k) nhi += 2;
    }
  return 0ull - ((uint64_t)nlo * C_LO + (uint64_t)nhi * C_HI);
}

struct El {
  double d[L]; // integers in [0, 2^52), exactly representable
};

__device__ __forceinline__ double to_double52(uint64_t limb) { return __longlong_as_double((long long)(limb | C_LO)) - 4503599627370496.0; }

__device__ __forceinline__ void mad52(uint64_t& col_lo, uint64_t& col_hi, double a, double b)
{
  const double c1 = 20282409603651670423947251286016.0;          // 2^104
  const double c2 = 20282409603651670423947251286016.0 + 4503599627370496.0; // 2^104 + 2^52
  const double hi = __fma_rz(a, b, c1);
  const double lo = __fma_rz(a, b, c2 - hi);
  col_hi += (uint64_t)__double_as_longlong(hi);
  col_lo += (uint64_t)__double_as_longlong(lo);
}

// r = a * b * 2^-260 mod p, r < 2p for a, b < 8p
__device__ __forceinline__ El mul(const El& a, const El& b)
{
  uint64_t t[2 * L + 1];
#pragma unroll
  for (int k = 0; k <= 2 * L; k++) t[k] = col_init(k);
#pragma unroll
  for (int i = 0; i < L; i++)
#pragma unroll
    for (int j = 0; j < L; j++) mad52(t[i + j], t[i + j + 1], a.d[i], b.d[j]);
#pragma unroll
  for (int i = 0; i < L; i++) {
    const uint64_t q = (t[i] * NP52) & M52;
    const double qd = to_double52(q);
#pragma unroll
    for (int j = 0; j < L; j++) mad52(t[i + j], t[i + j + 1], qd, P52[j]);
    t[i + 1] += t[i] >> 52;
  }
  El r;
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < L; k++) {
    const uint64_t v = t[L + k] + c;
    r.d[k] = to_double52(v & M52);
    c = v >> 52;
  }
  return r;
}
} // namespace f52

typedef Fp<params::bn254_fq> Fq;

__global__ void __launch_bounds__(256) k_imad(uint32_t* out, int iters)
{
  Fq a, b;
  for (int i = 0; i < Fq::N; i++) { a.v[i] = threadIdx.x * 77 + i; b.v[i] = blockIdx.x * 13 + i * 5 + 1; }
  a.v[Fq::N - 1] &= 0x0fffffff; b.v[Fq::N - 1] &= 0x0fffffff;
  for (int i = 0; i < iters; i++) { a = a * b; b = b * a; }
  uint32_t r = 0;
  for (int i = 0; i < Fq::N; i++) r ^= a.v[i] ^ b.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

__device__ __forceinline__ void dfma_chain(uint32_t* out, int iters)
{
  f52::El a, b;
  for (int i = 0; i < f52::L; i++) { a.d[i] = (double)(threadIdx.x * 77 + i + 1); b.d[i] = (double)(blockIdx.x * 13 + i * 5 + 1); }
  for (int i = 0; i < iters; i++) { a = f52::mul(a, b); b = f52::mul(b, a); }
  double r = 0;
  for (int i = 0; i < f52::L; i++) r += a.d[i] + b.d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double_as_longlong(r);
}
__global__ void __launch_bounds__(256) k_dfma(uint32_t* out, int iters) { dfma_chain(out, iters); }

// both flavours in one grid: blocks with (blockIdx.x % den) < num run the FP64 flavour
__global__ void __launch_bounds__(256) k_mixed(uint32_t* out, int iters_imad, int iters_dfma, int num, int den)
{
  if ((int)(blockIdx.x % den) < num) {
    dfma_chain(out, iters_dfma);
  } else {
    Fq a, b;
    for (int i = 0; i < Fq::N; i++) { a.v[i] = threadIdx.x * 77 + i; b.v[i] = blockIdx.x * 13 + i * 5 + 1; }
    a.v[Fq::N - 1] &= 0x0fffffff; b.v[Fq::N - 1] &= 0x0fffffff;
    for (int i = 0; i < iters_imad; i++) { a = a * b; b = b * a; }
    uint32_t r = 0;
    for (int i = 0; i < Fq::N; i++) r ^= a.v[i] ^ b.v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  }
}

// validation samples: a, b as 5 limbs each (uint64), result 5 limbs
__global__ void k_check(const uint64_t* in, uint64_t* out, int n)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  f52::El a, b;
  for (int i = 0; i < 5; i++) { a.d[i] = f52::to_double52(in[t * 10 + i]); b.d[i] = f52::to_double52(in[t * 10 + 5 + i]); }
  f52::El r = f52::mul(a, b);
  f52::El r2 = f52::mul(r, r); // a second level: inputs < 2p
  for (int i = 0; i < 5; i++) {
    out[t * 10 + i] = (uint64_t)r.d[i];
    out[t * 10 + 5 + i] = (uint64_t)r2.d[i];
  }
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
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  uint32_t* out;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  const int iters = 400;
  const double prods = 2.0 * iters * (double)blocks * threads;
  float t_i = time_ms([&] { k_imad<<<blocks, threads>>>(out, iters); });
  float t_d = time_ms([&] { k_dfma<<<blocks, threads>>>(out, iters); });
  printf("IMAD.WIDE flavour (ff.cuh, 8x32)   : %8.2f G products/s  (%.3f ms)\n", prods / t_i * 1e-6, t_i);
  printf("FP64 flavour (5x52, fma_rz hi/lo)  : %8.2f G products/s  (%.3f ms)\n", prods / t_d * 1e-6, t_d);
  // mixed: choose per-flavour iteration counts so that both halves take about the same time when alone
  for (int num = 1; num <= 3; num++) {
    const int den = 4;
    // blocks of each flavour: dfma = num/den, imad = rest; give each flavour iterations in proportion to its solo speed
    const double sp_i = prods / t_i, sp_d = prods / t_d;
    const int it_d = iters, it_i = (int)(iters * (sp_i / sp_d) * ((double)num / (den - num)) + 0.5);
    if (it_i <= 0) continue;
    float t_m = time_ms([&] { k_mixed<<<blocks, threads>>>(out, it_i, it_d, num, den); });
    const double pr = 2.0 * threads * ((double)blocks * num / den * it_d + (double)blocks * (den - num) / den * it_i);
    printf("mixed %d/%d blocks FP64 (iters %d) + IMAD (iters %d): %8.2f G products/s  (%.3f ms)\n", num, den, it_d, it_i, pr / t_m * 1e-6, t_m);
  }
  // validation samples
  const int n = 1 << 16;
  std::vector<uint64_t> h_in((size_t)n * 10), h_out((size_t)n * 10);
  uint64_t x = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  const uint64_t top = 0x30644e72e131ull; // top limb of p: keep operands < p by keeping the top limb below it
  for (int t = 0; t < n; t++)
    for (int o = 0; o < 2; o++) {
      for (int i = 0; i < 4; i++) h_in[(size_t)t * 10 + o * 5 + i] = rnd() & f52::M52;
      h_in[(size_t)t * 10 + o * 5 + 4] = rnd() % top;
    }
  // a few edge cases: 0, p-1
  const uint64_t pm1[5] = {0x8c16d87cfd46ull, 0x916871ca8d3c2ull, 0x181585d97816aull, 0xa029b85045b68ull, 0x30644e72e131ull};
  for (int i = 0; i < 5; i++) { h_in[i] = 0; h_in[5 + i] = pm1[i]; h_in[10 + i] = pm1[i]; h_in[15 + i] = pm1[i]; }
  uint64_t *d_in, *d_out;
  cudaMalloc(&d_in, h_in.size() * 8);
  cudaMalloc(&d_out, h_out.size() * 8);
  cudaMemcpy(d_in, h_in.data(), h_in.size() * 8, cudaMemcpyHostToDevice);
  k_check<<<(n + 127) / 128, 128>>>(d_in, d_out, n);
  cudaMemcpy(h_out.data(), d_out, h_out.size() * 8, cudaMemcpyDeviceToHost);
  FILE* f = fopen("gpurun_out/dfma_check.bin", "wb");
  if (f) {
    fwrite(h_in.data(), 8, h_in.size(), f);
    fwrite(h_out.data(), 8, h_out.size(), f);
    fclose(f);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
