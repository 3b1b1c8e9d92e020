// Developer probe (GPU box): DRAM cost of random 32-byte / 64-byte gathers from a table much larger than L2, with the
// different PTX cache / prefetch-size qualifiers.  Motivation: ncu shows 128 B of DRAM reads per gathered 64-byte affine
// point (and per 32-byte x coordinate) in the MSM kernels (profiles/r1_ncu_launches_msm_2p24_pair_levels.txt).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu && ./gather_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__device__ __forceinline__ uint4 ldg(const uint4* p)
{
  uint4 r;
  if (MODE == 0) r = *p;
  else if (MODE == 1) asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == 2) asm volatile("ld.global.nc.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  else if (MODE == 6) asm volatile("ld.global.L1::evict_first.L2::64B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// 256-bit loads (sm_100+: LDG.E.ENL2.256): one request per 32-byte sector instead of two
__device__ __forceinline__ uint32_t ldg256(const uint4* p)
{
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(p));
  return r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
}

// each thread gathers `per` random records of BYTES bytes (16-byte pieces) from a table of n_rec 64-byte records
template <int MODE, int BYTES>
__global__ void k_gather(const uint4* __restrict__ table, uint32_t rec_mask, uint32_t per, uint32_t* out)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t x = t * 2654435761u + 12345u;
  uint32_t acc = 0;
  for (uint32_t i = 0; i < per; i++) {
    x = x * 1664525u + 1013904223u;
    const uint32_t rec = (x >> 4) & rec_mask;
    const uint4* p = table + (uint64_t)rec * 4; // 64-byte records
    if (MODE == 7) {
#pragma unroll
      for (int k = 0; k < BYTES / 32; k++) acc += ldg256(p + 2 * k);
    } else {
#pragma unroll
      for (int k = 0; k < BYTES / 16; k++) {
        uint4 v = ldg<MODE>(p + k);
        acc += v.x ^ v.y ^ v.z ^ v.w;
      }
    }
  }
  if (acc == 0x12345678u) out[t] = acc;
}

template <int MODE, int BYTES>
void run(const char* name, const uint4* table, uint32_t rec_mask, uint32_t* out)
{
  const uint32_t per = 64, blocks = 148 * 64, threads = 256;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_gather<MODE, BYTES><<<blocks, threads>>>(table, rec_mask, per, out);
  cudaEventRecord(a);
  for (int r = 0; r < 3; r++) k_gather<MODE, BYTES><<<blocks, threads>>>(table, rec_mask, per, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 3;
  const double gathers = (double)blocks * threads * per;
  printf("%-44s %d B/gather: %8.3f ms  %7.2f G gathers/s  useful %7.1f GB/s  (x128B = %7.1f GB/s, x64B = %7.1f GB/s)\n", name, BYTES, ms,
         gathers / ms * 1e-6, gathers * BYTES / ms * 1e-6, gathers * 128 / ms * 1e-6, gathers * 64 / ms * 1e-6);
}

int main()
{
  const uint64_t n_rec = 1ull << 26; // 4 GiB of 64-byte records
  uint4* table;
  uint32_t* out;
  cudaMalloc(&table, n_rec * 64);
  cudaMalloc(&out, 148 * 64 * 256 * 4);
  cudaMemset(table, 1, n_rec * 64);
  const uint32_t mask = (uint32_t)(n_rec - 1);
#define RUN(M, NAME)                     \
  run<M, 32>(NAME, table, mask, out);    \
  run<M, 64>(NAME, table, mask, out);
  RUN(0, "plain ld.global")
  RUN(1, "ld.global.nc")
  RUN(2, "ld.global.nc.L2::64B")
  RUN(3, "ld.global.nc.L2::128B")
  RUN(4, "ld.global.nc.L1::no_allocate.L2::64B")
  RUN(5, "ld.global.cs")
  RUN(6, "ld.global.L1::evict_first.L2::64B")
  RUN(7, "ld.global.nc.v8.u32 (256-bit)")
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
