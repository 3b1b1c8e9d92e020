// Bucket grouping of the MSM entry list without a general-purpose sort (small / medium inputs; see the measurement below).
//
// The reference has no sort at all: its workers add each point straight into bucket[digit] (icicle/backend/cpu/src/curve/
// cpu_msm.hpp:296-304).  The GPU schedule needs the (scalar, window) entries GROUPED by bucket, in any order inside a bucket, so
// instead of a 3-pass key-value radix sort (42 GB of traffic, bound by the ranking rate of ~150 G keys/s per pass) it does a
// two-kernel counting sort that ranks every entry exactly once:
//   k_digits_rank  scalar -> signed window digits (as k_digits); each non-zero digit takes its RANK inside its bucket with one
//                  returning atomicAdd on the bucket's counter (27 MB of counters for 13 x 2^19 buckets: L2 resident) and stores
//                  {bucket key, rank | sign} window-major (coalesced)
//   k_scan_*       exclusive scan of the counters = the bucket offset table off[] (what the pair levels and the slice
//                  accumulation consume; off[nb] = number of live entries; zero digits are simply dropped)
//   k_scatter      entry e goes to position off[key] + rank: scattered 4-byte stores of the key and of (point index | sign).  Entries
//                  are visited window by window, so the open write frontier is one window's buckets (2^19 x 2 sectors = 32 MB):
//                  it stays in the 126 MB L2 and reaches DRAM as full sectors.
// MEASURED (profiles/r2_msm_counting_sort_experiment.txt): on par with digits + cub up to ~2^22 points, but at 2^26 points the
// scatter of 872 M entries runs at 44 G scattered stores/s (39 ms) and the ranking atomics cost 5 ms: 47 ms against cub's 21 ms.
// It is therefore used only for entry lists below 2^25 (msm_impl.cuh); large inputs keep the library radix sort.
// The order inside a bucket depends on atomic arrival order; the bucket SUM (a group element) does not.
#pragma once
#include "common.cuh"

namespace b200 { namespace msm {

constexpr uint32_t SORT_SIGN_BIT = 0x80000000u;
constexpr uint32_t SORT_DROPPED = 0xffffffffu;

struct MsmPlan;

// ---- exclusive scan over uint32 counters (n up to 2^32-1 elements; sums must fit 32 bits) -------------------------------------
constexpr int SCAN_BLOCK = 1024; // elements per block (256 threads x 4)
static __global__ void __launch_bounds__(256) k_scan_block_sums(const uint32_t* __restrict__ in, uint64_t n, uint32_t* __restrict__ block_sums)
{
  __shared__ uint32_t red[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (base + j < n) s += in[base + j];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += red[w];
    block_sums[blockIdx.x] = t;
  }
}
// one CTA: exclusive scan of the block sums in place (m <= a few thousand at MSM sizes; loops in chunks of 1024)
static __global__ void __launch_bounds__(1024) k_scan_top(uint32_t* __restrict__ sums, uint32_t m)
{
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < m; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = (i < m) ? sums[i] : 0u;
    uint32_t x = v;
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)(threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = wsum[threadIdx.x];
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
        if ((int)threadIdx.x >= o) w += y;
      }
      wsum[threadIdx.x] = w; // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_excl = (threadIdx.x >> 5) ? wsum[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t incl = x + warp_excl;
    if (i < m) sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
}
static __global__ void __launch_bounds__(256) k_scan_apply(const uint32_t* __restrict__ in, uint64_t n, const uint32_t* __restrict__ block_excl,
                                                           uint32_t* __restrict__ out)
{
  __shared__ uint32_t wsum[8];
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x * 4;
  uint32_t v[4];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0u;
    s += v[j];
  }
  uint32_t x = s;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if ((int)(threadIdx.x & 31) >= o) x += y;
  }
  if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < (int)(threadIdx.x >> 5); w++) woff += wsum[w];
  uint32_t run = block_excl[blockIdx.x] + woff + x - s;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
}

struct ScanScratch {
  Scratch sums;
  uint32_t blocks = 0;
  int prepare(uint64_t n_max, cudaStream_t s)
  {
    blocks = (uint32_t)((n_max + SCAN_BLOCK - 1) / SCAN_BLOCK);
    return sums.alloc((size_t)std::max<uint32_t>(blocks, 1) * 4, s);
  }
};
// out[i] = sum_{j < i} in[j], i in [0, n); in and out may alias.  n <= the n_max given to prepare().
static inline int exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint64_t n, ScanScratch& sc, cudaStream_t s)
{
  const uint32_t blocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  if (blocks == 0) return B200_SUCCESS;
  k_scan_block_sums<<<blocks, 256, 0, s>>>(in, n, sc.sums.as<uint32_t>()); B200_LAUNCHED(1);
  k_scan_top<<<1, 1024, 0, s>>>(sc.sums.as<uint32_t>(), blocks); B200_LAUNCHED(1);
  k_scan_apply<<<blocks, 256, 0, s>>>(in, n, sc.sums.as<uint32_t>(), out); B200_LAUNCHED(1);
  B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  return B200_SUCCESS;
}

// entry e (window-major, see k_digits) -> sorted position off[key] + rank
static __global__ void __launch_bounds__(256) k_scatter(const uint32_t* __restrict__ keys0, const uint32_t* __restrict__ rank0, uint64_t n_ent,
                                                        uint32_t n, uint32_t nwin, uint32_t nbm, uint32_t pf, bool shared_points,
                                                        const uint32_t* __restrict__ off, uint32_t* __restrict__ keys1, uint32_t* __restrict__ vals1)
{
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_ent) return;
  const uint32_t key = keys0[e];
  if (key == SORT_DROPPED) return;
  const uint32_t r = rank0[e];
  // e = (b * nwin + w) * n + i  ->  the point the digit multiplies: bases[(shared ? 0 : b*n*pf) + i*pf + w / nbm]
  const uint32_t i = (uint32_t)(e % n);
  const uint32_t bw = (uint32_t)(e / n);
  const uint32_t b = bw / nwin, w = bw % nwin;
  const uint32_t val = ((shared_points ? 0u : b * n * pf) + i * pf + w / nbm) | (r & SORT_SIGN_BIT);
  const uint32_t pos = off[key] + (r & ~SORT_SIGN_BIT);
  keys1[pos] = key;
  vals1[pos] = val;
}

}} // namespace b200::msm
