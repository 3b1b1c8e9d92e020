// Multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Replaces (reference, CPU): cpu_msm / Msm::run_msm (icicle/backend/cpu/src/curve/cpu_msm.hpp:430-442, 61-75):
//   phase 1  worker_run_phase1           cpu_msm.hpp:259-314   signed-digit windows + bucket accumulation
//   phase 2  worker_collapse_segment     cpu_msm.hpp:317-362   running-sum ("line"/"triangle") bucket reduction
//   phase 3  phase3_final_accumulator    cpu_msm.hpp:365-417   Horner over bucket modules with c doublings
//   and cpu_msm_precompute_bases         cpu_msm.hpp:454-481
// Semantics kept: result[b] = sum_i s[b][i] * P[i]; scalars taken mod 2^bitsize when bitsize != 0; optional Montgomery
// inputs; affine zero (0,0) bases skipped (cpu_msm.hpp:282); batch with shared / per-MSM bases (cpu_msm.hpp:436-437);
// precompute_factor bases laid out as out[pf*i + j] = 2^(j*shift) * P_i (cpu_msm.hpp:468-478, shift derived from our c).
// Output is the reference's homogeneous projective point in standard form; the reference compares group elements,
// not representatives (projective.h:228-231).
//
// GPU schedule (one pass over the scalars, one gather pass over the points):
//   K6  k_digits        scalar -> signed c-bit digits; emits (bucket key, point index|sign) per (scalar, window)
//   K7  bucket grouping (cub::DeviceRadixSort above 2^25 entries; own counting sort, msm_sort.cuh, below) groups entries by bucket
//   K7' pair levels     (msm_pairs.cuh, large inputs) the sorted list is halved up to five times by batched-affine adds
//                       (6 products per add, inversions shared across the grid) before the bucket accumulation
//   K8  k_accumulate    each thread owns a fixed SLICE of the sorted entry list and sums runs of equal keys with
//                       mixed XYZZ adds (8M+2S, Montgomery chains in registers); complete runs go straight to the
//                       bucket, runs cut by a slice boundary become partials -> perfectly balanced for any scalar
//                       distribution (no large-bucket special case)
//       k_resolve       stitches the partials of runs that span slices
//   K9  k_bucket_chunks per-chunk running sums: sum_k (k+1)*B_k = triangle + offset*line
//       k_sum_groups    tree-sum of chunk results per bucket module
//   K10 k_final         Horner over bucket modules (c doublings each), XYZZ -> projective, out of Montgomery form
// Host-pointer calls of >= 2^23 points run the same stages chunk by chunk behind the H2D copies (msm_chunked): every chunk
// accumulates into one shared bucket array (window size of the whole MSM), the reduction runs once.
// The bucket accumulation is integer-multiply bound (6-10 Montgomery products of 140 IMAD.WIDE per point-window) except
// level 0 of the pair tree, which is bound by HBM's random-sector rate; algorithmic HBM bytes are only |scalar| + |affine|
// per point (96 B for BN254 G1).
#pragma once
#include "common.cuh"
#include "msm_pairs.cuh"
#include "msm_sort.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <memory>
#include <thread>
#include <vector>

namespace b200 { namespace msm {

constexpr int MSM_THREADS = 128;
constexpr uint32_t SIGN_BIT = 0x80000000u;

template <class F>
struct alignas(16) AffineRaw {
  uint32_t w[2 * F::N];
};

template <class F>
B200_D Affine<F> load_affine(const uint32_t* p, bool wide = false)
{
  return {load_el_gather<F>(p, wide), load_el_gather<F>(p + F::N, wide)};
}
template <class F>
B200_D XYZZ<F> load_xyzz(const uint32_t* p)
{
  return {load_el<F>(p), load_el<F>(p + F::N), load_el<F>(p + 2 * F::N), load_el<F>(p + 3 * F::N)};
}
template <class F>
B200_D void store_xyzz(uint32_t* p, const XYZZ<F>& a)
{
  store_el(p, a.x);
  store_el(p + F::N, a.y);
  store_el(p + 2 * F::N, a.zz);
  store_el(p + 3 * F::N, a.zzz);
}

struct MsmPlan {
  int c;              // window bits
  int bits;           // scalar bits considered
  int nwin;           // number of c-bit windows covering bits+1
  int pf;             // precompute factor
  int nbm;            // bucket modules = ceil(nwin / pf)
  uint32_t bm_buckets; // buckets per module = 2^(c-1)
};

// ---------------------------------------------------------------------------------------------------------------------
// K6: digits.  Thread per scalar.  keys/vals are laid out [batch_local][window][i] so writes are coalesced.
// key = (batch_local*nbm + bm) * 2^(c-1) + (|d|-1);  d == 0 -> sentinel (total_buckets).  val = point index | sign.
// ---------------------------------------------------------------------------------------------------------------------
template <class S>
__global__ void __launch_bounds__(256) k_digits(
  const uint32_t* __restrict__ scalars, uint32_t n, uint32_t batch_local, MsmPlan pl, bool scalars_mont, bool shared_points,
  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, uint32_t sentinel)
{
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)n * batch_local) return;
  const uint32_t b = (uint32_t)(g / n), i = (uint32_t)(g % n);
  S s = load_fp<S>(scalars + g * S::N);
  if (scalars_mont) s = s.from_mont();
  const uint32_t c = pl.c, half = 1u << (c - 1), full = 1u << c;
  uint32_t carry = 0;
  const uint64_t ent_base = (uint64_t)b * pl.nwin * n;
  const uint32_t pt_base = shared_points ? 0u : b * n * pl.pf;
  for (int w = 0; w < pl.nwin; w++) {
    // raw = bits [w*c, w*c + c) of s, clipped to pl.bits
    const uint32_t lsb = w * c;
    uint32_t raw = 0;
    if (lsb < (uint32_t)pl.bits) {
      const uint32_t limb = lsb >> 5, off = lsb & 31;
      uint64_t two = s.v[limb];
      if (limb + 1 < S::N) two |= (uint64_t)s.v[limb + 1] << 32;
      raw = (uint32_t)(two >> off) & (full - 1);
      const uint32_t avail = pl.bits - lsb;
      if (avail < c) raw &= (1u << avail) - 1;
    }
    raw += carry;
    uint32_t mag, neg;
    if (raw > half) {
      mag = full - raw;
      neg = 1;
      carry = 1;
    } else {
      mag = raw;
      neg = 0;
      carry = 0;
    }
    const uint32_t j = w / pl.nbm, bm = w % pl.nbm;
    uint32_t key = sentinel;
    if (mag) key = (b * pl.nbm + bm) * half + (mag - 1);
    const uint64_t e = ent_base + (uint64_t)w * n + i;
    keys[e] = key;
    vals[e] = (pt_base + i * pl.pf + j) | (neg ? SIGN_BIT : 0u);
  }
}

// K6 + ranking (msm_sort.cuh): as k_digits, but every non-zero digit also takes its rank inside its bucket with one returning
// atomicAdd on the bucket counter; zero digits are marked dropped.  rank word = rank | sign bit of the digit.
template <class S>
__global__ void __launch_bounds__(256) k_digits_rank(
  const uint32_t* __restrict__ scalars, uint32_t n, uint32_t batch_local, MsmPlan pl, bool scalars_mont, uint32_t* __restrict__ keys,
  uint32_t* __restrict__ ranks, uint32_t* __restrict__ cnt)
{
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (uint64_t)n * batch_local) return;
  const uint32_t b = (uint32_t)(g / n), i = (uint32_t)(g % n);
  S s = load_fp<S>(scalars + g * S::N);
  if (scalars_mont) s = s.from_mont();
  const uint32_t c = pl.c, half = 1u << (c - 1), full = 1u << c;
  uint32_t carry = 0;
  const uint64_t ent_base = (uint64_t)b * pl.nwin * n;
  for (int w = 0; w < pl.nwin; w++) {
    const uint32_t lsb = w * c;
    uint32_t raw = 0;
    if (lsb < (uint32_t)pl.bits) {
      const uint32_t limb = lsb >> 5, off = lsb & 31;
      uint64_t two = s.v[limb];
      if (limb + 1 < S::N) two |= (uint64_t)s.v[limb + 1] << 32;
      raw = (uint32_t)(two >> off) & (full - 1);
      const uint32_t avail = pl.bits - lsb;
      if (avail < c) raw &= (1u << avail) - 1;
    }
    raw += carry;
    uint32_t mag, neg;
    if (raw > half) {
      mag = full - raw;
      neg = SORT_SIGN_BIT;
      carry = 1;
    } else {
      mag = raw;
      neg = 0;
      carry = 0;
    }
    const uint32_t bm = w % pl.nbm;
    uint32_t key = SORT_DROPPED, rk = 0;
    if (mag) {
      key = (b * pl.nbm + bm) * half + (mag - 1);
      rk = atomicAdd(cnt + key, 1u) | neg;
    }
    const uint64_t e = ent_base + (uint64_t)w * n + i;
    keys[e] = key;
    ranks[e] = rk;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// K8: slice accumulation over the sorted entries.
// partial slot layout: pkey[2*t + {0,1}] (0xffffffff = empty), pflag bit0 = run starts in this slice, bit1 = run ends here.
// ---------------------------------------------------------------------------------------------------------------------
constexpr uint32_t P_EMPTY = 0xffffffffu;

// DIRECT = false: entries are the radix-sorted (key, point index | sign) pairs, points gathered from `points`.
// DIRECT = true : entries are the affine points of a pair-tree level (msm_pairs.cuh): point e = points[e], no sign, no
//                 sentinel, and the live entry count is read from device memory (*n_dev) because it is only known there.
template <class F, bool DIRECT>
__global__ void __launch_bounds__(MSM_THREADS) k_accumulate(
  const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n_entries, const uint32_t* __restrict__ n_dev, uint32_t slice,
  uint32_t sentinel, const uint32_t* __restrict__ points, uint32_t* __restrict__ buckets, uint32_t* __restrict__ pkey,
  uint32_t* __restrict__ pflag, uint32_t* __restrict__ ppt, uint64_t n_slices, bool wide)
{
  constexpr int AW = 2 * F::N, XW = 4 * F::N;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_slices) return;
  if (n_dev) n_entries = *n_dev; // the live entry count is only known on the device (pair levels; counting sort: off[nb])
  const uint64_t beg = t * slice;
  const uint64_t end = (beg + slice < n_entries) ? beg + slice : n_entries;
  pkey[2 * t] = P_EMPTY;
  pkey[2 * t + 1] = P_EMPTY;
  if (beg >= n_entries) return;

  uint32_t cur = keys[beg];
  if (cur == sentinel) return;
  bool run_starts_here = (beg == 0) || (keys[beg - 1] != cur);
  int slot = 0; // head partial goes to slot 0, tail partial to slot 1
  XYZZ<F> acc = XYZZ<F>::inf();

  uint32_t v = DIRECT ? (uint32_t)beg : vals[beg];
  Affine<F> nxt = load_affine<F>(points + (uint64_t)(v & ~SIGN_BIT) * AW, wide);
  bool nxt_neg = !DIRECT && (v & SIGN_BIT) != 0;

  for (uint64_t e = beg; e < end; e++) {
    Affine<F> p = nxt;
    const bool neg = nxt_neg;
    uint32_t knext = sentinel;
    if (e + 1 < n_entries) knext = keys[e + 1];
    if (e + 1 < end && knext != sentinel) {
      uint32_t v2 = DIRECT ? (uint32_t)(e + 1) : vals[e + 1];
      nxt = load_affine<F>(points + (uint64_t)(v2 & ~SIGN_BIT) * AW, wide);
      nxt_neg = !DIRECT && (v2 & SIGN_BIT) != 0;
    }
    if (neg) p.y = p.y.neg();
    acc.add_affine(p);

    const bool last_in_slice = (e + 1 == end);
    if (knext != cur || last_in_slice) {
      const bool run_ends_here = (knext != cur);
      if (run_starts_here && run_ends_here) {
        store_xyzz<F>(buckets + (uint64_t)cur * XW, acc); // run fully inside this slice: sole owner of the bucket
      } else {
        const uint64_t ps = 2 * t + slot;
        pkey[ps] = cur;
        pflag[ps] = (run_starts_here ? 1u : 0u) | (run_ends_here ? 2u : 0u);
        store_xyzz<F>(ppt + ps * XW, acc);
      }
      if (last_in_slice || knext == sentinel) return;
      cur = knext;
      run_starts_here = true;
      slot = 1;
      acc = XYZZ<F>::inf();
    }
  }
}

// One level of the hierarchical stitch: each thread owns `seg` consecutive partial slots, sums the slots of each run
// (runs are delimited by the start/end flags the previous level wrote), stores runs that both start and end inside its
// segment straight to their bucket and emits at most two next-level partials (a run ending here that started earlier,
// and a run still open at the end of the segment).  The slot list shrinks by seg/2 per level, so a bucket that spans
// thousands of slices (top window, skewed scalars, bitsize = 1) is reduced in O(log) depth instead of by one thread.
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_reduce_partials(
  const uint32_t* __restrict__ in_key, const uint32_t* __restrict__ in_flag, const uint32_t* __restrict__ in_pt, uint64_t n_in,
  uint32_t seg, uint32_t* __restrict__ out_key, uint32_t* __restrict__ out_flag, uint32_t* __restrict__ out_pt, uint64_t n_threads,
  uint32_t* __restrict__ buckets)
{
  constexpr int XW = 4 * F::N;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_threads) return;
  out_key[2 * t] = P_EMPTY;
  out_key[2 * t + 1] = P_EMPTY;
  const uint64_t beg = t * seg;
  const uint64_t end = (beg + seg < n_in) ? beg + seg : n_in;
  bool active = false, started = false;
  uint32_t cur = 0;
  int slot = 0;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint64_t j = beg; j < end; j++) {
    const uint32_t key = in_key[j];
    if (key == P_EMPTY) continue;
    const uint32_t fl = in_flag[j];
    XYZZ<F> o = load_xyzz<F>(in_pt + j * XW);
    if (!active) {
      acc = o;
      cur = key;
      started = (fl & 1u) != 0;
      active = true;
    } else {
      acc.add(o);
    }
    if (fl & 2u) { // run ends at this slot
      if (started) {
        store_xyzz<F>(buckets + (uint64_t)cur * XW, acc);
      } else {
        const uint64_t ps = 2 * t + slot;
        out_key[ps] = cur;
        out_flag[ps] = 2u;
        store_xyzz<F>(out_pt + ps * XW, acc);
      }
      slot = 1;
      active = false;
    }
  }
  if (active) { // run continues into the next segment
    const uint64_t ps = 2 * t + slot;
    out_key[ps] = cur;
    out_flag[ps] = started ? 1u : 0u;
    store_xyzz<F>(out_pt + ps * XW, acc);
  }
}

// Stitch runs that span several slices: the slot that starts a run walks forward until the slot that ends it.
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_resolve(
  const uint32_t* __restrict__ pkey, const uint32_t* __restrict__ pflag, const uint32_t* __restrict__ ppt, uint64_t n_slots,
  uint32_t* __restrict__ buckets)
{
  constexpr int XW = 4 * F::N;
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_slots) return;
  const uint32_t key = pkey[j];
  if (key == P_EMPTY) return;
  const uint32_t fl = pflag[j];
  if (!(fl & 1u)) return; // not the start of a run
  XYZZ<F> acc = load_xyzz<F>(ppt + j * XW);
  if (!(fl & 2u)) {
    for (uint64_t k = j + 1; k < n_slots; k++) {
      if (pkey[k] == P_EMPTY) continue;
      XYZZ<F> o = load_xyzz<F>(ppt + k * XW);
      acc.add(o);
      if (pflag[k] & 2u) break;
    }
  }
  store_xyzz<F>(buckets + (uint64_t)key * XW, acc);
}

// ---------------------------------------------------------------------------------------------------------------------
// K9: bucket reduction.  Module m has buckets B_0..B_{nb-1} with weights 1..nb.  Thread (m, g) handles the chunk
// [g*L, (g+1)*L): tri = sum (k - g*L + 1) B_k, line = sum B_k (descending running sums, cpu_msm.hpp:338-351), and
// emits tri + (g*L) * line.
// ---------------------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_bucket_chunks(
  const uint32_t* __restrict__ buckets, uint32_t n_modules, uint32_t nb_log, uint32_t chunk_log, uint32_t* __restrict__ out)
{
  constexpr int XW = 4 * F::N;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t chunks_per_module_log = nb_log - chunk_log;
  if (t >= ((uint64_t)n_modules << chunks_per_module_log)) return;
  const uint32_t g = (uint32_t)(t & ((1ull << chunks_per_module_log) - 1));
  const uint64_t first = t << chunk_log; // global bucket index of the chunk start (modules are contiguous)
  const uint32_t L = 1u << chunk_log;
  XYZZ<F> line = XYZZ<F>::inf(), tri = XYZZ<F>::inf();
  for (int k = (int)L - 1; k >= 0; k--) {
    XYZZ<F> b = load_xyzz<F>(buckets + (first + k) * XW);
    line.add(b);
    tri.add(line);
  }
  // tri += (g*L) * line
  uint32_t off = g << chunk_log;
  if (off) {
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int bit = 31 - __clz(off); bit >= 0; bit--) {
      acc = acc.dbl();
      if ((off >> bit) & 1) acc.add(line);
    }
    tri.add(acc);
  }
  store_xyzz<F>(out + t * XW, tri);
}

// out[i] = sum_{j < k} in[i*k + j]
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_sum_groups(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n_out, uint32_t k)
{
  constexpr int XW = 4 * F::N;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  XYZZ<F> acc = load_xyzz<F>(in + i * k * XW);
  for (uint32_t j = 1; j < k; j++) {
    XYZZ<F> o = load_xyzz<F>(in + (i * k + j) * XW);
    acc.add(o);
  }
  store_xyzz<F>(out + i * XW, acc);
}

// K10: per batch element: Horner over modules, then XYZZ -> homogeneous projective, out of Montgomery form.
template <class F>
__global__ void k_final(const uint32_t* __restrict__ module_sums, uint32_t nbm, uint32_t c, uint32_t batch, uint32_t* __restrict__ results)
{
  constexpr int XW = 4 * F::N;
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  XYZZ<F> acc = load_xyzz<F>(module_sums + ((uint64_t)b * nbm + (nbm - 1)) * XW);
  for (int m = (int)nbm - 2; m >= 0; m--) {
    for (uint32_t k = 0; k < c; k++) acc = acc.dbl();
    XYZZ<F> o = load_xyzz<F>(module_sums + ((uint64_t)b * nbm + m) * XW);
    acc.add(o);
  }
  Projective<F> pr = acc.to_projective();
  uint32_t* o = results + (uint64_t)b * 3 * F::N;
  store_el(o, pr.x.from_mont());
  store_el(o + F::N, pr.y.from_mont());
  store_el(o + 2 * F::N, pr.z.from_mont());
}

// coordinate-wise to-Montgomery over a flat array of base-field elements (points -> Montgomery form)
template <class B>
__global__ void __launch_bounds__(256) k_to_mont(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    store_fp<B>(out + i * B::N, load_fp<B>(in + i * B::N).to_mont());
}

// precompute: out[pf*i + j] = 2^(j*shift) * in[i] (affine).  Thread per input point.
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_precompute(
  const uint32_t* __restrict__ in, uint32_t n, uint32_t pf, uint32_t shift, bool in_mont, bool out_mont, uint32_t* __restrict__ out)
{
  constexpr int AW = 2 * F::N;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<F> p = load_affine<F>(in + i * AW);
  // copy of the original point, form preserved as given (cpu_msm.hpp:469)
  store_el(out + (i * pf) * AW, p.x);
  store_el(out + (i * pf) * AW + F::N, p.y);
  if (!in_mont) {
    p.x = p.x.to_mont();
    p.y = p.y.to_mont();
  }
  XYZZ<F> q = XYZZ<F>::from_affine(p);
  for (uint32_t j = 1; j < pf; j++) {
    for (uint32_t k = 0; k < shift; k++) q = q.dbl();
    F x = F::zero(), y = F::zero();
    if (!q.is_inf()) {
      F zi = inv_el(q.zzz);        // 1/ZZZ
      F zi2 = F::sqr(zi);          // ZZ^3 = ZZZ^2  =>  1/ZZ = ZZ^2 / ZZZ^2
      x = q.x * F::sqr(q.zz) * zi2; // X/ZZ = X*ZZ^2/ZZZ^2
      y = q.y * zi;                 // Y/ZZZ
    }
    if (!out_mont) {
      x = x.from_mont();
      y = y.from_mont();
    }
    store_el(out + (i * pf + j) * AW, x);
    store_el(out + (i * pf + j) * AW + F::N, y);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
inline int ilog2_ceil(uint64_t x)
{
  int l = 0;
  while ((1ull << l) < x) l++;
  return l;
}

inline int auto_c(int msm_size, int bits, int pf)
{
  // Cost model (EC adds): n * nwin  (accumulation)  +  ~3 * nbm * 2^(c-1)  (reduction incl. offset scalar-muls),
  // full adds weighted 1.4x a mixed add.  Evaluated for c in [4, 24].
  double best = 1e300;
  int best_c = 8;
  for (int c = 4; c <= 24; c++) {
    int nwin = (bits + 1 + c - 1) / c;
    int nbm = (nwin + pf - 1) / pf;
    double acc = (double)msm_size * nwin;
    double red = 1.4 * 3.0 * (double)nbm * (double)(1u << (c - 1));
    double cost = acc + red;
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

template <class C>
MsmPlan make_plan(int msm_size, const b200_msm_config* cfg)
{
  MsmPlan pl;
  typedef typename C::Scalar S;
  pl.bits = (cfg->bitsize > 0 && cfg->bitsize <= S::P::BITS) ? cfg->bitsize : S::P::BITS;
  pl.pf = cfg->precompute_factor > 0 ? cfg->precompute_factor : 1;
  pl.c = cfg->c > 0 ? cfg->c : auto_c(msm_size, pl.bits, pl.pf);
  if (pl.c < 2) pl.c = 2;
  if (pl.c > 24) pl.c = 24;
  pl.nwin = (pl.bits + 1 + pl.c - 1) / pl.c;
  pl.nbm = (pl.nwin + pl.pf - 1) / pl.pf;
  pl.bm_buckets = 1u << (pl.c - 1);
  return pl;
}

// number of MSMs of a batch processed together (entry counts must fit 31 bits and memory stays bounded)
inline int batch_chunk(uint32_t n, const MsmPlan& pl, int batch, bool shared, const b200_msm_config* cfg, size_t coord_bytes = 32)
{
  const uint64_t ent_per_msm = (uint64_t)n * pl.nwin;
  uint64_t max_entries = 1ull << 30;
  if (batch > 1 && ent_per_msm * 2 <= max_entries) {
    // the sorted (key, value) lists (16 B/entry) plus the planar level buffers of the pair tree (~2.25 coordinates' worth per
    // entry) must fit comfortably: wide curves (48- / 96-byte coordinates) get smaller batch chunks instead of losing the levels
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const uint64_t per_entry = 16 + (uint64_t)(2.25 * (double)coord_bytes * 2);
      max_entries = std::max<uint64_t>(ent_per_msm, std::min<uint64_t>(max_entries, (uint64_t)(0.45 * (double)free_b) / per_entry));
    } else {
      (void)cudaGetLastError();
    }
  }
  int chunk = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)batch, max_entries / std::max<uint64_t>(ent_per_msm, 1)));
  if (cfg->ext_nof_chunks > 0) chunk = std::min(chunk, std::max(1, (batch + cfg->ext_nof_chunks - 1) / cfg->ext_nof_chunks)); // never above the 2^30-entry cap
  if (!shared) { // point indices must fit 31 bits
    while (chunk > 1 && (uint64_t)chunk * n * pl.pf >= (1ull << 31)) chunk--;
  }
  return chunk;
}

// Number of batched-affine pair levels (msm_pairs.cuh) run before the XYZZ accumulation.  L levels halve the sorted list L
// times with ~6-product affine adds instead of 10-product mixed adds.  Measured on B200 (profiles/r1_msm_pair_levels.txt):
// each level costs ~0.7 ms of small inversion-tree kernels on top of its per-pair work, so the levels pay off from ~2^26
// entries, and the deeper levels are worth it while the bucket runs stay >= 8 entries long.
inline int choose_pair_levels(uint64_t max_ent, uint64_t max_buckets)
{
  int levels = 0;
  const double avg_run = (double)max_ent / (double)std::max<uint64_t>(max_buckets, 1);
  if (max_ent >= (1ull << 26) && avg_run >= 8.0) {
    while ((8 << levels) <= avg_run && levels < 5) levels++;
  }
  if (tune(T_MSM_PAIR_LEVELS) >= 0) levels = std::min(8, tune(T_MSM_PAIR_LEVELS));
  return levels;
}

// Device-resident core: scalars (standard or Montgomery form), points already in Montgomery form, results written to d_res
// (device).  Everything is enqueued on `s`; nothing here synchronises.
// `phases`: MSM_ACC = digits .. stitched bucket sums, MSM_RED = bucket reduction + final; the host-pointer pipeline
// (msm_chunked) runs MSM_ACC once per point-range chunk into `ext_bkt` (batch == 1) and MSM_RED once at the end.
enum : int { MSM_ACC = 1, MSM_RED = 2, MSM_ALL = 3 };

template <class C>
int msm_core(const void* d_scal, const uint32_t* pts_m, uint32_t n, const MsmPlan& pl, int batch, bool shared, const b200_msm_config* cfg,
             void* d_res, cudaStream_t s, StageTimer& prof, uint32_t* ext_bkt = nullptr, int phases = MSM_ALL)
{
  typedef typename C::Scalar S;
  typedef typename C::Base F;
  constexpr int AW = 2 * F::N, XW = 4 * F::N, PW = 3 * F::N;
  int err;
  // ---- batch chunking so that entry counts fit 31 bits and memory stays bounded ----------------------------------------
  const uint64_t ent_per_msm = (uint64_t)n * pl.nwin;
  int chunk = batch_chunk(n, pl, batch, shared, cfg, (size_t)F::BYTES);
  if ((uint64_t)n * pl.pf >= (1ull << 31) || ent_per_msm >= (1ull << 32)) return B200_INVALID_ARGUMENT;

  const uint64_t max_ent = ent_per_msm * chunk;
  const uint64_t max_modules = (uint64_t)pl.nbm * chunk;
  const uint64_t max_buckets = max_modules << (pl.c - 1);
  if (max_buckets >= 0xfffffff0ull) return B200_INVALID_ARGUMENT;

  // slice length: aim for >= 4 slices per resident thread, between 8 and 128 entries
  const uint64_t target_threads = (uint64_t)num_sms() * 2048;
  uint32_t slice = (uint32_t)std::min<uint64_t>(128, std::max<uint64_t>(8, max_ent / target_threads));
  const uint64_t max_slices = (max_ent + slice - 1) / slice;

  // chunk length for the bucket reduction: >= 64K threads if possible
  int chunk_target = 16; // log2 of the thread count k_bucket_chunks aims for
  if (tune(T_MSM_CHUNK_TARGET) >= 0) chunk_target = std::max(10, std::min(22, tune(T_MSM_CHUNK_TARGET)));
  int chunk_log = std::max(0, std::min(7, (pl.c - 1 + ilog2_ceil(max_modules)) - chunk_target));
  if (chunk_log > pl.c - 1) chunk_log = pl.c - 1;
  const uint64_t max_chunks = max_buckets >> chunk_log;

  if (ext_bkt && (batch != 1 || chunk != 1)) return B200_INVALID_ARGUMENT;
  const bool do_acc = (phases & MSM_ACC) != 0, do_red = (phases & MSM_RED) != 0;
  Scratch s_k0, s_k1, s_v0, s_v1, s_cub, s_bkt, s_pkey, s_pflag, s_ppt, s_pkey2, s_pflag2, s_ppt2, s_red0, s_red1, s_cnt0, s_off0;
  ScanScratch scan_sc;
  size_t cub_bytes = 0;
  // bucket grouping: cub radix sort for large entry lists, the counting sort of msm_sort.cuh below 2^25 entries (on par / fewer
  // launches there; at 2^26 points its fully scattered stores lose 47 vs 21 ms: profiles/r2_msm_counting_sort_experiment.txt).
  // msm_sort knob: 1 = always cub, 2 = always counting sort.
  const bool use_cub = tune(T_MSM_SORT) == 1 || (tune(T_MSM_SORT) != 2 && max_ent > (1ull << 25));
  if (do_acc) {
    if ((err = s_k0.alloc(max_ent * 4, s))) return err;
    if ((err = s_k1.alloc(max_ent * 4, s))) return err;
    if ((err = s_v0.alloc(max_ent * 4, s))) return err;
    if ((err = s_v1.alloc(max_ent * 4, s))) return err;
    if (use_cub) {
      cub::DoubleBuffer<uint32_t> dk(s_k0.as<uint32_t>(), s_k1.as<uint32_t>()), dv(s_v0.as<uint32_t>(), s_v1.as<uint32_t>());
      cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, (int64_t)max_ent, 0, 32, s);
      if ((err = s_cub.alloc(cub_bytes, s))) return err;
    } else {
      if ((err = s_cnt0.alloc((size_t)(max_buckets + 1) * 4, s))) return err;
      if ((err = s_off0.alloc((size_t)(max_buckets + 1) * 4, s))) return err;
    }
    if ((err = scan_sc.prepare(max_buckets + 1, s))) return err;
    if ((err = s_pkey.alloc(max_slices * 2 * 4, s))) return err;
    if ((err = s_pflag.alloc(max_slices * 2 * 4, s))) return err;
    if ((err = s_ppt.alloc(max_slices * 2 * XW * 4, s))) return err;
    const uint64_t max_slots2 = 2 * ((max_slices * 2 + 15) / 16) + 2;
    if ((err = s_pkey2.alloc(max_slots2 * 4, s))) return err;
    if ((err = s_pflag2.alloc(max_slots2 * 4, s))) return err;
    if ((err = s_ppt2.alloc(max_slots2 * XW * 4, s))) return err;
  }
  if (!ext_bkt && (err = s_bkt.alloc(max_buckets * XW * 4, s))) return err;
  uint32_t* const bkt = ext_bkt ? ext_bkt : s_bkt.as<uint32_t>();
  if (do_red) {
    if ((err = s_red0.alloc(std::max<uint64_t>(max_chunks, 1) * XW * 4, s))) return err;
    if ((err = s_red1.alloc(std::max<uint64_t>(max_chunks / 2 + 1, 1) * XW * 4, s))) return err;
  }

  // ---- batched-affine pair levels (msm_pairs.cuh) ---------------------------------------------------------------------
  int levels = do_acc ? choose_pair_levels(max_ent, max_buckets) : 0;
  constexpr uint32_t PAIR_J = 32, INV_G = 64;
  const uint32_t nb_max = (uint32_t)max_buckets;
  uint64_t cap[10];
  cap[0] = max_ent;
  for (int l = 0; l < levels; l++) cap[l + 1] = cap[l] / 2 + nb_max + 1;
  Scratch s_off[2], s_cnt, s_lvl[2], s_lkey[2], s_pbuf, s_tree, s_tpre;
  uint64_t tree_m[8];
  int tree_levels = 0;
  uint64_t tree_total = 0;
  if (levels > 0) {
    uint64_t cap_max = cap[0]; // sparse lists (few entries, many buckets) have a growing upper bound
    for (int l = 1; l < levels; l++) cap_max = std::max(cap_max, cap[l]);
    uint64_t m = (((cap_max + 1) / 2 + (uint64_t)PAIR_J * PAIR_THREADS - 1) / ((uint64_t)PAIR_J * PAIR_THREADS)) * PAIR_THREADS;
    for (;;) {
      tree_m[tree_levels++] = m;
      tree_total += m;
      if (m <= INV_G) break;
      m = (m + INV_G - 1) / INV_G;
    }
    bool ok = true;
    for (int b = 0; b < 2 && ok; b++) {
      ok = ok && !s_off[b].alloc(((size_t)nb_max + 1) * 4, s);
      if (b < levels) {
        ok = ok && !s_lvl[b].alloc((size_t)cap[b + 1] * AW * 4, s);
        ok = ok && !s_lkey[b].alloc((size_t)cap[b + 1] * 4, s);
      }
    }
    ok = ok && !s_cnt.alloc(((size_t)nb_max + 1) * 4, s) &&
         !s_pbuf.alloc(((size_t)cap[0] / 2 + 1) * F::BYTES, s) && !s_tree.alloc((size_t)tree_total * F::BYTES, s) &&
         !s_tpre.alloc((size_t)tree_total * F::BYTES, s);
    if (!ok) { // not enough memory for the level buffers: fall back to the XYZZ-only schedule
      levels = 0;
      for (int b = 0; b < 2; b++) { s_off[b].release(); s_lvl[b].release(); s_lkey[b].release(); }
      s_cnt.release(); s_pbuf.release(); s_tree.release(); s_tpre.release();
    }
  }
  if (levels > 0) {
    uint32_t sl = (uint32_t)std::min<uint64_t>(128, std::max<uint64_t>(8, cap[levels] / target_threads));
    while ((cap[levels] + sl - 1) / sl > max_slices) sl++; // the partial-slot buffers were sized for max_slices
    slice = sl;
  }

  for (int b0 = 0; b0 < batch; b0 += chunk) {
    const int bl = std::min(chunk, batch - b0);
    const uint64_t n_ent = ent_per_msm * bl;
    const uint32_t n_modules = (uint32_t)pl.nbm * bl;
    const uint64_t n_buckets = (uint64_t)n_modules << (pl.c - 1);
    const uint32_t sentinel = (uint32_t)n_buckets;
    const uint32_t* sc = (const uint32_t*)d_scal + (uint64_t)b0 * n * S::N;
    const uint32_t* pts = pts_m + (shared ? 0 : (uint64_t)b0 * n * pl.pf * AW);
    const bool pts_wide = (((uintptr_t)pts) & 31u) == 0 && (F::BYTES % 32 == 0) && tune(T_MSM_NO_WIDE_LOADS) <= 0;

    if (do_acc) {
    const uint32_t nb = (uint32_t)n_buckets;
    const uint32_t* sorted_k;          // entries grouped by bucket: keys ...
    const uint32_t* sorted_v;          // ... and (point index | sign)
    const uint32_t* n_live = nullptr;  // device word holding the number of live entries (counting sort), or nullptr = n_ent
    uint32_t* off_first = (levels > 0) ? s_off[0].as<uint32_t>() : s_off0.as<uint32_t>(); // off[k] = first entry of bucket k
    uint32_t acc_sentinel = sentinel;
    if (use_cub) {
      // K6 + K7 (round 1): digits with a sentinel key for zero digits, library radix sort
      uint64_t th = (uint64_t)n * bl;
      k_digits<S><<<(unsigned)((th + 255) / 256), 256, 0, s>>>(
        sc, n, (uint32_t)bl, pl, cfg->are_scalars_montgomery_form, shared, s_k0.as<uint32_t>(), s_v0.as<uint32_t>(), sentinel); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
      prof.mark("digits");
      cub::DoubleBuffer<uint32_t> dk(s_k0.as<uint32_t>(), s_k1.as<uint32_t>()), dv(s_v0.as<uint32_t>(), s_v1.as<uint32_t>());
      const int key_bits = std::max(1, ilog2_ceil((uint64_t)sentinel + 1));
      B200_CUDA_TRY(cub::DeviceRadixSort::SortPairs(s_cub.p, cub_bytes, dk, dv, (int64_t)n_ent, 0, key_bits, s), B200_UNKNOWN_ERROR);
      sorted_k = dk.Current();
      sorted_v = dv.Current();
      if (levels > 0) {
        k_bounds<<<(nb + 1 + 255) / 256, 256, 0, s>>>(sorted_k, (uint32_t)n_ent, nb, off_first); B200_LAUNCHED(1);
      }
      prof.mark("sort");
    } else {
      // K6 + K7 (msm_sort.cuh): digits ranked inside their bucket by one atomic each, scan of the counters, scatter
      B200_CUDA_TRY(cudaMemsetAsync(s_cnt0.p, 0, ((size_t)nb + 1) * 4, s), B200_UNKNOWN_ERROR);
      uint64_t th = (uint64_t)n * bl;
      k_digits_rank<S><<<(unsigned)((th + 255) / 256), 256, 0, s>>>(
        sc, n, (uint32_t)bl, pl, cfg->are_scalars_montgomery_form, s_k0.as<uint32_t>(), s_v0.as<uint32_t>(), s_cnt0.as<uint32_t>()); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
      prof.mark("digits");
      if ((err = exclusive_scan_u32(s_cnt0.as<uint32_t>(), off_first, (uint64_t)nb + 1, scan_sc, s))) return err;
      k_scatter<<<(unsigned)((n_ent + 255) / 256), 256, 0, s>>>(s_k0.as<uint32_t>(), s_v0.as<uint32_t>(), n_ent, n, (uint32_t)pl.nwin, (uint32_t)pl.nbm,
                                                               (uint32_t)pl.pf, shared, off_first, s_k1.as<uint32_t>(), s_v1.as<uint32_t>()); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
      sorted_k = s_k1.as<uint32_t>();
      sorted_v = s_v1.as<uint32_t>();
      n_live = off_first + nb;
      acc_sentinel = 0xffffffffu; // zero digits never enter the list
      prof.mark("sort");
    }
    // K8
    B200_CUDA_TRY(cudaMemsetAsync(bkt, 0, n_buckets * XW * 4, s), B200_UNKNOWN_ERROR);
    uint64_t n_slices;
    if (levels == 0) {
      n_slices = (n_ent + slice - 1) / slice;
      k_accumulate<F, false><<<(unsigned)((n_slices + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(
        sorted_k, sorted_v, n_ent, n_live, slice, acc_sentinel, pts, bkt, s_pkey.as<uint32_t>(),
        s_pflag.as<uint32_t>(), s_ppt.as<uint32_t>(), n_slices, pts_wide); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    } else {
      const unsigned gb = (nb + 1 + 255) / 256;
      uint64_t capl = n_ent; // upper bound of the current level's length (the exact length lives in off[nb] on the device)
      const uint32_t *lk = sorted_k, *lp = pts;
      uint64_t lcap = 0;     // plane stride of the current (planar) level buffer
      uint4* pbuf = s_pbuf.as<uint4>();
      const uint64_t pcap = cap[0] / 2 + 1;
      for (int l = 0; l < levels; l++) {
        uint32_t* off_c = s_off[l & 1].as<uint32_t>();
        uint32_t* off_n = s_off[(l + 1) & 1].as<uint32_t>();
        k_pair_counts<<<gb, 256, 0, s>>>(off_c, nb, s_cnt.as<uint32_t>()); B200_LAUNCHED(1);
        if ((err = exclusive_scan_u32(s_cnt.as<uint32_t>(), off_n, (uint64_t)nb + 1, scan_sc, s))) return err;
        const uint64_t nq = (capl + 1) / 2;
        const unsigned gp = (unsigned)((nq + (uint64_t)PAIR_J * PAIR_THREADS - 1) / ((uint64_t)PAIR_J * PAIR_THREADS));
        const uint32_t nthr = gp * PAIR_THREADS;
        uint32_t* tree = s_tree.as<uint32_t>();
        uint32_t* tpre = s_tpre.as<uint32_t>();
        uint32_t* out_p = s_lvl[l & 1].as<uint32_t>();
        uint32_t* out_k = s_lkey[l & 1].as<uint32_t>();
        const uint64_t ocap = cap[(l & 1) + 1];
        const bool last = (l + 1 == levels);
        if (l == 0) {
          PairSrc<F, true> src = {lk, sorted_v, lp, 0, pts_wide};
          k_pair_prefix<F, true><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, nb, PAIR_J, pbuf, pcap, tree);
        } else {
          PairSrc<F, false> src = {lk, nullptr, lp, lcap, false};
          k_pair_prefix<F, false><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, nb, PAIR_J, pbuf, pcap, tree);
        }
        B200_LAUNCHED(1);
        // batch inversion of the nthr per-thread totals (in place in tree[0 .. nthr))
        {
          uint64_t m[8], o[8];
          int R = 0;
          uint64_t mm = nthr, oo = 0;
          for (;;) {
            m[R] = mm; o[R] = oo; R++;
            if (mm <= INV_G) break;
            oo += mm;
            mm = (mm + INV_G - 1) / INV_G;
          }
          for (int r = 0; r + 1 < R; r++) {
            k_inv_up<F><<<(unsigned)((m[r + 1] + 127) / 128), 128, 0, s>>>(
              tree + o[r] * F::N, (uint32_t)m[r], INV_G, tpre + o[r] * F::N, tree + o[r + 1] * F::N, (uint32_t)m[r + 1]); B200_LAUNCHED(1);
          }
          k_inv_top<F><<<1, 32, 0, s>>>(tree + o[R - 1] * F::N, (uint32_t)m[R - 1], tpre + o[R - 1] * F::N); B200_LAUNCHED(1);
          for (int r = R - 2; r >= 0; r--) {
            k_inv_down<F><<<(unsigned)((m[r + 1] + 127) / 128), 128, 0, s>>>(
              tree + o[r] * F::N, (uint32_t)m[r], INV_G, tpre + o[r] * F::N, tree + o[r + 1] * F::N, (uint32_t)m[r + 1]); B200_LAUNCHED(1);
          }
        }
        if (l == 0) {
          PairSrc<F, true> src = {lk, sorted_v, lp, 0, pts_wide};
          if (last) k_pair_apply<F, true, false><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, off_n, nb, PAIR_J, pbuf, pcap, tree, out_p, ocap, out_k);
          else k_pair_apply<F, true, true><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, off_n, nb, PAIR_J, pbuf, pcap, tree, out_p, ocap, out_k);
        } else {
          PairSrc<F, false> src = {lk, nullptr, lp, lcap, false};
          if (last) k_pair_apply<F, false, false><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, off_n, nb, PAIR_J, pbuf, pcap, tree, out_p, ocap, out_k);
          else k_pair_apply<F, false, true><<<gp, PAIR_THREADS, 0, s>>>(src, off_c, off_n, nb, PAIR_J, pbuf, pcap, tree, out_p, ocap, out_k);
        }
        B200_LAUNCHED(1);
        B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
        lk = out_k; lp = out_p; lcap = ocap;
        capl = capl / 2 + nb + 1;
      }
      prof.mark("pair_levels");
      n_slices = (capl + slice - 1) / slice;
      k_accumulate<F, true><<<(unsigned)((n_slices + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(
        lk, nullptr, 0, s_off[levels & 1].as<uint32_t>() + nb, slice, 0xffffffffu, lp, bkt, s_pkey.as<uint32_t>(),
        s_pflag.as<uint32_t>(), s_ppt.as<uint32_t>(), n_slices, true); B200_LAUNCHED(1); // level buffers are our own (aligned) scratch
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    }
    prof.mark("accumulate");
    {
      // hierarchical stitch of the boundary partials (see k_reduce_partials), then a final owner walk over <= 512 slots
      uint32_t *ik = s_pkey.as<uint32_t>(), *ifl = s_pflag.as<uint32_t>(), *ipt = s_ppt.as<uint32_t>();
      uint32_t *ok = s_pkey2.as<uint32_t>(), *ofl = s_pflag2.as<uint32_t>(), *opt = s_ppt2.as<uint32_t>();
      uint64_t n_slots = 2 * n_slices;
      const uint32_t seg = 16;
      while (n_slots > 512) {
        const uint64_t nt = (n_slots + seg - 1) / seg;
        k_reduce_partials<F><<<(unsigned)((nt + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(
          ik, ifl, ipt, n_slots, seg, ok, ofl, opt, nt, bkt); B200_LAUNCHED(1);
        B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
        std::swap(ik, ok);
        std::swap(ifl, ofl);
        std::swap(ipt, opt);
        n_slots = 2 * nt;
      }
      k_resolve<F><<<(unsigned)((n_slots + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(ik, ifl, ipt, n_slots, bkt); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    }
    prof.mark("resolve");
    } // do_acc
    if (!do_red) continue;
    // K9
    const uint64_t n_chunks = n_buckets >> chunk_log;
    k_bucket_chunks<F><<<(unsigned)((n_chunks + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(
      bkt, n_modules, (uint32_t)(pl.c - 1), (uint32_t)chunk_log, s_red0.as<uint32_t>()); B200_LAUNCHED(1);
    B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    uint32_t* cur = s_red0.as<uint32_t>();
    uint32_t* other = s_red1.as<uint32_t>();
    uint64_t per_module = 1ull << (pl.c - 1 - chunk_log);
    while (per_module > 1) {
      const uint32_t k = (uint32_t)std::min<uint64_t>(per_module, 8);
      const uint64_t n_out = (uint64_t)n_modules * (per_module / k);
      k_sum_groups<F><<<(unsigned)((n_out + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(cur, other, n_out, k); B200_LAUNCHED(1);
      B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
      std::swap(cur, other);
      per_module /= k;
    }
    prof.mark("bucket_reduce");
    // K10
    k_final<F><<<(bl + 31) / 32, 32, 0, s>>>(cur, (uint32_t)pl.nbm, (uint32_t)pl.c, (uint32_t)bl, (uint32_t*)d_res + (uint64_t)b0 * PW); B200_LAUNCHED(1);
    B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  }
  return B200_SUCCESS;
}

// dst[k] += src[k] over two XYZZ bucket arrays (host-pointer pipeline: per-chunk bucket sums into the running ones)
template <class F>
__global__ void __launch_bounds__(MSM_THREADS) k_bucket_merge(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n)
{
  constexpr int XW = 4 * F::N;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  XYZZ<F> b = load_xyzz<F>(src + i * XW);
  if (b.is_inf()) return;
  XYZZ<F> a = load_xyzz<F>(dst + i * XW);
  a.add(b);
  store_xyzz<F>(dst + i * XW, a);
}

// ---------------------------------------------------------------------------------------------------------------------
// Chunked MSM: one MSM processed in point-range chunks that all accumulate into ONE bucket array (window size of the whole
// MSM); the bucket reduction + Horner run once at the end.  Two users:
//  * host-pointer calls of >= 2^23 points (the reference-facing call, what `e2e` measures): PCIe moves 96 B/point at ~55 GB/s,
//    so a 2^26 MSM is transfer-bound unless the copies hide behind the arithmetic.  The whole input gets a device staging
//    area; copies run ahead on private copy streams and never wait for the compute stream, which processes chunk i as soon as
//    it has landed.  PINNED sources are copied directly (one cudaMemcpyAsync per chunk and array).  PAGEABLE sources -- what
//    Rust / Go / C++ callers normally pass -- are moved by a few host copier threads through a ring of pinned slots
//    (memcpy user -> slot, cudaMemcpyAsync slot -> device), because cudaMemcpyAsync from pageable memory blocks the calling
//    thread and runs at a fraction of the PCIe rate;
//  * device-resident calls whose entry list would not fit 32-bit indexing / memory (n * windows > 2^30, i.e. > 2^26 points at
//    c = 20): no copies, equal chunks.
// Chunk 0 accumulates straight into the running bucket array, later chunks into a second array that k_bucket_merge folds in.
// ---------------------------------------------------------------------------------------------------------------------
// B200_MSM_PIPELINE_MIN=<points> (tests): smallest host-pointer MSM that takes the pipeline, and the chunks may then be small
inline uint32_t pipeline_min_points()
{
  if (tune(T_MSM_PIPELINE_MIN) >= 0) return (uint32_t)std::max(2, tune(T_MSM_PIPELINE_MIN));
  return 1u << 23;
}

// Chunk schedule of the host-pointer pipeline, as point counts.  The copy stream never waits, so chunk k is on the device at
// (points up to k) / PCIe rate; the compute stream is the slower of the two (~0.44 vs ~0.57 G points/s for BN254 G1), so
// the exposed transfer is the FIRST chunk only and the cost of splitting is the shorter bucket runs per chunk (fewer
// batched-affine levels, one k_bucket_merge per chunk).  Hence: two small chunks to start early, then doubling, then
// quarters -- 1/16, 1/16, 1/8, 1/4, 1/4, 1/4 (measured on B200, profiles/r1_e2e_pipeline.txt).  B200_MSM_PIPELINE_CHUNKS=k
// forces k equal chunks instead.
inline uint32_t pipeline_schedule(uint32_t n, uint32_t* sizes, uint32_t max_chunks)
{
  const bool tiny_ok = tune(T_MSM_PIPELINE_MIN) >= 0; // tests: allow tiny chunks
  const uint32_t min_chunk = tiny_ok ? 1u : (1u << 20);
  uint32_t k = 0, left = n;
  if (tune(T_MSM_PIPELINE_CHUNKS) > 0) {
    const uint32_t want = (uint32_t)std::max(1, std::min((int)max_chunks, tune(T_MSM_PIPELINE_CHUNKS)));
    const uint32_t each = std::max<uint32_t>((n + want - 1) / want, min_chunk);
    while (left > 0 && k < max_chunks) {
      sizes[k] = (k + 1 == max_chunks) ? left : std::min(each, left);
      left -= sizes[k++];
    }
    return k;
  }
  static const uint32_t shift[6] = {4, 4, 3, 2, 2, 2};
  for (int i = 0; i < 6 && left > 0; i++) {
    uint32_t want = std::max<uint32_t>(n >> shift[i], min_chunk);
    if (i == 5 || left < want + min_chunk) want = left; // the last chunk takes the remainder
    sizes[k] = std::min(want, left);
    left -= sizes[k++];
  }
  return k;
}

// largest point count whose entry list (n * windows) stays within 2^30 entries: bigger device-resident MSMs are chunked
inline uint32_t device_chunk_points(const MsmPlan& pl) { return (uint32_t)std::max<uint64_t>(1, (1ull << 30) / (uint64_t)pl.nwin); }

template <class C>
int msm_chunked(const void* scalars, const void* bases, uint32_t n, const b200_msm_config* cfg, void* results, HostKind ks, HostKind kp)
{
  typedef typename C::Scalar S;
  typedef typename C::Base F;
  typedef typename base_fp<F>::type B;
  constexpr int AW = 2 * F::N, XW = 4 * F::N, PW = 3 * F::N;
  constexpr uint32_t MAX_CHUNKS = 32;
  cudaStream_t s = (cudaStream_t)cfg->stream;
  DeviceRes* res = device_res();
  if (!res) return B200_UNKNOWN_ERROR;
  cudaStream_t cs = res->copy_stream;
  b200_msm_config sub = *cfg;
  sub.batch_size = 1;
  const MsmPlan pl = make_plan<C>((int)n, &sub); // one window size for all chunks: they share the bucket array
  const uint32_t pf = (uint32_t)pl.pf;
  const bool any_host = (ks != HK_DEVICE) || (kp != HK_DEVICE);
  uint32_t csize[MAX_CHUNKS], coff[MAX_CHUNKS];
  uint32_t nchunks;
  if (any_host) {
    nchunks = pipeline_schedule(n, csize, MAX_CHUNKS);
  } else {
    const uint32_t cap = device_chunk_points(pl);
    nchunks = std::min<uint32_t>(MAX_CHUNKS, (n + cap - 1) / cap);
    const uint32_t each = (n + nchunks - 1) / nchunks;
    uint32_t left = n;
    for (uint32_t i = 0; i < nchunks; i++) {
      csize[i] = std::min(each, left);
      left -= csize[i];
    }
    if (left) return B200_INVALID_ARGUMENT; // > 32 * 2^30 / windows points
  }
  uint32_t max_chunk = 0;
  for (uint32_t i = 0, o = 0; i < nchunks; o += csize[i], i++) {
    coff[i] = o;
    max_chunk = std::max(max_chunk, csize[i]);
  }
  const uint64_t n_buckets = (uint64_t)pl.nbm << (pl.c - 1);
  int err;
  Scratch d_s, d_p, d_pm, d_bkt, d_tmp, s_res;
  ChunkArray as, ap;
  as.kind = ks; as.elem_bytes = S::BYTES; as.host = (const uint8_t*)scalars; as.dev = (uint8_t*)const_cast<void*>(scalars);
  ap.kind = kp; ap.elem_bytes = (size_t)pf * AW * 4; ap.host = (const uint8_t*)bases; ap.dev = (uint8_t*)const_cast<void*>(bases);
  if (ks != HK_DEVICE) {
    if ((err = d_s.alloc((size_t)n * S::BYTES, s))) return err;
    as.dev = d_s.as<uint8_t>();
  } else if (misaligned16(scalars)) {
    return B200_INVALID_POINTER;
  }
  if (kp != HK_DEVICE) {
    if ((err = d_p.alloc((size_t)n * ap.elem_bytes, s))) return err;
    ap.dev = d_p.as<uint8_t>();
  } else {
    if (misaligned16(bases)) return B200_INVALID_POINTER;
    // the caller's device points stay untouched: each chunk is converted into this scratch
    if (!cfg->are_points_montgomery_form && (err = d_pm.alloc((size_t)max_chunk * ap.elem_bytes, s))) return err;
  }
  if ((err = d_bkt.alloc((size_t)n_buckets * XW * 4, s))) return err;
  if (nchunks > 1 && (err = d_tmp.alloc((size_t)n_buckets * XW * 4, s))) return err;
  void* d_res;
  if ((err = stage_out(d_res, results, (size_t)PW * 4, cfg->are_results_on_device, s, s_res))) return err;

  // ---- copies ------------------------------------------------------------------------------------------------------------
  PageableCopy pageable;
  cudaEvent_t ready = res->event(0);
  const bool direct_copies = (ks == HK_PINNED) || (kp == HK_PINNED);
  const bool ring_copies = (ks == HK_PAGEABLE) || (kp == HK_PAGEABLE);
  if (any_host) {
    cudaEventRecord(ready, s); // staging buffers exist (stream-ordered allocation) and earlier work on s is ordered before the copies
    if (direct_copies) {
      cudaStreamWaitEvent(cs, ready, 0);
      for (uint32_t i = 0; i < nchunks; i++) {
        const size_t off = coff[i], cn = csize[i];
        if (ks == HK_PINNED) cudaMemcpyAsync(as.dev + off * as.elem_bytes, as.host + off * as.elem_bytes, cn * as.elem_bytes, cudaMemcpyHostToDevice, cs);
        if (kp == HK_PINNED) cudaMemcpyAsync(ap.dev + off * ap.elem_bytes, ap.host + off * ap.elem_bytes, cn * ap.elem_bytes, cudaMemcpyHostToDevice, cs);
        cudaEventRecord(res->event(1 + i), cs);
      }
    }
    if (ring_copies) {
      if (ks == HK_PAGEABLE) pageable.add(as, coff, csize, nchunks);
      if (kp == HK_PAGEABLE) pageable.add(ap, coff, csize, nchunks);
      if ((err = pageable.start(res, nchunks, ready))) return err;
    }
  }

  StageTimer prof;
  prof.begin(s);
  int rc = B200_SUCCESS;
  for (uint32_t i = 0; i < nchunks && rc == B200_SUCCESS; i++) {
    const uint32_t off = coff[i], cn = csize[i];
    if (direct_copies) cudaStreamWaitEvent(s, res->event(1 + i), 0);
    if (ring_copies) pageable.wait_chunk(i, s);
    const uint32_t* cp = (const uint32_t*)(ap.dev + (size_t)off * ap.elem_bytes);
    if (!cfg->are_points_montgomery_form) {
      uint32_t* dst = (kp != HK_DEVICE) ? const_cast<uint32_t*>(cp) : d_pm.as<uint32_t>(); // our own staging copy converts in place
      const uint64_t ncoord = (uint64_t)cn * pf * 2 * (F::N / B::N);
      unsigned g = (unsigned)std::min<uint64_t>((ncoord + 255) / 256, (uint64_t)num_sms() * 16);
      k_to_mont<B><<<g, 256, 0, s>>>(cp, dst, ncoord); B200_LAUNCHED(1);
      cp = dst;
    }
    uint32_t* target = (i == 0) ? d_bkt.as<uint32_t>() : d_tmp.as<uint32_t>();
    rc = msm_core<C>(as.dev + (size_t)off * S::BYTES, cp, cn, pl, 1, true, &sub, nullptr, s, prof, target, MSM_ACC);
    if (rc == B200_SUCCESS && i > 0) {
      k_bucket_merge<F><<<(unsigned)((n_buckets + MSM_THREADS - 1) / MSM_THREADS), MSM_THREADS, 0, s>>>(d_bkt.as<uint32_t>(), d_tmp.as<uint32_t>(), n_buckets);
      B200_LAUNCHED(1);
      if (cudaGetLastError() != cudaSuccess) rc = B200_UNKNOWN_ERROR;
      prof.mark("merge");
    }
  }
  if (rc == B200_SUCCESS) rc = msm_core<C>(nullptr, nullptr, max_chunk, pl, 1, true, &sub, d_res, s, prof, d_bkt.as<uint32_t>(), MSM_RED);
  prof.mark("final");
  prof.finish(any_host ? "msm_pipelined" : "msm_chunked");
  pageable.join();
  if (pageable.failed.load()) rc = B200_COPY_FAILED;
  // host sources: this call lasts ~100+ ms; wait for it without spinning (ranks share the container's CPU quota)
  if (rc == B200_SUCCESS && any_host && !(cfg->are_results_on_device && cfg->is_async) && stream_sync_blocking(s) != cudaSuccess) rc = B200_SYNCHRONIZATION_FAILED;
  if (rc == B200_SUCCESS) rc = finish_out(results, d_res, (size_t)PW * 4, cfg->are_results_on_device, cfg->is_async, s);
  if (rc != B200_SUCCESS || any_host) {
    // host sources: the staging areas (Scratch, freed on `s`) must outlive the copies of the private streams
    if (direct_copies) cudaStreamSynchronize(cs);
    for (CopierCtx* c : pageable.ctx) cudaStreamSynchronize(c->st);
    if (rc != B200_SUCCESS) cudaStreamSynchronize(s);
  }
  return rc;
}

template <class C>
int msm_impl(const void* scalars, const void* bases, int msm_size, const b200_msm_config* cfg, void* results)
{
  typedef typename C::Scalar S;
  typedef typename C::Base F;
  typedef typename base_fp<F>::type B;
  constexpr int AW = 2 * F::N, XW = 4 * F::N, PW = 3 * F::N;
  cudaStream_t s = (cudaStream_t)cfg->stream;
  const int batch = cfg->batch_size > 0 ? cfg->batch_size : 1;
  if (msm_size <= 0) return B200_INVALID_ARGUMENT;
  const MsmPlan pl = make_plan<C>(msm_size, cfg);
  const uint32_t n = (uint32_t)msm_size;
  if ((uint64_t)n * pl.pf >= (1ull << 31)) return B200_INVALID_ARGUMENT;
  const HostKind ks = host_kind(scalars, cfg->are_scalars_on_device), kp = host_kind(bases, cfg->are_points_on_device);
  if (batch == 1) {
    const bool any_host = (ks != HK_DEVICE) || (kp != HK_DEVICE);
    const bool pipeline = any_host && n >= pipeline_min_points() && tune(T_MSM_NO_PIPELINE) <= 0;
    const bool too_big = !any_host && n > device_chunk_points(pl);
    if (pipeline || too_big || (any_host && n > device_chunk_points(pl))) return msm_chunked<C>(scalars, bases, n, cfg, results, ks, kp);
  }
  const bool shared = cfg->are_points_shared_in_batch || batch == 1;
  const uint64_t n_points = (uint64_t)n * pl.pf * (shared ? 1 : batch);
  int err;

  // ---- inputs --------------------------------------------------------------------------------------------------------
  Scratch s_scal, s_pts, s_pts_m, s_res;
  const void *d_scal, *d_pts;
  void* d_res;
  if ((err = stage_in(d_scal, scalars, (size_t)n * batch * S::BYTES, ks == HK_DEVICE, s, s_scal))) return err;
  if ((err = stage_in(d_pts, bases, (size_t)n_points * AW * 4, kp == HK_DEVICE, s, s_pts))) return err;
  if ((err = stage_out(d_res, results, (size_t)batch * PW * 4, cfg->are_results_on_device, s, s_res))) return err;
  const uint32_t* pts_m = (const uint32_t*)d_pts;
  if (!cfg->are_points_montgomery_form) {
    uint32_t* dst;
    if (d_pts != bases) {
      dst = s_pts.as<uint32_t>(); // our own staging copy: convert in place
    } else {
      if ((err = s_pts_m.alloc((size_t)n_points * AW * 4, s))) return err;
      dst = s_pts_m.as<uint32_t>();
    }
    const uint64_t ncoord = n_points * 2 * (F::N / B::N);
    unsigned g = (unsigned)std::min<uint64_t>((ncoord + 255) / 256, (uint64_t)num_sms() * 16);
    k_to_mont<B><<<g, 256, 0, s>>>((const uint32_t*)d_pts, dst, ncoord); B200_LAUNCHED(1);
    B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    pts_m = dst;
  }

  StageTimer prof;
  prof.begin(s);
  if ((err = msm_core<C>(d_scal, pts_m, n, pl, batch, shared, cfg, d_res, s, prof))) return err;
  prof.mark("final");
  prof.finish("msm");
  return finish_out(results, d_res, (size_t)batch * PW * 4, cfg->are_results_on_device, cfg->is_async, s);
}

template <class C>
int precompute_impl(const void* in, int n, const b200_msm_config* cfg, void* out)
{
  typedef typename C::Base F;
  constexpr int AW = 2 * F::N;
  cudaStream_t s = (cudaStream_t)cfg->stream;
  if (n <= 0) return B200_INVALID_ARGUMENT;
  const MsmPlan pl = make_plan<C>(n, cfg);
  const uint32_t shift = (uint32_t)pl.c * pl.nbm;
  Scratch s_in, s_out;
  const void* d_in;
  void* d_out;
  int err;
  const size_t bytes_in = (size_t)n * AW * 4, bytes_out = bytes_in * pl.pf;
  if ((err = stage_in(d_in, in, bytes_in, cfg->are_points_on_device, s, s_in))) return err;
  if ((err = stage_out(d_out, out, bytes_out, cfg->are_results_on_device, s, s_out))) return err;
  k_precompute<F><<<(n + MSM_THREADS - 1) / MSM_THREADS, MSM_THREADS, 0, s>>>(
    (const uint32_t*)d_in, (uint32_t)n, (uint32_t)pl.pf, shift, cfg->are_points_montgomery_form, cfg->are_points_montgomery_form, (uint32_t*)d_out); B200_LAUNCHED(1);
  B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  return finish_out(out, d_out, bytes_out, cfg->are_results_on_device, cfg->is_async, s);
}


// K12: sum of n homogeneous projective points given in standard form (multi-GPU combine of per-GPU partial MSM results
// after the NCCL all-gather; the reference has no multi-device reduction -- docs/docs/start/architecture/multi-device.md).
template <class F>
__global__ void k_proj_sum(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out)
{
  constexpr int PW = 3 * F::N;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  XYZZ<F> acc = XYZZ<F>::inf();
  for (uint32_t i = 0; i < n; i++) {
    Projective<F> p = {load_el<F>(in + (uint64_t)i * PW).to_mont(), load_el<F>(in + (uint64_t)i * PW + F::N).to_mont(),
                       load_el<F>(in + (uint64_t)i * PW + 2 * F::N).to_mont()};
    XYZZ<F> q = XYZZ<F>::from_projective(p);
    acc.add(q);
  }
  Projective<F> r = acc.to_projective();
  store_el(out, r.x.from_mont());
  store_el(out + F::N, r.y.from_mont());
  store_el(out + 2 * F::N, r.z.from_mont());
}

template <class C>
int ec_sum_impl(const void* points, int n, const b200_vec_ops_config* cfg, void* out)
{
  typedef typename C::Base F;
  constexpr int PW = 3 * F::N;
  cudaStream_t s = (cudaStream_t)cfg->stream;
  if (n <= 0) return B200_INVALID_ARGUMENT;
  Scratch s_in, s_out;
  const void* d_in;
  void* d_out;
  int err;
  if ((err = stage_in(d_in, points, (size_t)n * PW * 4, cfg->is_a_on_device, s, s_in))) return err;
  if ((err = stage_out(d_out, out, (size_t)PW * 4, cfg->is_result_on_device, s, s_out))) return err;
  k_proj_sum<F><<<1, 32, 0, s>>>((const uint32_t*)d_in, (uint32_t)n, (uint32_t*)d_out); B200_LAUNCHED(1);
  B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
  return finish_out(out, d_out, (size_t)PW * 4, cfg->is_result_on_device, cfg->is_async, s);
}

}} // namespace b200::msm
